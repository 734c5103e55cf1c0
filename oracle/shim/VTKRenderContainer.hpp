// TEST INFRASTRUCTURE (oracle shim) -- never linked into the product library.
// Stub of the un-vendored VTKFunctionLibrary render container used only by the serial GA's optional
// live plot (Source/GeneticAlgorithm.cpp:14-19,247-260,285-291). Every method is a no-op.
#ifndef PNOL_ORACLE_SHIM_VTKRENDERCONTAINER_HPP_
#define PNOL_ORACLE_SHIM_VTKRENDERCONTAINER_HPP_
#include <vector>
struct ShimVtkWindow { void SetSize(int, int) {} void Render() {} };
struct ShimVtkRenderer { void RemoveAllViewProps() {} };
struct ShimVtkInteractor { void Start() {} };
class RenderContainerVTK {
  public:
	ShimVtkWindow windowObj; ShimVtkRenderer rendererObj; ShimVtkInteractor interactorObj;
	ShimVtkWindow * renderWindow; ShimVtkRenderer * renderer; ShimVtkInteractor * renderWindowInteractor;
	RenderContainerVTK() : renderWindow(&windowObj), renderer(&rendererObj), renderWindowInteractor(&interactorObj) {}
	template <typename... Args> void plotPoint(Args...) {}
};
template <typename... Args> inline void plotParabolicSurf(Args...) {}
#endif
