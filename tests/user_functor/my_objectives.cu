/* The ONE translation unit a user adds for objectives of his own (compiled out of tree, see tests/test_gpu_user_functor.py):
 *   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -fPIC,-ffp-contract=off -I<repo>/include -shared \
 *        -o libmy_objectives.so my_objectives.cu -L<repo>/parallelnonlinearoptimizationlibrary_b200/lib -lpnol_b200 -lpnol_b200_host
 * The three macros instantiate the library's kernel templates for the functors and register their launch tables under the kinds
 * chosen in my_objectives.hpp when this library is loaded; libpnol_b200.so itself is untouched. */
#include "my_objectives.hpp"

#include "pnol/device/functor_kernels.cuh"

PNOL_REGISTER_SCALAR_FUNCTOR(MY_F_TRID, TridFunctor, 0)
PNOL_REGISTER_SCALAR_FUNCTOR(MY_F_STYBLINSKI, StyblinskiFunctor, 0)
PNOL_REGISTER_RESIDUAL_FUNCTOR(MY_F_GAUSSFIT, GaussFitFunctor, 2)

/* ---- the plugin classes of these objectives: the reference-style user code (Objective / MultiObjective subclasses) ---- */
#include "pnol/BFGS_with_linesearch.hpp"
#include "pnol/ExampleObjectives.hpp"
#include "pnol/LevenbergMarquardt.hpp"

class TridObject : public pnol::FunctorObjective<TridFunctor> {};
class StyblinskiObject : public pnol::FunctorObjective<StyblinskiFunctor> {
  public:
	explicit StyblinskiObject(double scale) { P.scalars[0] = scale; }
};
class GaussFitObjective : public pnol::FunctorMultiObjective<GaussFitFunctor> {
  public:
	GaussFitObjective(const double * t, const double * y, long long m, double c0)
	{
		cols.resize(2);
		cols[0].assign(t, t + m);
		cols[1].assign(y, y + m);
		P.scalars[0] = c0;
		bindColumns();
	}
};

/* ---- C face for the test (ctypes): host objEval of the same functors and two drivers through the plugin classes ---- */
extern "C" {

int my_registration_status(void) { return pnol_reg_TridFunctor | pnol_reg_StyblinskiFunctor | pnol_reg_GaussFitFunctor; }

double my_host_eval(int kind, double scale, const double * x, int n)
{
	std::vector<double> X(x, x + n);
	if (kind == MY_F_TRID) { TridObject o; return o.objEval(X); }
	StyblinskiObject o(scale);
	return o.objEval(X);
}

void my_host_residual(const double * t, const double * y, long long m, double c0, const double * x, double * F)
{
	GaussFitObjective o(t, y, m, c0);
	std::vector<double> X(x, x + 3), Fv((size_t) m);
	o.objEval(X, Fv);
	for (long long i = 0; i < m; i++) F[i] = Fv[(size_t) i];
}

/* BFGS::findMin (Source/BFGS_with_linesearch.cpp:14-141) on the Trid objective: every stencil and line-search evaluation runs the
 * user's kernels. Returns the iteration status 0, -1 on a library error (text on stderr). */
int my_bfgs_trid(double * x, int n, int maxiter, double * f0, double * fopt)
{
	try {
		TridObject obj;
		BFGS bfgs;
		bfgs.setObjPtr(obj);
		bfgs.setParams(1e-4, 0.9, 1e-6, 1.0, 1000, 1e-7, 1e-3, maxiter, 1e-9, 1e-9, 0, 0);
		std::vector<double> X(x, x + n);
		bfgs.findMin(X, *f0, *fopt);
		for (int i = 0; i < n; i++) x[i] = X[(size_t) i];
		return 0;
	} catch (const std::exception & e) {
		fprintf(stderr, "my_bfgs_trid: %s\n", e.what());
		return -1;
	}
}

/* LevMarq::findMin (Source/LevenbergMarquardt.cpp:12-150) on the Gaussian fit */
int my_lm_gaussfit(const double * t, const double * y, long long m, double c0, double * x, int maxiter)
{
	try {
		GaussFitObjective obj(t, y, m, c0);
		LevMarq lm;
		lm.setObjPtr(obj);
		lm.setParams(0.001, 10, 1e-7, maxiter, 1e-12, 0);
		std::vector<double> X(x, x + 3), F0((size_t) m), F((size_t) m);
		lm.findMin(X, F0, F);
		for (int i = 0; i < 3; i++) x[i] = X[(size_t) i];
		return 0;
	} catch (const std::exception & e) {
		fprintf(stderr, "my_lm_gaussfit: %s\n", e.what());
		return -1;
	}
}

}
