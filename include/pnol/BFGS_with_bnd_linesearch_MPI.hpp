/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * BFGS_with_bnd_linesearch_MPI.hpp -- BFGSBnd_MPI: the older box-bounded BFGS with the pooled SECANT line search and the
 * one-level active-set recursion, interface of /root/reference/Source/BFGS_with_bnd_linesearch_MPI.hpp:35-122 (same
 * setParams order :78-95, defaults :99-116). SURVEY.md 8(f) item 3.
 *
 * The reference evaluates one step length per MPI rank (Source/BFGS_with_bnd_linsearch_MPI.cpp:262-353, pool width = Nprocs,
 * :374); here the pool is one batched kernel launch (pnol_alpha_pool) and its width is pnol::Runtime::poolWidth() or
 * setPoolWidth(). Gradients, p = -D g and updateHessianInv are device work; every scalar decision stays on the host as
 * the reference wrote it. computeAlphaBnd / checkAlphaPoolBnd (which the reference keeps in the same .cpp, :665-743) are in
 * Box_boundary_functions.hpp.
 */
#ifndef PNOL_BFGS_WITH_BND_LINESEARCH_MPI_HPP_
#define PNOL_BFGS_WITH_BND_LINESEARCH_MPI_HPP_

#include <vector>

#include "UtilityFunctions.hpp"
#include "PNOL_Algorithm.hpp"
#include "Box_boundary_functions.hpp"
#include "BFGS_with_linesearch.hpp"
#include "BFGS_with_linesearch_MPI.hpp"

using namespace std;

class BFGSBnd_MPI : public AlgorithmBnd {
  private:
	// defaults of Source/BFGS_with_bnd_linesearch_MPI.hpp:99-116
	double c1 = 1e-4, c2 = 0.1;                    // sufficient-decrease and curvature constants of the line search
	double maxAlphaMult = 4, alphaGuess = 1;       // pool spread and centre
	int maxIterLineSearch = 1000;
	double alphaMin = 1e-16;                       // smallest step; below it both the zoom and the outer loop stop
	double dXGrad = 1e-6, dXHess = 1e-3;           // stencil steps (dXGrad is also the "on the bound" tolerance)
	double xMinDiff = 1e-5, minGrad2Norm = 1e-5;   // stop tests
	double FStepTolerance = 1e-5;                  // smaller gain than this: retry along steepest descent
	int maxIter = 10000;
	bool initHessFD = false, verbose = false;
	int poolWidth = 0;                             // 0: take pnol::Runtime::poolWidth()
	int iterationsDone = 0, poolLaunches = 0;      // of the last findMinBnd (outer + recursive iterations; alpha-pool launches)

  public:
	BFGSBnd_MPI() {}
	~BFGSBnd_MPI() {}

	// same argument order as the reference (:78-95)
	void setParams( double c1In, double c2In, double alphaMinIn, double maxAlphaMultIn, double alphaGuessIn, int maxIterLineSearchIn, double dXGradIn,
			double dXHessIn, double maxIterIn, double xMinDiffIn, double minGrad2NormIn, double FStepToleranceIn, bool initHessFDIn, bool verboseIn )
	{
		c1 = c1In; c2 = c2In; alphaMin = alphaMinIn; maxAlphaMult = maxAlphaMultIn; alphaGuess = alphaGuessIn;
		maxIterLineSearch = maxIterLineSearchIn; dXGrad = dXGradIn; dXHess = dXHessIn; maxIter = (int) maxIterIn; xMinDiff = xMinDiffIn;
		minGrad2Norm = minGrad2NormIn; FStepTolerance = FStepToleranceIn; initHessFD = initHessFDIn; verbose = verboseIn;
	}
	void setPoolWidth( int w ){ poolWidth = w; }
	int iterations() const { return iterationsDone; }
	int poolEvaluations() const { return poolLaunches; }

	// main optimization function (AlgorithmBnd)
	void findMinBnd( vector <double> & X, vector <double> & Xlb, vector <double> & Xub, double & f0, double & fOpt );

	// the pieces, public as in the reference (:60-74); D is device-resident here
	void mainBFGSLoop( double & F, vector <double> & X, vector<double> & dFdX, pnol::InverseHessian & D,
			vector <double> & Xlb, vector <double> & Xub, vector<double> & dX, vector<double> & constantX, vector<bool> & constantIndicator,
			bool & optimFlag, bool & recurFlag );
	double lineSearchObj( double alpha, vector <double> & X, vector <double> & p, vector<double> & constantX, vector<bool> & constantIndicator );
	void evalAlphaPoolMPI( vector <double> & alphaPool, vector <double> & phiPool, vector <double> & X, vector <double> & p,
			vector<double> & constantX, vector<bool> & constantIndicator );
	void secantLineSearchBnd( vector <double> & X, vector <double> & Xlb, vector <double> & Xub, double FX,
			vector <double> & dFdX, vector <double> & p, double & alphaOpt, double & Fopt, vector<double> & constantX, vector<bool> & constantIndicator );
	void boundaryAssessment( double & F, vector <double> & X, vector <double> & p, vector<double> & dFdX, pnol::InverseHessian & D,
			vector <double> & Xlb, vector <double> & Xub, vector<double> & dX, vector<double> & constantX, vector<bool> & constantIndicator,
			bool & optimFlag, bool & recurFlag );
};

#endif
