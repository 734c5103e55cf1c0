/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * BFGS_with_linesearch_MPI.hpp -- BFGS_MPI (pooled secant line search), interface of
 * /root/reference/Source/BFGS_with_linesearch_MPI.hpp:32-103. The reference evaluates a pool of Nprocs step lengths,
 * one per MPI rank (Source/BFGS_with_linesearch_MPI.cpp:163-223, 235); here the pool is one batched kernel launch and
 * its width is pnol::Runtime::poolWidth() (or setPoolWidth()).
 */
#ifndef PNOL_BFGS_WITH_LINESEARCH_MPI_HPP_
#define PNOL_BFGS_WITH_LINESEARCH_MPI_HPP_

#include <vector>

#include "UtilityFunctions.hpp"
#include "PNOL_Algorithm.hpp"
#include "BFGS_with_linesearch.hpp"

using namespace std;

class BFGS_MPI : public Algorithm {
  private:
	double c1, c2;
	double maxAlphaMult;
	double alphaGuess;
	int maxIterLineSearch;
	double dXGrad;
	double dXHess;
	double xMinDiff;
	double minGrad2Norm;
	int maxIter;
	bool initHessFD;
	bool verbose;
	int poolWidth;          // 0: take pnol::Runtime::poolWidth()
	int iterationsDone;

  public:
	void findMin( vector <double> & X, double & f0, double & fOpt );
	double lineSearchObj( double alpha, vector <double> & X, vector <double> & p );
	void evalAlphaPoolMPI( vector <double> & alphaPool, vector <double> & phiPool, vector <double> & X, vector <double> & p );
	void secantLineSearch( vector <double> & X, double FX,
			vector <double> & dFdX, vector <double> & p, double & alphaOpt, double & Fopt );

	void setParams( double c1In, double c2In, double maxAlphaMultIn, double alphaGuessIn, int maxIterLineSearchIn, double dXGradIn, double dXHessIn,
			double maxIterIn, double xMinDiffIn, double minGrad2NormIn, bool initHessFDIn, bool verboseIn )
	{
		c1 = c1In; c2 = c2In; maxAlphaMult = maxAlphaMultIn; alphaGuess = alphaGuessIn; maxIterLineSearch = maxIterLineSearchIn;
		dXGrad = dXGradIn; dXHess = dXHessIn; maxIter = maxIterIn; xMinDiff = xMinDiffIn; minGrad2Norm = minGrad2NormIn;
		initHessFD = initHessFDIn; verbose = verboseIn;
	}
	void setPoolWidth( int w ){ poolWidth = w; }
	int iterations() const { return iterationsDone; }

	BFGS_MPI()
	{
		c1 = 1e-4; c2 = 0.1; maxAlphaMult = 4; alphaGuess = 1; maxIterLineSearch = 1000;
		dXGrad = 1e-6; dXHess = 1e-3; maxIter = 10000; xMinDiff = 1e-5; minGrad2Norm = 1e-5;
		verbose = 0; initHessFD = 0; poolWidth = 0; iterationsDone = 0;
	}
	~BFGS_MPI(){}
};

void findPoolBounds( vector<double> & alphaPool, vector<double> & phiPool, double alpha0, double phi0,
		double & alpha_lo, double & alpha_hi, double & phi_lo, double & phi_hi );

#endif
