/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * SimplexSearch.hpp -- SimplexSearch (the reference's Nelder-Mead class, /root/reference/Source/SimplexSearch.hpp:23-62) behind the
 * unchanged API: setObjPtr / setSimplexParams / findMin(X, f0, fOpt).
 *
 * the method is sequential (one or two evaluations per iteration) and its control flow stays on the host; what
 * runs on the device is every objective evaluation -- single points (reflection, expansion, contraction) and the batched
 * evaluateVariableSet (the n + 1 vertices at the start and after a shrink, Source/SimplexSearch.cpp:258-266) -- through
 * pnol_eval_batch on the objective's device twin. There is no CPU fallback: an objective without a device functor throws.
 *
 * The start simplex of the reference is X0 + initRandMax * (timeRand() - 0.5) * 2 after srand(time(0)) (:57-64). Here the draws
 * come from the runtime's random stream (Runtime::setRandomStream, the same host-supplied stream the genetic algorithms use), so
 * that a run can be compared with the reference draw for draw; streamPosition() is the number of draws consumed.
 */
#ifndef PNOL_SIMPLEXSEARCH_HPP_
#define PNOL_SIMPLEXSEARCH_HPP_

#include <vector>

#include "PNOL_Algorithm.hpp"

class SimplexSearch : public Algorithm {
  private:
	// simplex parameters (Source/SimplexSearch.hpp:27-33)
	double alpha, rho, gamma, sigma;
	double initRandMax, xMinDiff;
	int maxIter;
	bool verbose;
	int iterations_;
	unsigned long long streamPos_;

  public:
	// main optimization function (Source/SimplexSearch.cpp:13-226)
	void findMin( std::vector <double> & X, double & f0, double & fOpt );

	// evaluate one point / all vertices (Source/SimplexSearch.cpp:243-266); vertices are rows of a dense row-major block
	double evaluateVariableArray( double * x, std::vector <double> & X );
	void evaluateVariableSet( double ** xvec, int Nsimplex, std::vector <double> & X, double * fvec );

	void setSimplexParams( double alphaIn, double gammaIn, double rhoIn, double sigmaIn, int maxIterIn,
			double initRandMaxIn, double xMinDiffIn, bool verboseIn )
	{ alpha = alphaIn; gamma = gammaIn; rho = rhoIn; sigma = sigmaIn;
	  maxIter = maxIterIn; initRandMax = initRandMaxIn; xMinDiff = xMinDiffIn; verbose = verboseIn; }

	int iterations() const { return iterations_; }
	unsigned long long streamPosition() const { return streamPos_; }

	SimplexSearch()
	{
		alpha = 1.0; gamma = 2.0; rho = 0.5; sigma = 0.5;
		maxIter = 10000;
		initRandMax = 1;
		xMinDiff = 1e-7;
		verbose = 0;
		iterations_ = 0;
		streamPos_ = 0;
	}
	~SimplexSearch(){}
};

// ascending sort of the vertices by function value (Source/SimplexSearch.cpp:272-326) and the largest coordinate distance of any
// vertex from the best one (:330-349)
void simplexSort( double * fvec, double ** xvec, int Nd );
double simplexDiff( double ** xvec, int Nd, int Nsimplex );

#endif /* PNOL_SIMPLEXSEARCH_HPP_ */
