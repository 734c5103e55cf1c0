/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * LevenbergMarquardt.hpp -- LevMarq (the reference's serial twin, /root/reference/Source/LevenbergMarquardt.hpp:24-58).
 * The reference's two classes differ only in which stencil they call and in rank-0 print guards
 * (SURVEY.md section 2); here both run the same device path.
 */
#ifndef PNOL_LEVENBERGMARQUARDT_HPP_
#define PNOL_LEVENBERGMARQUARDT_HPP_

#include "LevenbergMarquardtMPI.hpp"

class LevMarq : public MultiAlgorithm {
  private:
	double lambda0;
	double dXGrad;
	double xMinDiff;
	int maxIter;
	double lambdaFactor;
	int verbose;
	pnol::LMReport report;

  public:
	void findMin( vector <double> & X, vector <double> & f0, vector <double> & fOpt )
	{ pnol::LocalScope serial;   /* the serial class never touches the communicator (Source/LevenbergMarquardt.cpp) */
	  pnol::lmFindMin( mObjPtr, lambda0, lambdaFactor, dXGrad, maxIter, xMinDiff, verbose, X, f0, fOpt, report ); }

	void setParams( double lambda0In, double lambdaFactorIn, double dXGradIn, double maxIterIn, double xMinDiffIn, int verboseIn )
	{  maxIter = maxIterIn; xMinDiff = xMinDiffIn; verbose = verboseIn; dXGrad = dXGradIn; lambda0 = lambda0In; lambdaFactor = lambdaFactorIn; }

	const pnol::LMReport & lastReport() const { return report; }

	LevMarq()
	{
		dXGrad = 1e-7;
		lambda0 = 0.001;
		maxIter = 10000;
		xMinDiff = 1e-7;
		verbose = 1;
		lambdaFactor = 10;
	}
	~LevMarq(){}
};

#endif
