/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * BFGS_bnd_linesearch.hpp -- BFGS_Bnd: the serial box-bounded BFGS of /root/reference/Source/BFGS_bnd_linesearch.hpp:30-146
 * (same setParams order, defaults :125-141, setGradVec / setinitialScalingVec).
 *
 * In the reference this class and BFGS_Bnd_MPI_SW carry word-for-word copies of mainBFGSLoop and boundaryAssessment
 * (Source/BFGS_bnd_linesearch.cpp:116-205, 521-735 vs Source/BFGS_bnd_linesearch_MPI_SW.cpp:116-207, 741-967); they differ in
 * the line search only (one trial step at a time here, pools there). It is therefore implemented ON the SW class with its
 * serial line search switched on; every evaluation still runs on the device.
 */
#ifndef PNOL_BFGS_BND_LINESEARCH_HPP_
#define PNOL_BFGS_BND_LINESEARCH_HPP_

#include "BFGS_bnd_linesearch_MPI_SW.hpp"

class BFGS_Bnd : public AlgorithmBnd {
  private:
	BFGS_Bnd_MPI_SW impl;

  public:
	void findMinBnd( vector <double> & X, vector <double> & Xlb, vector <double> & Xub, double & f0, double & fOpt )
	{
		pnol::LocalScope serial;         // the serial class never touches the communicator (Source/BFGS_bnd_linesearch.cpp)
		impl.setObjPtr( *objPtr );
		impl.findMinBnd( X, Xlb, Xub, f0, fOpt );
	}

	void setParams( double c1In, double c2In, double dalphaIn, double alphaGuessIn, double alphaTolIn, double alphaMultIn,
			int maxIterLineSearchIn, double bndTolIn, double dXGradIn, double dXHessIn, double maxIterIn,
			double xMinDiffIn, double minGrad2NormIn, bool initHessFDIn, int verboseIn )
	{
		impl.setParams( c1In, c2In, dalphaIn, alphaGuessIn, alphaTolIn, alphaMultIn, maxIterLineSearchIn, bndTolIn, dXGradIn, dXHessIn,
				maxIterIn, xMinDiffIn, minGrad2NormIn, initHessFDIn, verboseIn );
	}
	void setGradVec( vector <double> & dXGradVecIn ){ impl.setGradVec( dXGradVecIn ); }
	void setinitialScalingVec( vector <double> & initialScalingVecIn ){ impl.setinitialScalingVec( initialScalingVecIn ); }
	int iterations() const { return impl.iterations(); }

	BFGS_Bnd()
	{
		impl.setSerialLineSearch( true );
		// Source/BFGS_bnd_linesearch.hpp:125-141
		impl.setParams( 1e-4, 0.9, 1e-6, 1, 1e-20, 2, 50, 1e-5, 1e-6, 1e-3, 10000, 1e-5, 1e-5, false, 0 );
	}
	~BFGS_Bnd(){}
};

#endif
