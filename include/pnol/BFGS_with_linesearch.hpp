/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * BFGS_with_linesearch.hpp -- BFGS (unbounded, cubic-interpolation strong-Wolfe line search), interface of
 * /root/reference/Source/BFGS_with_linesearch.hpp:25-105, and the two free functions every BFGS variant shares.
 * The inverse Hessian lives on the device; gradient, search direction, pool evaluations and the update are kernels.
 */
#ifndef PNOL_BFGS_WITH_LINESEARCH_HPP_
#define PNOL_BFGS_WITH_LINESEARCH_HPP_

#include <vector>

#include "UtilityFunctions.hpp"
#include "PNOL_Algorithm.hpp"

using namespace std;

class BFGS : public Algorithm {
  private:
	double c1, c2;
	double dalpha;
	double alphaGuess;
	int maxIterLineSearch;
	double dXGrad;
	double dXHess;
	double xMinDiff;
	double minGrad2Norm;
	int maxIter;
	bool initHessFD;
	bool verbose;
	int iterationsDone;

  public:
	void findMin( vector <double> & X, double & f0, double & fOpt );
	double lineSearchObj( double alpha, vector <double> & X, vector <double> & p );
	double lineSearchFDDerivative( double alpha, double phialpha, vector <double> & X, vector <double> & p );
	void lineSearchZoom( double alpha_lo, double alpha_hi, double phi_lo, double phi_hi, double dphi_lo_dalpha, double dphi_hi_dalpha,
			double phi0, double dphi0dalpha, vector <double> & X, vector <double> & p, double & alphaOpt, double & phiOpt, double & dphiOptdalpha );
	void cubicInterpolationLineSearch( vector <double> & X, double FX,
			vector <double> & dFdX, vector <double> & p, double & alphaOpt, double & Fopt );

	void setParams( double c1In, double c2In, double dalphaIn, double alphaGuessIn, int maxIterLineSearchIn, double dXGradIn, double dXHessIn, double maxIterIn,
			double xMinDiffIn, double minGrad2NormIn, bool initHessFDIn, bool verboseIn )
	{
		c1 = c1In; c2 = c2In; dalpha = dalphaIn; alphaGuess = alphaGuessIn; maxIterLineSearch = maxIterLineSearchIn;
		dXGrad = dXGradIn; dXHess = dXHessIn; maxIter = maxIterIn; xMinDiff = xMinDiffIn; minGrad2Norm = minGrad2NormIn;
		initHessFD = initHessFDIn; verbose = verboseIn;
	}
	int iterations() const { return iterationsDone; }

	BFGS()
	{
		c1 = 1e-4; c2 = 0.9; dalpha = 1e-6; alphaGuess = 1; maxIterLineSearch = 1000;
		dXGrad = 1e-6; dXHess = 1e-3; maxIter = 10000; xMinDiff = 1e-5; minGrad2Norm = 1e-5;
		verbose = 0; initHessFD = 0; iterationsDone = 0;
	}
	~BFGS(){}

  private:
	// phi(alpha) and its forward-difference slope from ONE kernel launch
	void evalPhiAndSlope( double alpha, vector <double> & X, vector <double> & p, double & phi, double & dphi );
};

// local functions (Source/BFGS_with_linesearch.hpp:107-109)
double cubicInterpMin( double alpha_lo, double alpha_hi, double phi_lo, double phi_hi, double dphi_lo_dalpha, double dphi_hi_dalpha,
		vector <double> & X, vector <double> & p );
// host-matrix form kept for source compatibility: uploads D, runs the device update, downloads D
void updateHessianInv( vector<vector<double> > & D, vector<double> & g, vector<double> & s );

namespace pnol {
// device-resident inverse Hessian used by every BFGS variant here
class InverseHessian {
  public:
	explicit InverseHessian( int n );
	int size() const { return n_; }
	void setIdentity();
	void setDiagonal( const vector<double> & d );
	void setFromHost( const vector<vector<double> > & D );
	void toHost( vector<vector<double> > & D ) const;
	void direction( const vector<double> & dFdX, vector<double> & p );          // p = -D dFdX
	void update( const vector<double> & g, const vector<double> & s );          // updateHessianInv
	void setFromInverseOfFDHessian( Objective * obj, vector<double> & X, double dXHess );
  private:
	int n_;
	DeviceArray D_;
};
}

#endif
