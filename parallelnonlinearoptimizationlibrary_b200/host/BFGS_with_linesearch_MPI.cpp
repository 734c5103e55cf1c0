/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
// BFGS_with_linesearch_MPI.cpp -- BFGS_MPI: BFGS with the pooled secant line search of
// Source/BFGS_with_linesearch_MPI.cpp. The pool of step lengths the reference spreads over MPI ranks
// (evalAlphaPoolMPI, :163-223) is one batched kernel launch here.
#include "pnol/BFGS_with_linesearch_MPI.hpp"

#include <cmath>
#include <iostream>

double BFGS_MPI::lineSearchObj( double alpha, vector <double> & X, vector <double> & p )
{
	vector <double> Xalphap( X.size(), 0 );
	for( size_t i = 0; i < X.size(); i++ ) Xalphap[i] = X[i] + alpha*p[i];
	return objPtr->objEval( Xalphap );
}

// Source/BFGS_with_linesearch_MPI.cpp:163-223 (no NaN sentinel in this variant: raw values are returned)
void BFGS_MPI::evalAlphaPoolMPI( vector <double> & alphaPool, vector <double> & phiPool, vector <double> & X, vector <double> & p )
{
	pnol::Runtime & rt = pnol::Runtime::instance();
	pnol_functor * f = objPtr->deviceFunctor();
	if( !f ) throw pnol::Error( PNOL_ERR_NO_FUNCTOR, "BFGS_MPI: the objective has no device functor (no CPU fallback)" );
	int bad = 0;
	rt.check( pnol_alpha_pool( rt.ctx(), f, X.data(), p.data(), (int) X.size(), alphaPool.data(), (int) phiPool.size(), 0.0, nullptr,
			nullptr, nullptr, (int) X.size(), phiPool.data(), nullptr, &bad ) );
	objPtr->noteDeviceEvaluations( (long long) phiPool.size() );
	if( bad > 0 )
		for( size_t i = 0; i < phiPool.size(); i++ )
			if( phiPool[i] == 1e10 ) phiPool[i] = lineSearchObj( alphaPool[i], X, p );
}

// Source/BFGS_with_linesearch_MPI.cpp:12-142
void BFGS_MPI::findMin( vector <double> & X, double & f0, double & fOpt )
{
	int Nparam = (int) X.size();
	vector<double> Xprev( Nparam, 0 );
	vector<double> dX( Nparam, dXGrad );
	vector<double> dFdX( Nparam, 0 );
	vector<double> dFdX_prev( Nparam, 0 );
	vector<double> p( Nparam, 0 ), s( Nparam, 0 ), g( Nparam, 0 );

	pnol::InverseHessian D( Nparam );
	if( initHessFD ) D.setFromInverseOfFDHessian( objPtr, X, dXHess );
	else D.setIdentity();

	objPtr->gradientApproximationMPI( X, dX, dFdX );                          // (:64)
	for( int i = 0; i < Nparam; i++ ) dFdX_prev[i] = dFdX[i];
	double F = objPtr->objEval( X );
	f0 = F;

	int iter = 0;
	double xdiff = xMinDiff*2;
	double grad2Norm = 2*minGrad2Norm;
	while( iter < maxIter && xdiff > xMinDiff && grad2Norm > minGrad2Norm )      // (:74)
	{
		for( int i = 0; i < Nparam; i++ ) dFdX_prev[i] = dFdX[i];
		D.direction( dFdX, p );                                              // (:81-82)

		double alpha, Fopt;
		secantLineSearch( X, F, dFdX, p, alpha, Fopt );                      // (:87)

		for( int i = 0; i < Nparam; i++ )
		{
			Xprev[i] = X[i];
			X[i] = X[i] + alpha*p[i];
		}
		F = Fopt;

		objPtr->gradientApproximationMPI( X, dX, dFdX );                      // (:101)
		for( int i = 0; i < Nparam; i++ )
		{
			s[i] = alpha*p[i];
			g[i] = dFdX[i] - dFdX_prev[i];
		}
		D.update( g, s );                                                    // (:109)

		xdiff = 0;
		for( int i = 0; i < Nparam; i++ ) xdiff += fabs( X[i] - Xprev[i] );
		grad2Norm = vector2Norm( dFdX );
		if( verbose == true )
		{
			cout << "At iter = " << iter << " the mean abs xdiff is " << xdiff << " and the grad2norm = " << grad2Norm << endl;
			cout << " with a minimum function evaluation of " << F << endl;
		}
		iter = iter + 1;
	}
	iterationsDone = iter;
	fOpt = F;
	if( verbose == true )                                                    // (:130-138)
	{
		cout << endl << "Completed bfgs. f0 = " << f0 << ", fOpt = " << fOpt << " with variable:" << endl;
		cout << "X = "; print1DVector( X );
	}
}

// Source/BFGS_with_linesearch_MPI.cpp:226-492
void BFGS_MPI::secantLineSearch( vector <double> & X, double FX,
		vector <double> & dFdX, vector <double> & p, double & alphaOpt, double & Fopt )
{
	// As many evaluations as "available processors" (:235)
	int Npool = poolWidth > 0 ? poolWidth : pnol::Runtime::instance().poolWidth();
	vector<double> alphaPool( Npool, 0 );
	vector<double> phiPool( Npool, 0 );
	vector<double> alphaPoolPrev( Npool, -1 );
	vector<double> phiPoolPrev( Npool, 0 );
	vector<double> poolSecantSlope( Npool, 0 );

	alphaOpt = 0;
	Fopt = FX;

	double alpha0 = 0;
	double phi0 = FX;
	double dphi0dalpha = dotProd( dFdX, p );

	// Initial pool (:253-261): geometric around alphaGuess
	int idxMin = -ceil( (Npool-1.0)/2.0 );
	int idxMax = floor( (Npool-1.0)/2.0 );
	double r = pow( maxAlphaMult, 1.0/(double) idxMax );
	int idx = idxMin;
	for( int k = 0; k < Npool; k++ )
	{
		alphaPool[k] = alphaGuess*pow( r, idx );
		idx++;
	}

	bool firstFlag = true;
	bool zoomFlag = false;
	int iter = 0;
	while( iter < maxIterLineSearch && firstFlag )
	{
		evalAlphaPoolMPI( alphaPool, phiPool, X, p );

		// 1. sufficient decrease (:274-281)
		for( int i = 0; i < Npool; i++ )
			if( phiPool[i] > phi0 + c1*alphaPool[i]*dphi0dalpha ) { zoomFlag = true; firstFlag = false; }

		// 2. curvature against secant slopes (:284-300)
		poolSecantSlope[0] = ( phiPool[0] - phi0 )/( alphaPool[0] - alpha0 );
		for( int i = 1; i < Npool; i++ )
			poolSecantSlope[i] = ( phiPool[i] - phiPool[i-1] )/( alphaPool[i] - alphaPool[i-1] );
		if( firstFlag )
			for( int i = 0; i < Npool; i++ )
				if( fabs(poolSecantSlope[i]) <= fabs( c2*dphi0dalpha ) ) { zoomFlag = false; firstFlag = false; }

		// 3. positive secant slopes: zoom (:304-314)
		if( firstFlag )
			for( int i = 0; i < Npool; i++ )
				if( poolSecantSlope[i] >= 0 ) { zoomFlag = true; firstFlag = false; }

		// 4. otherwise extend the interval (:317-330)
		if( firstFlag )
		{
			double alphaMax; int indexMax;
			vectorMax( alphaPool, (int) alphaPool.size(), alphaMax, indexMax );
			r = pow( maxAlphaMult, 1.0/(double) Npool );
			for( int i = 0; i < Npool; i++ )
			{
				alphaPoolPrev[i] = alphaPool[i];
				phiPoolPrev[i] = phiPool[i];
				double power = i+1;
				alphaPool[i] = alphaMax*pow( r, power );
			}
		}
		iter++;
	}

	// pool bounds (:362-381)
	double alpha_lo, alpha_hi, phi_lo, phi_hi;
	if( alphaPoolPrev[0] < 0 )
	{
		findPoolBounds( alphaPool, phiPool, alpha0, phi0, alpha_lo, alpha_hi, phi_lo, phi_hi );
	}
	else
	{
		vector<double> alphaPoolEval( Npool*2, 0 );
		vector<double> phiPoolEval( Npool*2, 0 );
		for( int i = 0; i < Npool; i++ )
		{
			alphaPoolEval[i] = alphaPoolPrev[i];
			alphaPoolEval[i+Npool] = alphaPool[i];
			phiPoolEval[i] = phiPoolPrev[i];
			phiPoolEval[i+Npool] = phiPool[i];
		}
		findPoolBounds( alphaPoolEval, phiPoolEval, alpha0, phi0, alpha_lo, alpha_hi, phi_lo, phi_hi );
	}

	// second loop: zoom (:389-473)
	iter = 0;
	vector <double> alphaPool2( Npool+2, 0 );
	vector <double> phiPool2( Npool+2, 0 );
	while( iter < maxIterLineSearch && zoomFlag )
	{
		linspace( alpha_lo, alpha_hi, Npool+2, alphaPool2 );
		phiPool2[0] = phi_lo;
		phiPool2[Npool+1] = phi_hi;
		alphaPool2[0] = alpha_lo;
		alphaPool2[Npool+1] = alpha_hi;

		for( int i = 0; i < Npool; i++ )
		{
			alphaPool[i] = alphaPool2[i+1];
			phiPool[i] = phiPool2[i+1];
		}
		evalAlphaPoolMPI( alphaPool, phiPool, X, p );
		for( int i = 0; i < Npool; i++ )
		{
			alphaPool2[i+1] = alphaPool[i];
			phiPool2[i+1] = phiPool[i];
		}

		for( int i = 0; i < Npool; i++ )
			poolSecantSlope[i] = ( phiPool2[i+1] - phiPool2[i] )/( alphaPool2[i+1] - alphaPool2[i] );
		for( int i = 0; i < Npool; i++ )
			if( fabs(poolSecantSlope[i]) <= fabs( c2*dphi0dalpha ) ) zoomFlag = false;

		if( zoomFlag )
			findPoolBounds( alphaPool2, phiPool2, alpha0, phi0, alpha_lo, alpha_hi, phi_lo, phi_hi );
		iter++;
	}

	// minimum of the last zoom pool (:477-484); before any zoom pass the pool is all zeros exactly as in the reference
	double phiMin;
	vectorMin( phiPool2, (int) phiPool2.size(), phiMin, idxMin );
	alphaOpt = alphaPool2[idxMin];
	Fopt = phiMin;
}

// Source/BFGS_with_linesearch_MPI.cpp:496-530. The reference reads alphaPool[idxMin+1] even when idxMin is the last
// index (undefined behaviour); here that read is clamped to the last element.
void findPoolBounds( vector<double> & alphaPool, vector<double> & phiPool, double alpha0, double phi0,
		double & alpha1, double & alpha2, double & phi1, double & phi2 )
{
	double phiMin;
	int idxMin;
	vectorMin( phiPool, (int) phiPool.size(), phiMin, idxMin );
	int last = (int) phiPool.size() - 1;
	int up = idxMin + 1 > last ? last : idxMin + 1;
	if( phi0 < phiMin )
	{
		alpha1 = alpha0; alpha2 = alphaPool[0];
		phi1 = phi0; phi2 = phiPool[0];
	}
	else if( phi0 >= phiMin && idxMin == 0 )
	{
		alpha1 = alpha0; alpha2 = alphaPool[up];
		phi1 = phi0; phi2 = phiPool[up];
	}
	else
	{
		alpha1 = alphaPool[idxMin-1]; alpha2 = alphaPool[up];
		phi1 = phiPool[idxMin-1]; phi2 = phiPool[up];
	}
}
