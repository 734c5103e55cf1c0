/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
// BFGS_with_bnd_linsearch_MPI.cpp -- BFGSBnd_MPI: box-bounded BFGS with the pooled secant line search and a one-level
// active-set recursion. The host control flow follows Source/BFGS_with_bnd_linsearch_MPI.cpp of the
// reference decision for decision (iterates must match it); the FD gradients, p = -D g, the alpha pools and
// updateHessianInv are device work. The pool of step lengths the reference spreads over MPI ranks (:262-353) is one
// batched kernel launch here.
#include "pnol/BFGS_with_bnd_linesearch_MPI.hpp"

#include <cmath>
#include <iostream>

// Source/BFGS_with_bnd_linsearch_MPI.cpp:14-80
void BFGSBnd_MPI::findMinBnd( vector <double> & X, vector <double> & Xlb, vector <double> & Xub, double & f0, double & fOpt )
{
	iterationsDone = 0;
	poolLaunches = 0;
	int Nparam = (int) X.size();

	vector<double> constantX( Nparam, 0 );
	vector<bool> constantIndicator( Nparam, false );
	vector<double> dX( Nparam, dXGrad );
	vector<double> dFdX( Nparam, 0 );

	checkBoxBounds( X, Xlb, Xub );                                           // (:35)

	pnol::InverseHessian D( Nparam );
	if( initHessFD ) D.setFromInverseOfFDHessian( objPtr, X, dXHess );        // (:39-47)
	else D.setIdentity();

	objPtr->gradientApproximationMPI( X, dX, dFdX );                          // (:56)
	double F = objPtr->objEval( X );
	f0 = F;
	bool optimFlag = true;
	bool recurFlag = false;

	mainBFGSLoop( F, X, dFdX, D, Xlb, Xub, dX, constantX, constantIndicator, optimFlag, recurFlag );   // (:63)

	fOpt = F;
	if( verbose == true )
	{
		cout << endl << "Completed bounded bfgs." << endl;
		cout << "f0 = " << f0 << ", fOpt = " << fOpt << " with variable:" << endl;
		cout << "X = "; print1DVector( X );
		cout << "Xlb = "; print1DVector( Xlb );
		cout << "Xub = "; print1DVector( Xub );
	}
}

// Source/BFGS_with_bnd_linsearch_MPI.cpp:84-243. Called once on the full problem and once more, from boundaryAssessment,
// on the variables that stay free (X, Xlb, Xub, dX, D are then the reduced ones; constantX / constantIndicator always have
// the full length).
void BFGSBnd_MPI::mainBFGSLoop( double & F, vector <double> & X, vector<double> & dFdX, pnol::InverseHessian & D,
		vector <double> & Xlb, vector <double> & Xub, vector<double> & dX, vector<double> & constantX,
		vector<bool> & constantIndicator, bool & optimFlag, bool & recurFlag )
{
	double Fprev = 2*F;
	if( verbose == true )
	{
		cout << endl << "Starting bounded BFGS loop with." << endl;
		cout << "    X = "; print1DVector( X );
		cout << "    F(X) = " << F << endl;
		cout << "    dFdX = "; print1DVector( dFdX );
	}

	int Nparam = (int) X.size();
	vector<double> Xprev( Nparam, 0 );
	vector<double> dFdX_prev( Nparam, 0 );
	vector<double> p( Nparam, 0 ), s( Nparam, 0 ), g( Nparam, 0 );

	for( int i = 0; i < Nparam; i++ ) dFdX_prev[i] = dFdX[i];                // (:128)
	D.direction( dFdX, p );                                                  // (:131-132)

	int iter = 0;
	double xdiff = xMinDiff*2;
	double grad2Norm = 2*minGrad2Norm;
	double alpha = alphaMin*2;
	while( iter < maxIter && xdiff > xMinDiff && grad2Norm > minGrad2Norm && alpha > alphaMin && optimFlag )   // (:141)
	{
		double Fopt;
		secantLineSearchBnd( X, Xlb, Xub, F, dFdX, p, alpha, Fopt, constantX, constantIndicator );   // (:147)

		// the quasi-Newton direction gained too little: fresh gradient, steepest descent (:150-162)
		if( F - Fopt < FStepTolerance )
		{
			cout << "Line search failed in the quasi-newton direction. Recomputing gradient and attempting steepest descent instead." << endl;
			objPtr->gradientApproximationMPIRecur( X, dX, dFdX, constantX, constantIndicator );
			for( int i = 0; i < Nparam; i++ ) p[i] = -dFdX[i];
			secantLineSearchBnd( X, Xlb, Xub, F, dFdX, p, alpha, Fopt, constantX, constantIndicator );
		}

		for( int i = 0; i < Nparam; i++ )                                    // (:166-172)
		{
			Xprev[i] = X[i];
			X[i] = X[i] + alpha*p[i];
		}
		Fprev = F;
		F = Fopt;

		objPtr->gradientApproximationMPIRecur( X, dX, dFdX, constantX, constantIndicator );   // (:176)

		for( int i = 0; i < Nparam; i++ )                                    // (:180-188)
		{
			s[i] = alpha*p[i];
			g[i] = dFdX[i] - dFdX_prev[i];
		}
		if( dotProd( g, s ) != 0 ) D.update( g, s );

		for( int i = 0; i < Nparam; i++ ) dFdX_prev[i] = dFdX[i];            // (:192)
		D.direction( dFdX, p );                                              // (:195-196)

		// an increase means the finite-difference error took over: stop (:199-202)
		if( F > Fprev ) optimFlag = false;

		xdiff = 0;
		for( int i = 0; i < Nparam; i++ ) xdiff += fabs( X[i] - Xprev[i] );
		grad2Norm = vector2Norm( dFdX );

		if( verbose == true )
		{
			cout << endl << "---> At iter = " << iter << " the mean abs xdiff is " << xdiff << " and the grad2norm = " << grad2Norm << endl;
			cout << "    X = "; print1DVector( X );
			cout << "    with a minimum function evaluation of " << F << endl;
		}

		if( !recurFlag )                                                     // (:234-235)
			boundaryAssessment( F, X, p, dFdX, D, Xlb, Xub, dX, constantX, constantIndicator, optimFlag, recurFlag );

		iter = iter+1;
		iterationsDone = iterationsDone + 1;
	}
}

// Source/BFGS_with_bnd_linsearch_MPI.cpp:246-258 (single point; the pools go through evalAlphaPoolMPI)
double BFGSBnd_MPI::lineSearchObj( double alpha, vector <double> & X, vector <double> & p, vector<double> & constantX, vector<bool> & constantIndicator )
{
	vector <double> Xalphap( X.size(), 0 );
	for( size_t i = 0; i < X.size(); i++ ) Xalphap[i] = X[i] + alpha*p[i];
	return objPtr->objEvalRecur( Xalphap, constantX, constantIndicator );
}

// Source/BFGS_with_bnd_linsearch_MPI.cpp:262-353: phi[k] = f( assemble( X + alpha[k] p ) ) for the whole pool in one launch.
// The reference prints and exit(0)s when a value is NaN or infinite (:325-337); here that is a pnol::Error.
void BFGSBnd_MPI::evalAlphaPoolMPI( vector <double> & alphaPool, vector <double> & phiPool, vector <double> & X, vector <double> & p,
		vector<double> & constantX, vector<bool> & constantIndicator )
{
	int N = (int) phiPool.size();
	if( N == 0 ) return;
	pnol::Runtime & rt = pnol::Runtime::instance();
	pnol_functor * f = objPtr->deviceFunctor();
	if( !f ) throw pnol::Error( PNOL_ERR_NO_FUNCTOR, "BFGSBnd_MPI: the objective has no device functor (no CPU fallback)" );
	if( verbose == true )
	{
		cout << "Evaluating alphaPool =  [";
		for( size_t i = 0; i < alphaPool.size(); i++ ) cout << alphaPool[i] << "  ";
		cout << "]" << '\r' << flush;
	}
	vector<unsigned char> ind( constantIndicator.size() );
	for( size_t i = 0; i < ind.size(); i++ ) ind[i] = constantIndicator[i] ? 1 : 0;
	int bad = 0;
	rt.check( pnol_alpha_pool( rt.ctx(), f, X.data(), p.data(), (int) X.size(), alphaPool.data(), N, 0.0, nullptr,
			constantX.data(), ind.data(), (int) constantX.size(), phiPool.data(), nullptr, &bad ) );
	poolLaunches = poolLaunches + 1;
	objPtr->noteDeviceEvaluations( N );
	if( bad > 0 )
		throw pnol::Error( PNOL_ERR_NONFINITE, "BFGSBnd_MPI: line search crashed (objective returned NaN or inf on the alpha pool)" );
}

// Source/BFGS_with_bnd_linsearch_MPI.cpp:358-660
void BFGSBnd_MPI::secantLineSearchBnd( vector <double> & X, vector <double> & Xlb, vector <double> & Xub, double FX,
		vector <double> & dFdX, vector <double> & p, double & alphaOpt, double & Fopt, vector<double> & constantX, vector<bool> & constantIndicator )
{
	int idxMin, idxMax, idx;
	double r;

	// As many evaluations as "available processors" (:374)
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
	double dphi0dalpha = dotProd( dFdX, p );                                 // (:388)

	// Initial pool: geometric around alphaGuess (:392-402)
	bool bndIndicator = false;
	idxMin = -ceil( (Npool-1.0)/2.0 );
	idxMax = floor( (Npool-1.0)/2.0 );
	r = pow( maxAlphaMult, 1.0/(double) idxMax );
	idx = idxMin;
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
		// keep the pool inside the box: beyond the bound it becomes an even grid up to the bound (:411)
		checkAlphaPoolBnd( bndIndicator, alphaPool, X, Xlb, Xub, p, constantX, constantIndicator );

		evalAlphaPoolMPI( alphaPool, phiPool, X, p, constantX, constantIndicator );   // (:414)

		// 1. sufficient decrease (:417-424)
		for( int i = 0; i < Npool; i++ )
			if( phiPool[i] > phi0 + c1*alphaPool[i]*dphi0dalpha ) { zoomFlag = true; firstFlag = false; }

		// 2. curvature against secant slopes (:427-443)
		poolSecantSlope[0] = ( phiPool[0] - phi0 )/( alphaPool[0] - alpha0 );
		for( int i = 1; i < Npool; i++ )
			poolSecantSlope[i] = ( phiPool[i] - phiPool[i-1] )/( alphaPool[i] - alphaPool[i-1] );
		if( firstFlag )
			for( int i = 0; i < Npool; i++ )
				if( fabs(poolSecantSlope[i]) <= fabs( c2*dphi0dalpha ) ) { zoomFlag = false; firstFlag = false; }

		// 3. positive secant slopes: zoom (:447-457)
		if( firstFlag )
			for( int i = 0; i < Npool; i++ )
				if( poolSecantSlope[i] >= 0 ) { zoomFlag = true; firstFlag = false; }

		// 4. otherwise extend the interval, unless the box has been reached (:460-483)
		if( firstFlag && !bndIndicator )
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
		else if( bndIndicator )
		{
			firstFlag = false;
			if( verbose == true )
			{
				cout << endl << "Line search reached boundary. Attempting to find an acceptable point in the domain interior." << endl;
				cout << "alpha = "; print1DVector( alphaPool );
				cout << "phi = "; print1DVector( phiPool );
			}
		}
		iter++;
	}

	// pool bounds (:515-535)
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

	// second loop: zoom (:543-646)
	iter = 0;
	vector <double> alphaPool2( Npool+2, 0 );
	vector <double> phiPool2( Npool+2, 0 );
	while( iter < maxIterLineSearch && zoomFlag )
	{
		// 1. new pool locations (:569-573)
		linspace( alpha_lo, alpha_hi, Npool+2, alphaPool2 );
		phiPool2[0] = phi_lo;
		phiPool2[Npool+1] = phi_hi;
		alphaPool2[0] = alpha_lo;
		alphaPool2[Npool+1] = alpha_hi;

		// 2. evaluate the interior points (:577-589)
		for( int i = 0; i < Npool; i++ )
		{
			alphaPool[i] = alphaPool2[i+1];
			phiPool[i] = phiPool2[i+1];
		}
		evalAlphaPoolMPI( alphaPool, phiPool, X, p, constantX, constantIndicator );
		for( int i = 0; i < Npool; i++ )
		{
			alphaPool2[i+1] = alphaPool[i];
			phiPool2[i+1] = phiPool[i];
		}

		// 3. curvature against secant slopes (:594-607)
		for( int i = 0; i < Npool; i++ )
			poolSecantSlope[i] = ( phiPool2[i+1] - phiPool2[i] )/( alphaPool2[i+1] - alphaPool2[i] );
		for( int i = 0; i < Npool; i++ )
			if( fabs(poolSecantSlope[i]) <= fabs( c2*dphi0dalpha ) ) zoomFlag = false;

		if( zoomFlag )                                                       // (:609-610)
			findPoolBounds( alphaPool2, phiPool2, alpha0, phi0, alpha_lo, alpha_hi, phi_lo, phi_hi );

		// 5. the pool has shrunk below the smallest step (:638-641)
		double alphaMax2; int indexMax2;
		vectorMax( alphaPool2, (int) alphaPool2.size(), alphaMax2, indexMax2 );
		if( alphaMax2 < alphaMin ) zoomFlag = 0;

		iter++;
	}

	// minimum of the last evaluated pool (:650-654)
	double phiMin;
	vectorMin( phiPool, (int) phiPool.size(), phiMin, idxMin );
	alphaOpt = alphaPool[idxMin];
	Fopt = phiMin;
}

// The freeze rule of boundaryAssessment (Source/BFGS_with_bnd_linsearch_MPI.cpp:761-781 with v = p, outward = -1, and
// :891-912 with v = dFdX, outward = +1, after every indicator has been cleared): a free variable within tol of its lower
// bound is frozen when outward*v > 0 there, within tol of its upper bound when outward*v < 0. X, Xlb, Xub, v are indexed by
// the running count of free variables, the indicator by the full index, as in the reference.
static bool freezeAtBounds( const vector<double> & X, const vector<double> & Xlb, const vector<double> & Xub, const vector<double> & v,
		double outward, double tol, bool clearFirst, vector<double> & constantX, vector<bool> & constantIndicator )
{
	bool any = false;
	int k = 0;
	for( size_t i = 0; i < constantIndicator.size(); i++ )
	{
		if( clearFirst ) constantIndicator[i] = false;
		if( constantIndicator[i] ) continue;
		bool atLower = fabs( X[k] - Xlb[k] ) < tol && outward*v[k] > 0;
		bool atUpper = !atLower && fabs( X[k] - Xub[k] ) < tol && outward*v[k] < 0;
		if( atLower || atUpper )
		{
			any = true;
			constantIndicator[i] = true;
			constantX[i] = X[k];
		}
		k++;
	}
	return any;
}

// Source/BFGS_with_bnd_linsearch_MPI.cpp:748-934. Runs on the full problem only (mainBFGSLoop skips it inside the recursion).
void BFGSBnd_MPI::boundaryAssessment( double & F, vector <double> & X, vector <double> & p, vector<double> & dFdX, pnol::InverseHessian & D,
		vector <double> & Xlb, vector <double> & Xub, vector<double> & dX, vector<double> & constantX, vector<bool> & constantIndicator,
		bool & optimFlag, bool & recurFlag )
{
	int Ndim = (int) constantX.size();

	// freeze the variables that sit on a bound with the search direction pointing outward (:761-781)
	bool bndFlag = freezeAtBounds( X, Xlb, Xub, p, -1.0, dXGrad, false, constantX, constantIndicator );

	vector<int> freeIdx;
	for( int i = 0; i < Ndim; i++ )
		if( !constantIndicator[i] ) freeIdx.push_back( i );
	int NdimRecur = (int) freeIdx.size();                                    // Ndim - Nconst (:783-787, :809)

	if( bndFlag && verbose )
	{
		cout << endl << " Optimizer reached box boundary and found that the steepest descent is directed outside of the boundary at" << endl;
		cout << "    indicator "; print1DVector( constantIndicator );
		cout << "    the constant values are "; print1DVector( constantX );
		cout << "    X = "; print1DVector( X );
		cout << "    F = " << F << endl;
	}
	if( !bndFlag || NdimRecur == 0 ) return;                                 // only continue if at least one variable remains (:810)

	// reduced problem on the free variables; the free block of D seeds its inverse Hessian (:813-850)
	double FRecur = F;
	vector <double> XRecur( NdimRecur ), dFdXRecur( NdimRecur ), XlbRecur( NdimRecur ), XubRecur( NdimRecur ), dXRecur( NdimRecur );
	vector<vector<double> > Dfull, DRecur( NdimRecur, vector<double>( NdimRecur ) );
	D.toHost( Dfull );
	for( int a = 0; a < NdimRecur; a++ )
	{
		int i = freeIdx[a];
		XRecur[a] = X[i]; dFdXRecur[a] = dFdX[i]; XlbRecur[a] = Xlb[i]; XubRecur[a] = Xub[i]; dXRecur[a] = dX[i];
		for( int b = 0; b < NdimRecur; b++ ) DRecur[a][b] = Dfull[i][freeIdx[b]];
	}
	pnol::InverseHessian DRecurDev( NdimRecur );
	DRecurDev.setFromHost( DRecur );

	recurFlag = true;                                                        // (:853-854)
	mainBFGSLoop( FRecur, XRecur, dFdXRecur, DRecurDev, XlbRecur, XubRecur, dXRecur, constantX, constantIndicator, optimFlag, recurFlag );

	// scatter the free variables back (:858-872). F keeps the value it had before the recursion: the reference never
	// copies FRecur back, so fOpt of a run that recursed is the pre-recursion value.
	for( int a = 0; a < NdimRecur; a++ )
	{
		int i = freeIdx[a];
		X[i] = XRecur[a]; dFdX[i] = dFdXRecur[a]; Xlb[i] = XlbRecur[a]; Xub[i] = XubRecur[a]; dX[i] = dXRecur[a];
	}

	D.setIdentity();                                                         // (:875-887)
	objPtr->gradientApproximationMPI( X, dX, dFdX );                          // (:888)

	// release everything, then re-freeze what the fresh gradient still pushes out of the box (:890-917)
	freezeAtBounds( X, Xlb, Xub, dFdX, +1.0, dXGrad, true, constantX, constantIndicator );
	int Nconst = 0;
	for( size_t i = 0; i < constantIndicator.size(); i++ ) Nconst = Nconst + constantIndicator[i];

	recurFlag = false;
	optimFlag = ( Nconst == 0 );                                             // (:919-940)
	if( verbose )
		cout << ( optimFlag ? "Optimization continuing after recursive boundary optimization, as the gradient through the boundary points to the domain interior."
		                    : "Optimization exiting after recursive boundary optimization, as the gradient through the boundary still points out of the domain." ) << endl;
}
