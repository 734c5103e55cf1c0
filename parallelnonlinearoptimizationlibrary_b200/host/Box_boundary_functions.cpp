/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
// Box_boundary_functions.cpp -- box-bound helpers (host arithmetic) behind the reference's names.
#include "pnol/Box_boundary_functions.hpp"
#include "pnol/Runtime.hpp"

#include <iostream>

// Source/Box_boundary_functions.cpp:11-40: a start value outside the box by more than |bound|/1000 goes to the midpoint
void checkBoxBounds( vector <double> & X, vector <double> & Xlb, vector <double> & Xub )
{
	for( size_t i = 0; i < X.size(); i++ )
	{
		if( X[i] - Xlb[i] < -fabs(Xlb[i])/1000 || X[i] - Xub[i] > fabs(Xub[i])/1000 )
		{
			cout << endl << "!!!!----------------- WARNING -----------------!!!!" << endl;
			cout << "X[" << i << "] = " << X[i] << " is outside of bounds Xlb[i] = " << Xlb[i] << ", Xub[i] = " << Xub[i] << endl;
			X[i] = (Xlb[i] + Xub[i])/2.0;
			cout << " Replaced X[" << i << "] with " << X[i] << endl;
			cout << "!!!!----------------- END WARNING -----------------!!!!" << endl << endl;
		}
	}
}

// Source/Box_boundary_functions.cpp:44-62: random start inside the box. The reference draws from a hardware RNG;
// here the draws come from the runtime's random stream (counter-based, reproducible).
void setHardRandValues( vector <double> & X, vector <double> & Xlb, vector <double> & Xub )
{
	pnol::Runtime & rt = pnol::Runtime::instance();
	static uint64_t position = 0;
	uint64_t seed = rt.haveRandomStream() ? rt.randomStream().seed : 0x5EEDULL;
	for( size_t i = 0; i < X.size(); i++ )
		X[i] = (Xub[i] - Xlb[i])*pnol_stream_uniform( seed ^ 0xB0C5ULL, position++, 1.0 ) + Xlb[i];
	cout << " Set random values of X: "; print1DVector( X );
}

// Source/BFGS_with_bnd_linsearch_MPI.cpp:665-708
double computeAlphaBnd( vector <double> & X, vector <double> & Xlb, vector <double> & Xub, vector <double> & p )
{
	return pnol_compute_alpha_bnd( X.data(), Xlb.data(), Xub.data(), p.data(), (int) X.size() );
}

// Source/BFGS_with_bnd_linsearch_MPI.cpp:711-743
void checkAlphaPoolBnd( bool & bndIndicator, vector <double> & alphaPool, vector <double> & X, vector <double> & Xlb, vector <double> & Xub,
		vector <double> & p, vector<double> &, vector<bool> & )
{
	int Npool = (int) alphaPool.size();
	bndIndicator = false;
	double alphaBnd = computeAlphaBnd( X, Xlb, Xub, p );
	for( int i = 0; i < Npool; i++ )
		if( alphaPool[i] > alphaBnd ) bndIndicator = true;
	if( bndIndicator )
	{
		double deltaAlpha = alphaBnd/(Npool);
		for( int i = 0; i < Npool; i++ ) alphaPool[i] = deltaAlpha*(i+1);
	}
	for( int i = 0; i < Npool; i++ )
		if( alphaPool[i] < 0 ) alphaPool[i] = 0;
}
