/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
// SimplexSearch.cpp -- the Nelder-Mead controller of Source/SimplexSearch.cpp:13-349 on the host; every objective evaluation is a
// pnol_eval_batch on the objective's device twin (one point, or all n + 1 vertices in one launch).
#include "pnol/SimplexSearch.hpp"
#include "pnol/UtilityFunctions.hpp"

#include <cmath>
#include <iostream>

using std::vector;

namespace {

// the vertices live in ONE dense row-major block so that evaluateVariableSet is a single batched device call; rows[] are the
// double** view the reference's helper signatures use
struct SimplexStore {
	int nd, ns;
	vector<double> block;
	vector<double *> rows;
	SimplexStore( int Ndim ) : nd( Ndim ), ns( Ndim + 1 ), block( (size_t) (Ndim + 1)*Ndim, 0.0 ), rows( Ndim + 1 )
	{
		for( int k = 0; k < ns; k++ ) rows[k] = block.data() + (size_t) k*nd;
	}
};

// the k-th draw of the stream `s` (explicit array or counter stream)
double nextUniform( const pnol_stream_desc & s, unsigned long long & pos )
{
	double u;
	if( s.values )
	{
		if( pos >= s.n_values ) throw pnol::Error( PNOL_ERR_STREAM, "SimplexSearch::findMin: the explicit random stream is exhausted" );
		u = s.values[pos];
	}
	else u = pnol_stream_uniform( s.seed, pos, s.scale );
	pos++;
	return u;
}

} // namespace

double SimplexSearch::evaluateVariableArray( double * x, vector <double> & X )
{
	// set variable (:246-249): X is the caller's vector, as in the reference it ends up holding the last evaluated point
	for( size_t j = 0; j < X.size(); j++ ) X[j] = x[j];
	pnol::Runtime & rt = pnol::Runtime::instance();
	pnol_functor * f = objPtr->deviceFunctor();
	if( !f ) throw pnol::Error( PNOL_ERR_NO_FUNCTOR, "SimplexSearch: the objective has no device functor (no CPU fallback)" );
	double fval = 0;
	const int n = (int) X.size();
	rt.check( pnol_eval_batch( rt.ctx(), f, X.data(), 1, n, n, nullptr, &fval ) );
	objPtr->noteDeviceEvaluations( 1 );
	return fval;
}

void SimplexSearch::evaluateVariableSet( double ** xvec, int Nsimplex, vector <double> & X, double * fvec )
{
	pnol::Runtime & rt = pnol::Runtime::instance();
	pnol_functor * f = objPtr->deviceFunctor();
	if( !f ) throw pnol::Error( PNOL_ERR_NO_FUNCTOR, "SimplexSearch: the objective has no device functor (no CPU fallback)" );
	const int n = (int) X.size();
	// rows of one dense block (the layout findMin uses) go to the device as they are; anything else is packed first
	bool dense = true;
	for( int k = 1; k < Nsimplex; k++ ) if( xvec[k] != xvec[0] + (size_t) k*n ) dense = false;
	vector<double> packed;
	const double * pts = xvec[0];
	if( !dense )
	{
		packed.resize( (size_t) Nsimplex*n );
		for( int k = 0; k < Nsimplex; k++ ) for( int j = 0; j < n; j++ ) packed[(size_t) k*n + j] = xvec[k][j];
		pts = packed.data();
	}
	rt.check( pnol_eval_batch( rt.ctx(), f, pts, Nsimplex, n, n, nullptr, fvec ) );
	objPtr->noteDeviceEvaluations( Nsimplex );
	// the reference evaluates the vertices one after the other through X (:262-264): X holds the last one afterwards
	for( int j = 0; j < n; j++ ) X[j] = xvec[Nsimplex - 1][j];
}

void SimplexSearch::findMin( vector <double> & X, double & f0, double & fOpt )
{
	pnol::Runtime & rt = pnol::Runtime::instance();
	const int Ndim = (int) X.size();
	const int Nsimplex = Ndim + 1;       // n + 1 points in the simplex (:19-20)
	if( Ndim < 1 ) throw pnol::Error( PNOL_ERR_INVALID, "SimplexSearch::findMin: empty parameter vector" );

	SimplexStore S( Ndim );
	double ** xvec = S.rows.data();
	vector<double> xbar( Ndim ), xref( Ndim ), xe( Ndim ), xc( Ndim ), fvec( Nsimplex );
	double fref, fe, fc;

	// first vertex = the start point, the others = start point + uniform noise (:47-65)
	for( int j = 0; j < Ndim; j++ ) xvec[0][j] = X[j];
	// without an explicit stream the reference seeds from the clock (srand((unsigned) time(0)), :57): so does the default stream
	const pnol_stream_desc stream = rt.defaultStream( 1.0 );
	unsigned long long pos = 0;
	for( int i = 1; i < Nsimplex; i++ )
		for( int j = 0; j < Ndim; j++ )
			xvec[i][j] = xvec[0][j] + initRandMax * (nextUniform( stream, pos ) - 0.5) * 2.0;
	streamPos_ = pos;

	// evaluate and sort the initial set (:67-73)
	evaluateVariableSet( xvec, Nsimplex, X, fvec.data() );
	f0 = fvec[0];
	simplexSort( fvec.data(), xvec, Ndim );

	int iter = 0;
	double xdiff = xMinDiff*2;
	while( iter < maxIter && xdiff > xMinDiff )
	{
		// centroid of the n best vertices, each term divided before it is added (:86-98)
		for( int i = 0; i < Ndim; i++ ) xbar[i] = 0.0;
		for( int i = 0; i < Ndim; i++ )
			for( int k = 0; k < Ndim; k++ )
				xbar[i] = xbar[i] + xvec[k][i] / ((double) Ndim);

		// reflection (:100-107)
		for( int i = 0; i < Ndim; i++ ) xref[i] = xbar[i] + alpha*(xbar[i] - xvec[Nsimplex - 1][i]);
		fref = evaluateVariableArray( xref.data(), X );

		double * worst = xvec[Nsimplex - 1];
		if( fref >= fvec[0] && fref < fvec[Ndim - 1] )
		{
			// 1. neither the best nor the worst: take it (:112-120)
			for( int i = 0; i < Ndim; i++ ) worst[i] = xref[i];
			fvec[Nsimplex - 1] = fref;
		}
		else if( fref < fvec[0] )
		{
			// 2. better than the best: try the expansion (:123-152)
			for( int i = 0; i < Ndim; i++ ) xe[i] = xbar[i] + gamma*(xbar[i] - worst[i]);
			fe = evaluateVariableArray( xe.data(), X );
			if( fe < fref )
			{
				for( int i = 0; i < Ndim; i++ ) worst[i] = xe[i];
				fvec[Nsimplex - 1] = fe;
			}
			else
			{
				for( int i = 0; i < Ndim; i++ ) worst[i] = xref[i];
				fvec[Nsimplex - 1] = fref;
			}
		}
		else
		{
			// 3. contraction towards the centroid (:154-172) ...
			for( int i = 0; i < Ndim; i++ ) xc[i] = xbar[i] + rho*( worst[i] - xbar[i] );
			fc = evaluateVariableArray( xc.data(), X );
			if( fc < fvec[Nsimplex - 1] )
			{
				for( int i = 0; i < Ndim; i++ ) worst[i] = xc[i];
				fvec[Nsimplex - 1] = fc;
			}
			else
			{
				// ... or, when even that fails, shrink everything towards the best vertex and re-evaluate ALL vertices in one
				// batched device call (:175-189; the reference re-evaluates the unchanged best vertex too)
				for( int k = 1; k < Nsimplex; k++ )
					for( int i = 0; i < Ndim; i++ )
						xvec[k][i] = xvec[0][i] + sigma*(xvec[k][i] - xvec[0][i]);
				evaluateVariableSet( xvec, Nsimplex, X, fvec.data() );
			}
		}

		simplexSort( fvec.data(), xvec, Ndim );
		xdiff = simplexDiff( xvec, Ndim, Nsimplex );

		if( verbose == true && iter % 10 == 0 )
			std::cout << "At iter = " << iter << " the xdiff is " << xdiff << " with a minimum function evaluation of " << fvec[0] << std::endl;
		iter = iter + 1;
	}
	iterations_ = iter;

	// optimal values (:205-210)
	fOpt = fvec[0];
	for( int j = 0; j < Ndim; j++ ) X[j] = xvec[0][j];
	if( verbose == true )
	{
		std::cout << "Completed simplex search. f0 = " << f0 << ", fOpt = " << fOpt << " with variable:" << std::endl;
		std::cout << "X = "; print1DVector( X );
	}
}

// Selection of the minimum Nd + 1 times, the taken entry overwritten with twice the largest value (:283-309). Two quirks of the
// reference are kept because they decide the order on ties and for non-positive maxima: the candidate index is NOT reset between
// passes (it starts at 1 and then stays on the last taken, now overwritten, entry), and the sentinel 2 * max is only larger than
// everything when max > 0.
void simplexSort( double * fvec, double ** xvec, int Nd )
{
	const int Nsimplex = Nd + 1;
	int min = 1;
	vector<double> xtmp( (size_t) Nsimplex*Nd ), ftmp( Nsimplex );
	double fvecMax = fvec[0];
	for( int i = 1; i < Nsimplex; i++ ) if( fvec[i] > fvecMax ) fvecMax = fvec[i];      // vectorMax: first maximum
	for( int k = 0; k < Nsimplex; k++ )
	{
		for( int i = 0; i < Nsimplex; i++ )
			if( fvec[i] < fvec[min] ) min = i;
		for( int j = 0; j < Nd; j++ ) xtmp[(size_t) k*Nd + j] = xvec[min][j];
		ftmp[k] = fvec[min];
		fvec[min] = 2*fvecMax;
	}
	for( int k = 0; k < Nsimplex; k++ )
	{
		fvec[k] = ftmp[k];
		for( int j = 0; j < Nd; j++ ) xvec[k][j] = xtmp[(size_t) k*Nd + j];
	}
}

double simplexDiff( double ** xvec, int Nd, int Nsimplex )
{
	double xDiffMax = 0;
	for( int k = 1; k < Nsimplex; k++ )
		for( int j = 0; j < Nd; j++ )
		{
			const double diff = fabs( xvec[0][j] - xvec[k][j] );
			if( diff > xDiffMax ) xDiffMax = diff;
		}
	return xDiffMax;
}
