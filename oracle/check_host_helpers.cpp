// check_host_helpers.cpp -- TEST INFRASTRUCTURE. The reference takes vector2Norm, dotProd, linspace, vectorMin/Max, mod, sign,
// matrixInverse ... from an un-vendored library, so their semantics are DEFINED twice in this repo: in oracle/shim (what the verbatim
// reference is compiled against = what the golden vectors were made with) and in include/pnol/UtilityFunctions.hpp (what the
// product's host classes call). This program checks that the two definitions agree -- bit for bit where the arithmetic is the same
// sequence of operations (sums, linspace, extrema and their tie rule), to 1e-12 where the algorithm differs (LU vs Gauss-Jordan
// inverse). Built and run by tests/test_host_helpers.py; exit code 0 = agreement.
#include "UtilityFunctions/utilityFunctions.hpp"     // oracle/shim, global namespace

namespace ours {
#include "pnol/UtilityFunctions.hpp"                  // the product's helpers (std headers are already in through the shim's)
}

#include <cstdio>
#include <cstring>

static unsigned long long state = 0x9E3779B97F4A7C15ULL;
static double uni()                                   // deterministic doubles in (-1, 1) with full mantissas
{
	state = state*6364136223846793005ULL + 1442695040888963407ULL;
	return ( (double) (state >> 11) / 9007199254740992.0 )*2.0 - 1.0;
}
static bool same( double a, double b ){ return memcmp( &a, &b, sizeof a ) == 0; }
static int failures = 0;
#define CHECK(cond, what) do { if( !(cond) ){ printf( "MISMATCH: %s\n", what ); failures++; } } while( 0 )

int main()
{
	// sums: sequential, left to right
	for( int n : { 1, 2, 3, 17, 1000, 4097 } )
	{
		vector<double> a( n ), b( n );
		for( int i = 0; i < n; i++ ){ a[i] = uni()*1e3; b[i] = uni(); }
		CHECK( same( vector2Norm( a ), ours::vector2Norm( a ) ), "vector2Norm" );
		CHECK( same( dotProd( a, b ), ours::dotProd( a, b ) ), "dotProd" );
	}
	// linspace: a + i (b - a)/(N - 1), end point forced
	for( int N : { 1, 2, 3, 10, 11, 1001 } )
	{
		vector<double> u, v;
		double lo = uni(), hi = lo + 3.7 + uni();
		linspace( lo, hi, N, u );
		ours::linspace( lo, hi, N, v );
		CHECK( u.size() == v.size(), "linspace size" );
		for( size_t i = 0; i < u.size() && i < v.size(); i++ ) CHECK( same( u[i], v[i] ), "linspace value" );
	}
	// extrema: first one wins on ties
	{
		vector<double> v = { 3.0, 1.0, 7.0, 1.0, 7.0, 2.0 };
		double a, b; int ia, ib;
		vectorMin( v, (int) v.size(), a, ia ); ours::vectorMin( v, (int) v.size(), b, ib );
		CHECK( same( a, b ) && ia == ib && ia == 1, "vectorMin tie rule" );
		vectorMax( v, (int) v.size(), a, ia ); ours::vectorMax( v, (int) v.size(), b, ib );
		CHECK( same( a, b ) && ia == ib && ia == 2, "vectorMax tie rule" );
		for( int rep = 0; rep < 50; rep++ )
		{
			vector<double> w( 1 + rep );
			for( size_t i = 0; i < w.size(); i++ ) w[i] = (double) ( (int) ( uni()*4 ) );      // many ties
			vectorMin( w, (int) w.size(), a, ia ); ours::vectorMin( w, (int) w.size(), b, ib );
			CHECK( same( a, b ) && ia == ib, "vectorMin random" );
			vectorMax( w, (int) w.size(), a, ia ); ours::vectorMax( w, (int) w.size(), b, ib );
			CHECK( same( a, b ) && ia == ib, "vectorMax random" );
		}
	}
	// scalar helpers
	for( int a = -7; a <= 7; a++ ) for( int b : { 1, 2, 3, 5 } ) CHECK( mod( a, b ) == ours::mod( a, b ), "mod" );
	for( double x : { -2.5, -0.0, 0.0, 1e-300, 3.0 } ) CHECK( same( sign( x ), ours::sign( x ) ), "sign" );
	// inverse: LU with partial pivoting (shim) vs Gauss-Jordan with partial pivoting (product header)
	for( int n : { 1, 2, 4, 9, 30 } )
	{
		vector<vector<double> > A( n, vector<double>( n ) ), X( n, vector<double>( n ) ), Y( n, vector<double>( n ) );
		for( int i = 0; i < n; i++ ) for( int j = 0; j < n; j++ ) A[i][j] = uni() + ( i == j ? 2.0 + n : 0.0 );
		if( n >= 2 ){ A[0].swap( A[1] ); }                      // forces a row exchange
		matrixInverse( A, X );
		ours::matrixInverse( A, Y );
		double worst = 0, scale = 0;
		for( int i = 0; i < n; i++ ) for( int j = 0; j < n; j++ ){ worst = fmax( worst, fabs( X[i][j] - Y[i][j] ) ); scale = fmax( scale, fabs( X[i][j] ) ); }
		CHECK( worst <= 1e-12*scale, "matrixInverse" );
		vector<vector<double> > I1( n, vector<double>( n, 5.0 ) ), I2( n, vector<double>( n, 5.0 ) );
		setIdentity( I1 ); ours::setIdentity( I2 );
		CHECK( I1 == I2, "setIdentity" );
	}
	printf( failures ? "%d mismatches\n" : "host helpers agree (%d mismatches)\n", failures );
	return failures ? 1 : 0;
}
