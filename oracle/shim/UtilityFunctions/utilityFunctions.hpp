// TEST INFRASTRUCTURE (oracle shim) -- never linked into the product library.
//
// Stand-in for the reference's un-vendored, un-pinned sibling library "UtilityFunctionLibrary"
// (located only by path in /root/reference/Build/configure-pnol:6 and
// /root/reference/Source/CMakeLists.txt:29-32; its source is not on this box). Signatures are those the
// reference call sites need; the SEMANTICS BELOW ARE OURS ("parity unpinned" at this boundary, see
// DESIGN.md): sequential left-to-right sums, i-j-k matrix product, first extremum on ties, LU with
// partial pivoting (first maximal pivot), linspace a + i*(b-a)/(N-1) with the last point forced to b.
//
// timeRand() does not call rand(): it pops the next value of a host-supplied random stream (either an
// explicit array or the counter-based SplitMix64 stream of include/pnol_b200.h) so that GA runs are
// reproducible and can be compared with the device path draw for draw.
#ifndef PNOL_ORACLE_SHIM_UTILITYFUNCTIONS_HPP_
#define PNOL_ORACLE_SHIM_UTILITYFUNCTIONS_HPP_

#include <cmath>
#include <cstdlib>
#include <iostream>
#include <iomanip>
#include <string>
#include <vector>
#include <mpi.h>   // the real library evidently exposes MPI: Source/Box_boundary_functions.cpp:16 uses it with no other include

using namespace std;

// ---- dense helpers (hot path call sites: LevenbergMarquardtMPI.cpp:51,64-65,83,88,108,138;
//      BFGS_with_linesearch.cpp:397,421-422; BFGS_bnd_linesearch_MPI_SW.cpp:57,143,226) ----
double vector2Norm( vector<double> & v );
double dotProd( vector<double> & a, vector<double> & b );
void matrixTranspose( vector<vector<double> > & A, vector<vector<double> > & AT );
void matrixMultiply( vector<vector<double> > & A, vector<vector<double> > & B, vector<vector<double> > & C );
void matrixVectorMultiply( vector<vector<double> > & A, vector<double> & x, vector<double> & y );
void luSolve( vector<vector<double> > & A, vector<double> & b, vector<double> & x );
void matrixInverse( vector<vector<double> > & A, vector<vector<double> > & Ainv );
void setIdentity( vector<vector<double> > & A );

// ---- small scalar helpers ----
int mod( int a, int b );
double sign( double x );
void linspace( double a, double b, int N, vector<double> & out );
void vectorMin( vector<double> & v, int N, double & val, int & idx );
void vectorMax( vector<double> & v, int N, double & val, int & idx );
void vectorMax( double * v, int N, double & val, int & idx );
void vectorMin( double * v, int N, double & val, int & idx );
double vectorMax( vector<double> & v );
double vectorMin( vector<double> & v );

// ---- random numbers: host-supplied stream ----
double timeRand();
double hardRand();
// stream control (shim-only API, used by the oracle drivers)
void shimStreamSetArray( const double * u, size_t n );           // explicit stream (not copied)
void shimStreamSetCounter( unsigned long long seed, double scale ); // counter-based stream
size_t shimStreamPosition();
void shimStreamSeek( size_t pos );

// ---- process info / printing ----
int getProcID();

template <typename T> void print1DVector( vector<T> & v )
{
	cout << "[";
	for( size_t i = 0; i < v.size(); i++ ){ cout << setprecision(17) << v[i]; if( i + 1 < v.size() ) cout << ", "; }
	cout << "]" << endl;
}
template <typename T> void print1DVector( const vector<T> & v )
{
	vector<T> c(v); print1DVector(c);
}
void print2DVector( vector<vector<double> > & A );
void print1DArrayLine( double * v, int N, int prec, string name );

#endif
