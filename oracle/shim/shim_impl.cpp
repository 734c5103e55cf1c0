// TEST INFRASTRUCTURE (oracle shim) -- never linked into the product library.
// Implementation of oracle/shim/UtilityFunctions/utilityFunctions.hpp and oracle/shim/mpi.h.
// See those headers for what is pinned by the reference (nothing here is: the reference's utility
// library is absent) and what is our stated convention.

#include "UtilityFunctions/utilityFunctions.hpp"
#include "mpi.h"

#include <pthread.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>
#include <cstdio>
#include <cstring>

// ------------------------------------------------------------------------------------------------
// dense helpers
// ------------------------------------------------------------------------------------------------
double vector2Norm( vector<double> & v )
{
	double s = 0;
	for( size_t i = 0; i < v.size(); i++ ) s = s + v[i]*v[i];
	return sqrt(s);
}

double dotProd( vector<double> & a, vector<double> & b )
{
	double s = 0;
	for( size_t i = 0; i < a.size(); i++ ) s = s + a[i]*b[i];
	return s;
}

void matrixTranspose( vector<vector<double> > & A, vector<vector<double> > & AT )
{
	for( size_t i = 0; i < A.size(); i++ )
		for( size_t j = 0; j < A[i].size(); j++ )
			AT[j][i] = A[i][j];
}

void matrixMultiply( vector<vector<double> > & A, vector<vector<double> > & B, vector<vector<double> > & C )
{
	size_t M = A.size(), K = B.size(), N = B[0].size();
	for( size_t i = 0; i < M; i++ )
		for( size_t j = 0; j < N; j++ )
		{
			double s = 0;
			for( size_t k = 0; k < K; k++ ) s = s + A[i][k]*B[k][j];
			C[i][j] = s;
		}
}

void matrixVectorMultiply( vector<vector<double> > & A, vector<double> & x, vector<double> & y )
{
	for( size_t i = 0; i < A.size(); i++ )
	{
		double s = 0;
		for( size_t k = 0; k < x.size(); k++ ) s = s + A[i][k]*x[k];
		y[i] = s;
	}
}

// Doolittle LU with partial pivoting on a copy of A (A, b are left untouched).
static bool luFactor( vector<vector<double> > & LU, vector<int> & piv )
{
	int n = LU.size();
	for( int i = 0; i < n; i++ ) piv[i] = i;
	for( int k = 0; k < n; k++ )
	{
		int p = k; double best = fabs(LU[k][k]);
		for( int i = k+1; i < n; i++ ) if( fabs(LU[i][k]) > best ){ best = fabs(LU[i][k]); p = i; }
		if( p != k ){ LU[p].swap(LU[k]); int t = piv[p]; piv[p] = piv[k]; piv[k] = t; }
		for( int i = k+1; i < n; i++ )
		{
			LU[i][k] = LU[i][k]/LU[k][k];
			for( int j = k+1; j < n; j++ ) LU[i][j] = LU[i][j] - LU[i][k]*LU[k][j];
		}
	}
	return true;
}

static void luBackSub( vector<vector<double> > & LU, vector<int> & piv, vector<double> & b, vector<double> & x )
{
	int n = LU.size();
	vector<double> y(n);
	for( int i = 0; i < n; i++ )
	{
		double s = b[piv[i]];
		for( int k = 0; k < i; k++ ) s = s - LU[i][k]*y[k];
		y[i] = s;
	}
	for( int i = n-1; i >= 0; i-- )
	{
		double s = y[i];
		for( int k = i+1; k < n; k++ ) s = s - LU[i][k]*x[k];
		x[i] = s/LU[i][i];
	}
}

void luSolve( vector<vector<double> > & A, vector<double> & b, vector<double> & x )
{
	vector<vector<double> > LU(A);
	vector<int> piv(A.size());
	luFactor(LU, piv);
	luBackSub(LU, piv, b, x);
}

void matrixInverse( vector<vector<double> > & A, vector<vector<double> > & Ainv )
{
	int n = A.size();
	vector<vector<double> > LU(A);
	vector<int> piv(n);
	luFactor(LU, piv);
	vector<double> e(n), col(n);
	for( int j = 0; j < n; j++ )
	{
		for( int i = 0; i < n; i++ ) e[i] = (i == j) ? 1.0 : 0.0;
		luBackSub(LU, piv, e, col);
		for( int i = 0; i < n; i++ ) Ainv[i][j] = col[i];
	}
}

void setIdentity( vector<vector<double> > & A )
{
	for( size_t i = 0; i < A.size(); i++ )
		for( size_t j = 0; j < A[i].size(); j++ )
			A[i][j] = (i == j) ? 1.0 : 0.0;
}

// ------------------------------------------------------------------------------------------------
// scalar helpers
// ------------------------------------------------------------------------------------------------
int mod( int a, int b ){ int r = a % b; if( r < 0 ) r += b; return r; }

double sign( double x ){ if( x > 0 ) return 1.0; if( x < 0 ) return -1.0; return 0.0; }

void linspace( double a, double b, int N, vector<double> & out )
{
	out.resize( N > 0 ? N : 0 );
	if( N <= 0 ) return;
	if( N == 1 ){ out[0] = a; return; }
	double h = (b - a)/(N - 1);
	for( int i = 0; i < N; i++ ) out[i] = a + i*h;
	out[N-1] = b;
}

void vectorMin( double * v, int N, double & val, int & idx )
{
	val = v[0]; idx = 0;
	for( int i = 1; i < N; i++ ) if( v[i] < val ){ val = v[i]; idx = i; }
}
void vectorMax( double * v, int N, double & val, int & idx )
{
	val = v[0]; idx = 0;
	for( int i = 1; i < N; i++ ) if( v[i] > val ){ val = v[i]; idx = i; }
}
void vectorMin( vector<double> & v, int N, double & val, int & idx ){ vectorMin( v.data(), N, val, idx ); }
void vectorMax( vector<double> & v, int N, double & val, int & idx ){ vectorMax( v.data(), N, val, idx ); }
double vectorMax( vector<double> & v ){ double val; int idx; vectorMax( v.data(), (int) v.size(), val, idx ); return val; }
double vectorMin( vector<double> & v ){ double val; int idx; vectorMin( v.data(), (int) v.size(), val, idx ); return val; }

// ------------------------------------------------------------------------------------------------
// random stream
// ------------------------------------------------------------------------------------------------
static const double * gStreamArray = 0;
static size_t gStreamArrayLen = 0;
static unsigned long long gStreamSeed = 12345ULL;
static double gStreamScale = 1.0;
static size_t gStreamPos = 0;

static inline double splitmixUniform( unsigned long long seed, unsigned long long k, double scale )
{
	// Same generator as include/pnol_b200.h: pnol_stream_uniform(). Restated here on purpose.
	unsigned long long z = seed + (k + 1ULL)*0x9E3779B97F4A7C15ULL;
	z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27))*0x94D049BB133111EBULL;
	z = z ^ (z >> 31);
	return ((double)(z >> 11)*(1.0/9007199254740992.0))*scale;
}

void shimStreamSetArray( const double * u, size_t n ){ gStreamArray = u; gStreamArrayLen = n; gStreamPos = 0; }
void shimStreamSetCounter( unsigned long long seed, double scale ){ gStreamArray = 0; gStreamSeed = seed; gStreamScale = scale; gStreamPos = 0; }
size_t shimStreamPosition(){ return gStreamPos; }
void shimStreamSeek( size_t pos ){ gStreamPos = pos; }

double timeRand()
{
	if( gStreamArray )
	{
		if( gStreamPos >= gStreamArrayLen ){ fprintf(stderr, "oracle shim: random stream exhausted at %zu\n", gStreamPos); abort(); }
		return gStreamArray[gStreamPos++];
	}
	return splitmixUniform( gStreamSeed, gStreamPos++, gStreamScale );
}
double hardRand(){ return timeRand(); }

// ------------------------------------------------------------------------------------------------
// printing
// ------------------------------------------------------------------------------------------------
void print2DVector( vector<vector<double> > & A )
{
	for( size_t i = 0; i < A.size(); i++ ) print1DVector( A[i] );
}
void print1DArrayLine( double * v, int N, int prec, string name )
{
	cout << name << " = [";
	for( int i = 0; i < N; i++ ){ cout << setprecision(prec) << v[i]; if( i + 1 < N ) cout << ", "; }
	cout << "]" << endl;
}

// ------------------------------------------------------------------------------------------------
// mini-MPI (fork + shared mapping)
// ------------------------------------------------------------------------------------------------
namespace {

struct ShimWorld
{
	pthread_barrier_t barrier;
	int nprocs;
};

const size_t kSlotBytes = size_t(8) << 20;  // per-rank staging slot; larger messages are chunked

ShimWorld * gWorld = 0;
char * gSlots = 0;
int gRank = 0;
int gNprocs = 1;
pid_t gChildren[64];

inline void shimBarrier(){ if( gNprocs > 1 ) pthread_barrier_wait( &gWorld->barrier ); }

size_t dtSize( MPI_Datatype dt ){ return dt == MPI_DOUBLE ? sizeof(double) : sizeof(int); }

// recv[i] = sum over ranks (rank order) of slot_r[i]; writeResult selects who stores.
void reduceChunked( const void * sendbuf, void * recvbuf, int count, MPI_Datatype dt, bool writeResult )
{
	size_t es = dtSize(dt);
	if( gNprocs == 1 ){ if( writeResult && recvbuf != sendbuf ) memmove( recvbuf, sendbuf, es*count ); return; }
	size_t perChunk = kSlotBytes/es;
	for( size_t off = 0; off < (size_t) count; off += perChunk )
	{
		size_t nn = ((size_t) count - off < perChunk) ? (size_t) count - off : perChunk;
		memcpy( gSlots + gRank*kSlotBytes, (const char *) sendbuf + off*es, nn*es );
		shimBarrier();
		if( writeResult )
		{
			if( dt == MPI_DOUBLE )
			{
				double * out = (double *) recvbuf + off;
				for( size_t i = 0; i < nn; i++ )
				{
					double s = ((double *) (gSlots))[i];
					for( int r = 1; r < gNprocs; r++ ) s = s + ((double *) (gSlots + r*kSlotBytes))[i];
					out[i] = s;
				}
			}
			else
			{
				int * out = (int *) recvbuf + off;
				for( size_t i = 0; i < nn; i++ )
				{
					int s = 0;
					for( int r = 0; r < gNprocs; r++ ) s += ((int *) (gSlots + r*kSlotBytes))[i];
					out[i] = s;
				}
			}
		}
		shimBarrier();
	}
}

} // namespace

int MPI_Init( int *, char *** )
{
	const char * env = getenv("PNOL_SHIM_NPROCS");
	int P = env ? atoi(env) : 1;
	if( P < 1 ) P = 1;
	if( P > 64 ) P = 64;
	gNprocs = P; gRank = 0;
	if( P == 1 ) return MPI_SUCCESS;

	size_t bytes = sizeof(ShimWorld) + 4096 + kSlotBytes*P;
	char * base = (char *) mmap( 0, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0 );
	if( base == (char *) MAP_FAILED ){ perror("oracle shim mmap"); abort(); }
	gWorld = (ShimWorld *) base;
	gSlots = base + 4096;
	pthread_barrierattr_t attr;
	pthread_barrierattr_init( &attr );
	pthread_barrierattr_setpshared( &attr, PTHREAD_PROCESS_SHARED );
	pthread_barrier_init( &gWorld->barrier, &attr, P );
	gWorld->nprocs = P;
	fflush(stdout); fflush(stderr);
	for( int r = 1; r < P; r++ )
	{
		pid_t pid = fork();
		if( pid < 0 ){ perror("oracle shim fork"); abort(); }
		if( pid == 0 ){ gRank = r; return MPI_SUCCESS; }
		gChildren[r] = pid;
	}
	return MPI_SUCCESS;
}

int MPI_Finalize()
{
	fflush(stdout); fflush(stderr);
	if( gNprocs > 1 )
	{
		shimBarrier();
		if( gRank != 0 ) _exit(0);
		for( int r = 1; r < gNprocs; r++ ){ int st; waitpid( gChildren[r], &st, 0 ); }
	}
	return MPI_SUCCESS;
}

int MPI_Comm_size( MPI_Comm, int * size ){ *size = gNprocs; return MPI_SUCCESS; }
int MPI_Comm_rank( MPI_Comm, int * rank ){ *rank = gRank; return MPI_SUCCESS; }
int MPI_Barrier( MPI_Comm ){ shimBarrier(); return MPI_SUCCESS; }

int MPI_Allreduce( const void * sendbuf, void * recvbuf, int count, MPI_Datatype dt, MPI_Op, MPI_Comm )
{
	reduceChunked( sendbuf, recvbuf, count, dt, true );
	return MPI_SUCCESS;
}

int MPI_Reduce( const void * sendbuf, void * recvbuf, int count, MPI_Datatype dt, MPI_Op, int root, MPI_Comm )
{
	reduceChunked( sendbuf, recvbuf, count, dt, gRank == root );
	return MPI_SUCCESS;
}

int MPI_Bcast( void * buf, int count, MPI_Datatype dt, int root, MPI_Comm )
{
	if( gNprocs == 1 ) return MPI_SUCCESS;
	size_t es = dtSize(dt);
	size_t perChunk = kSlotBytes/es;
	for( size_t off = 0; off < (size_t) count; off += perChunk )
	{
		size_t nn = ((size_t) count - off < perChunk) ? (size_t) count - off : perChunk;
		if( gRank == root ) memcpy( gSlots, (char *) buf + off*es, nn*es );
		shimBarrier();
		if( gRank != root ) memcpy( (char *) buf + off*es, gSlots, nn*es );
		shimBarrier();
	}
	return MPI_SUCCESS;
}

double MPI_Wtime()
{
	struct timespec ts; clock_gettime( CLOCK_MONOTONIC, &ts );
	return ts.tv_sec + 1e-9*ts.tv_nsec;
}

int getProcID(){ return gRank; }
