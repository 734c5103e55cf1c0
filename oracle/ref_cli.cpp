// TEST INFRASTRUCTURE -- command-line driver around the VERBATIM reference (compiled from /root/reference/Source
// by oracle/Makefile into oracle/_ref/). It exists to (a) pin oracle/pnol_oracle.cpp, (b) generate the fixtures under
// tests/golden/ (tests/golden/make_golden.py), (c) serve as the CPU baseline of bench.py (--impl reference).
// Nothing here is product code and no reference source is copied: the reference headers are #included from
// where they lie.
//
//   pnol_ref_cli <command> key=value ...      arrays are raw little-endian float64 files; outputs go to
//                                             <out>.<name>.f64 (written by rank 0 only)
//   PNOL_SHIM_NPROCS=P selects the number of mini-MPI ranks (oracle/shim/mpi.h).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <ctime>
#include <cmath>
#include <iostream>
#include <iomanip>
#include <mpi.h>

using namespace std;

#include "UtilityFunctions/utilityFunctions.hpp"
#include "PNOL_Algorithm.hpp"
#include "PNOL_Objective.hpp"
#include "ExampleObjectives.hpp"
#include "LevenbergMarquardt.hpp"
#include "LevenbergMarquardtMPI.hpp"
#include "GeneticAlgorithm.hpp"
#include "GeneticAlgorithmMPI.hpp"
#include "BFGS_with_linesearch.hpp"
#include "BFGS_with_linesearch_MPI.hpp"
#include "BFGS_with_bnd_linesearch_MPI.hpp"
#include "BFGS_bnd_linesearch.hpp"
#include "BFGS_bnd_linesearch_MPI_SW.hpp"
#include "Box_boundary_functions.hpp"
#include "SimplexSearch.hpp"

#include "Examples.hpp"
#include "oracle_objectives.h"

static std::streambuf * gKeepCout = 0;

double computeAlphaBnd( vector <double> & X, vector <double> & Xlb, vector <double> & Xub, vector <double> & p );

// ---- our synthetic objectives as reference plugin classes (restated arithmetic of oracle_objectives.h) ----
class OracleScalarObjective : public Objective {
  public:
	OFunctor f;
	vector<double> scalars; vector<long long> ints; vector<vector<double> > colData; vector<const double *> colPtr;
	double objEval( vector<double> & X ){ return o_scalar( f, X.data(), (int) X.size() ); }
	void finish( int kind, long long m )
	{
		colPtr.resize( colData.size() );
		for( size_t c = 0; c < colData.size(); c++ ) colPtr[c] = colData[c].data();
		scalars.resize(8, 0.0); ints.resize(4, 0);
		f.kind = kind; f.scalars = scalars.data(); f.ints = ints.data(); f.cols = colPtr.data(); f.m = m;
	}
};
class OracleMultiObjective : public MultiObjective {
  public:
	OFunctor f;
	vector<double> scalars; vector<long long> ints; vector<vector<double> > colData; vector<const double *> colPtr;
	void objEval( vector<double> & X, vector<double> & F ){ o_residual( f, X.data(), (int) X.size(), F.data() ); }
	void finish( int kind, long long m )
	{
		colPtr.resize( colData.size() );
		for( size_t c = 0; c < colData.size(); c++ ) colPtr[c] = colData[c].data();
		scalars.resize(8, 0.0); ints.resize(4, 0);
		f.kind = kind; f.scalars = scalars.data(); f.ints = ints.data(); f.cols = colPtr.data(); f.m = m;
	}
};

// ---- helpers ----
static map<string, string> gArgs;
static int gRank = 0;

static string arg( const string & k, const string & dflt = "" )
{
	map<string, string>::iterator it = gArgs.find(k);
	return it == gArgs.end() ? dflt : it->second;
}
static double argd( const string & k, double dflt ){ string v = arg(k); return v.empty() ? dflt : atof(v.c_str()); }
static long long argi( const string & k, long long dflt ){ string v = arg(k); return v.empty() ? dflt : atoll(v.c_str()); }

static vector<double> readf64( const string & path )
{
	vector<double> v;
	FILE * fp = fopen( path.c_str(), "rb" );
	if( !fp ){ fprintf(stderr, "ref_cli: cannot open %s\n", path.c_str()); exit(2); }
	fseek( fp, 0, SEEK_END ); long sz = ftell(fp); fseek( fp, 0, SEEK_SET );
	v.resize( sz/8 );
	if( sz > 0 && fread( v.data(), 8, v.size(), fp ) != v.size() ){ fprintf(stderr, "ref_cli: short read %s\n", path.c_str()); exit(2); }
	fclose(fp);
	return v;
}
static void writef64( const string & name, const double * p, size_t n )
{
	if( gRank != 0 ) return;
	string path = arg("out", "ref_out") + "." + name + ".f64";
	FILE * fp = fopen( path.c_str(), "wb" );
	if( !fp ){ fprintf(stderr, "ref_cli: cannot write %s\n", path.c_str()); exit(2); }
	fwrite( p, 8, n, fp );
	fclose(fp);
}
static void writef64( const string & name, const vector<double> & v ){ writef64( name, v.data(), v.size() ); }
static void writeScalar( const string & name, double v ){ writef64( name, &v, 1 ); }
static vector<double> argvec( const string & k, int n, double dflt )
{
	string v = arg(k);
	if( v.empty() ) return vector<double>( n, dflt );
	if( v.find('/') != string::npos || v.find(".f64") != string::npos ) return readf64(v);
	return vector<double>( n, atof(v.c_str()) );
}
static void flatten( vector<vector<double> > & A, vector<double> & out )
{
	out.clear();
	for( size_t i = 0; i < A.size(); i++ ) out.insert( out.end(), A[i].begin(), A[i].end() );
}

// scalar objective factory: reference fixtures by their own classes, ours by OracleScalarObjective
static Objective * makeScalar( const string & spec )
{
	if( spec == "rosenbrock" ) return new RosenbrockObject();
	if( spec == "booth" ) return new BoothFunction();
	if( spec == "goldstein" ) return new GoldsteinFunction();
	if( spec.compare(0, 6, "power:") == 0 ){ PowerObject * o = new PowerObject(); o->setPower( atoi(spec.c_str() + 6) ); return o; }
	if( spec == "expsingle_ref" ) return new ExpCurveObjectiveSingle();
	OracleScalarObjective * o = new OracleScalarObjective();
	if( spec.compare(0, 10, "powerprod:") == 0 ){ o->finish(2, 0); o->ints[0] = atoi(spec.c_str() + 10); return o; }
	if( spec == "rastrigin" ){ o->finish(5, 0); return o; }
	if( spec == "expsingle" )
	{
		o->colData.resize(2);
		linspace( 0, 5, 100, o->colData[0] );
		o->colData[1].resize(100);
		for( int k = 0; k < 100; k++ ) o->colData[1][k] = 10.2*exp( 0.4*o->colData[0][k] ) + 0.1;   // ExampleObjectives.hpp:311
		o->finish(6, 100);
		return o;
	}
	fprintf(stderr, "ref_cli: unknown scalar objective %s\n", spec.c_str()); exit(2);
}

static MultiObjective * makeMulti( const string & spec, long long & m )
{
	if( spec == "expcurve_ref" ){ ExpCurveObjective * o = new ExpCurveObjective(); m = o->getDataSize(); return o; }
	if( spec == "cubic" ){ CubicObjective * o = new CubicObjective(); m = o->getDataSize(); return o; }
	OracleMultiObjective * o = new OracleMultiObjective();
	if( spec == "expcurve" )
	{
		o->colData.resize(2);
		linspace( 0, 5, 100, o->colData[0] );
		o->colData[1].resize(100);
		for( int k = 0; k < 100; k++ ) o->colData[1][k] = 10.2*exp( 0.4*o->colData[0][k] ) + 0.1;   // ExampleObjectives.hpp:145
		m = 100; o->finish(101, m);
		return o;
	}
	if( spec == "lorentz" )
	{
		o->colData.resize(2);
		o->colData[0] = readf64( arg("t") );
		o->colData[1] = readf64( arg("y") );
		m = (long long) o->colData[0].size();
		o->finish(103, m);
		o->scalars[0] = argd("w", 4.0);
		return o;
	}
	fprintf(stderr, "ref_cli: unknown residual objective %s\n", spec.c_str()); exit(2);
}

// ---- commands ----
static int cmdFdGrad()
{
	Objective * obj = makeScalar( arg("obj") );
	vector<double> X = readf64( arg("x") );
	int n = (int) X.size();
	vector<double> dX = argvec( "dx", n, 1e-6 );
	vector<double> g(n, 0), gm(n, 0);
	obj->gradientApproximation( X, dX, g );
	obj->gradientApproximationMPI( X, dX, gm );
	writef64( "g", g ); writef64( "g_mpi", gm );
	writeScalar( "f", obj->objEval(X) );
	return 0;
}

static int cmdRecur()
{
	Objective * obj = makeScalar( arg("obj") );
	vector<double> Xr = readf64( arg("x") );
	vector<double> constX = readf64( arg("constx") );
	vector<double> indf = readf64( arg("ind") );
	vector<bool> ind( indf.size() );
	for( size_t i = 0; i < indf.size(); i++ ) ind[i] = indf[i] != 0;
	int nr = (int) Xr.size();
	vector<double> dX = argvec( "dx", nr, 1e-6 );
	vector<double> g(nr, 0), gm(nr, 0);
	obj->gradientApproximationRecur( Xr, dX, g, constX, ind );
	obj->gradientApproximationMPIRecur( Xr, dX, gm, constX, ind );
	writef64( "g", g ); writef64( "g_mpi", gm );
	writeScalar( "f", obj->objEvalRecur( Xr, constX, ind ) );
	return 0;
}

static int cmdHessian()
{
	Objective * obj = makeScalar( arg("obj") );
	vector<double> X = readf64( arg("x") );
	int n = (int) X.size();
	vector<double> dX = argvec( "dx", n, 1e-3 );
	vector<vector<double> > B( n, vector<double>(n, 0) );
	obj->hessianApproximation( X, dX, B );
	vector<double> flat; flatten( B, flat );
	writef64( "B", flat );
	return 0;
}

static int cmdEval()
{
	Objective * obj = makeScalar( arg("obj") );
	vector<double> pts = readf64( arg("pts") );
	int n = (int) argi("n", 1);
	size_t B = pts.size()/n;
	vector<double> f(B), X(n);
	for( size_t b = 0; b < B; b++ ){ for( int j = 0; j < n; j++ ) X[j] = pts[b*n + j]; f[b] = obj->objEval(X); }
	writef64( "f", f );
	return 0;
}

static int cmdFdJac()
{
	long long m = 0;
	MultiObjective * obj = makeMulti( arg("obj"), m );
	vector<double> X = readf64( arg("x") );
	int n = (int) X.size();
	vector<double> dX = argvec( "dx", n, 1e-7 );
	vector<vector<double> > J( m, vector<double>(n, 0) ), Jm( m, vector<double>(n, 0) );
	vector<double> F( m, 0 );
	obj->objEval( X, F );
	obj->gradientApproximation( X, dX, J );
	obj->gradientApproximationMPI( X, dX, Jm );
	vector<double> flat;
	flatten( J, flat ); writef64( "J", flat );
	flatten( Jm, flat ); writef64( "J_mpi", flat );
	writef64( "F", F );
	return 0;
}

static int cmdLM()
{
	long long m = 0;
	MultiObjective * obj = makeMulti( arg("obj"), m );
	vector<double> X = readf64( arg("x") );
	vector<double> F0( m, 0 ), F( m, 0 );
	double lambda0 = argd("lambda0", 0.001), factor = argd("factor", 10), dXGrad = argd("dxgrad", 1e-6), xMinDiff = argd("xmindiff", 1e-6);
	double maxIter = argd("maxiter", 100);
	if( argi("serial", 0) )
	{
		LevMarq lm; lm.setObjPtr( *obj );
		lm.setParams( lambda0, factor, dXGrad, maxIter, xMinDiff, -1 );
		lm.findMin( X, F0, F );
	}
	else
	{
		LevMarqMPI lm; lm.setObjPtr( *obj );
		lm.setParams( lambda0, factor, dXGrad, maxIter, xMinDiff, -1 );
		lm.findMin( X, F0, F );
	}
	writef64( "X", X ); writef64( "F0", F0 ); writef64( "F", F );
	return 0;
}

static int cmdUpdHinv()
{
	vector<double> Dflat = readf64( arg("D") ), g = readf64( arg("g") ), s = readf64( arg("s") );
	int n = (int) g.size();
	vector<vector<double> > D( n, vector<double>(n) );
	for( int i = 0; i < n; i++ ) for( int j = 0; j < n; j++ ) D[i][j] = Dflat[(size_t) i*n + j];
	updateHessianInv( D, g, s );
	vector<double> flat; flatten( D, flat );
	writef64( "D", flat );
	return 0;
}

static int cmdBfgs( const string & variant )
{
	Objective * obj = makeScalar( arg("obj") );
	vector<double> X = readf64( arg("x") );
	int n = (int) X.size();
	double f0 = 0, fOpt = 0;
	double c1 = argd("c1", 1e-4), c2 = argd("c2", 0.9), dalpha = argd("dalpha", 1e-6), alphaGuess = argd("alphaguess", 1);
	int maxIterLS = (int) argi("maxiterls", 1000);
	double dXGrad = argd("dxgrad", 1e-7), dXHess = argd("dxhess", 1e-3), maxIter = argd("maxiter", 100);
	double xMinDiff = argd("xmindiff", 1e-5), minGrad = argd("mingrad", 1e-5);
	bool initHess = argi("inithess", 0) != 0;
	if( variant == "bfgs" )
	{
		BFGS b; b.setObjPtr( *obj );
		b.setParams( c1, c2, dalpha, alphaGuess, maxIterLS, dXGrad, dXHess, maxIter, xMinDiff, minGrad, initHess, 0 );
		b.findMin( X, f0, fOpt );
	}
	else if( variant == "bfgs_mpi" )
	{
		BFGS_MPI b; b.setObjPtr( *obj );
		b.setParams( c1, c2, argd("maxalphamult", 4), alphaGuess, maxIterLS, dXGrad, dXHess, maxIter, xMinDiff, minGrad, initHess, 0 );
		b.findMin( X, f0, fOpt );
	}
	else
	{
		vector<double> Xlb = argvec( "xlb", n, -5 ), Xub = argvec( "xub", n, 5 );
		double alphaTol = argd("alphatol", 1e-10), alphaMult = argd("alphamult", 2);
		if( variant == "bfgsbnd_mpi" )
		{
			// Source/BFGS_with_bnd_linsearch_MPI.cpp:14-80; pool width = mini-MPI rank count. Its progress prints go to stdout.
			BFGSBnd_MPI b; b.setObjPtr( *obj );
			b.setParams( c1, c2, argd("alphamin", 1e-16), argd("maxalphamult", 4), alphaGuess, maxIterLS, dXGrad, dXHess, maxIter, xMinDiff, minGrad,
					argd("fsteptol", 1e-5), initHess, argi("verbose", 0) != 0 );
			b.findMinBnd( X, Xlb, Xub, f0, fOpt );
		}
		else if( variant == "bfgs_bnd" )
		{
			BFGS_Bnd b; b.setObjPtr( *obj );
			b.setParams( c1, c2, dalpha, alphaGuess, alphaTol, alphaMult, maxIterLS, argd("bndtol", 1e-5), dXGrad, dXHess, maxIter, xMinDiff, minGrad, initHess, 0 );
			b.findMinBnd( X, Xlb, Xub, f0, fOpt );
		}
		else
		{
			BFGS_Bnd_MPI_SW b; b.setObjPtr( *obj );
			b.setParams( c1, c2, dalpha, alphaGuess, alphaTol, alphaMult, maxIterLS, argd("bndtol", 1e-5), dXGrad, dXHess, maxIter, xMinDiff, minGrad, initHess, 0 );
			b.findMinBnd( X, Xlb, Xub, f0, fOpt );
		}
	}
	writef64( "X", X ); writeScalar( "f0", f0 ); writeScalar( "fOpt", fOpt );
	return 0;
}

// the reference's own example drivers (Source/Examples.cpp), for side-by-side runs with oracle/_ref/pnol_examples_dropin
static int cmdExample()
{
	string d = arg("name");
	int rank = 0;
	MPI_Comm_rank( MPI_COMM_WORLD, &rank );
	if( rank == 0 ) cout.rdbuf( gKeepCout );          // the drivers print from every rank; keep the root's
	shimStreamSetCounter( (unsigned long long) argi("seed", 12345), argd("scale", 1.0 - 1.0/1048576.0) );
	if( d == "testBFGSBndMPISW" ) testBFGSBndMPISW();
	else if( d == "testBFGSBnd" ) testBFGSBnd();
	else if( d == "testBFGSBnd_MPI" ) testBFGSBnd_MPI();
	else if( d == "testLMExpMPI" ) testLMExpMPI();
	else if( d == "testBFGS_MPI" ) testBFGS_MPI();
	else if( d == "testBFGS_booth" ) testBFGS_booth();
	else if( d == "testBFGS" ) testBFGS();
	else if( d == "testHessian" ) testHessian();
	else if( d == "testGAParallel" ) testGAParallel();
	else if( d == "testGA" ) testGA();
	else if( d == "testLMExp" ) testLMExp();
	else if( d == "testLMCubicLinearCoef" ) testLMCubicLinearCoef();
	else if( d == "testSimplexSearch" ) testSimplexSearch();
	else if( d == "testCreateObject" ) testCreateObject();
	else if( d == "testGradientEvaluation" ) testGradientEvaluation();
	else if( d == "testGradientApproxMultMPI" ) testGradientApproxMultMPI();
	else if( d == "testGradientApproxMultMPIRecur" ) testGradientApproxMultMPIRecur();
	else return 2;
	return 0;
}

static void setStreamFromArgs()
{
	static vector<double> explicitStream;
	if( !arg("stream").empty() ){ explicitStream = readf64( arg("stream") ); shimStreamSetArray( explicitStream.data(), explicitStream.size() ); }
	else shimStreamSetCounter( (unsigned long long) argi("seed", 12345), argd("scale", 1.0) );
}

// SimplexSearch::findMin of the verbatim reference (Source/SimplexSearch.cpp:13-226); its srand(time(0)) is harmless, the start
// simplex comes from timeRand() = the host-supplied stream of the shim
static int cmdSimplex()
{
	Objective * obj = makeScalar( arg("obj") );
	vector<double> X = readf64( arg("x") );
	setStreamFromArgs();
	SimplexSearch s; s.setObjPtr( *obj );
	s.setSimplexParams( argd("alpha", 1.0), argd("gamma", 2.0), argd("rho", 0.5), argd("sigma", 0.5), (int) argi("maxiter", 10000),
			argd("initrandmax", 1.0), argd("xmindiff", 1e-7), false );
	double f0 = 0, fOpt = 0;
	s.findMin( X, f0, fOpt );
	writef64( "X", X ); writeScalar( "f0", f0 ); writeScalar( "fOpt", fOpt );
	writeScalar( "stream_pos", (double) shimStreamPosition() );
	return 0;
}

static int cmdGA()
{
	Objective * obj = makeScalar( arg("obj") );
	vector<double> X = readf64( arg("x") );
	int n = (int) X.size();
	vector<double> Xlb = argvec( "xlb", n, -5 ), Xub = argvec( "xub", n, 5 );
	setStreamFromArgs();
	double f0 = 0, fOpt = 0;
	int Npop = (int) argi("npop", 100), maxGen = (int) argi("maxgen", 10);
	if( argi("serial", 0) )
	{
		GeneticAlgorithm ga; ga.setObjPtr( *obj );
		ga.setGAParams( Npop, maxGen, argd("elitefrac", 0.1), argd("crossfrac", 0.3), argd("elitemutfrac", 0.2), argd("mutsize", 0.5),
				argd("elitemutsize", 0.01), 0.5, argd("nstatic", 50), false, false );
		ga.findMinBnd( X, Xlb, Xub, f0, fOpt );
	}
	else
	{
		GeneticAlgorithmMPI ga; ga.setObjPtr( *obj );
		ga.setGAParams( Npop, maxGen, argd("elitefrac", 0.1), argd("crossfrac", 0.3), argd("elitemutfrac", 0.2), argd("mutsize", 0.5),
				argd("elitemutsize", 0.01), 0.5, argd("nstatic", 50), false );
		ga.findMinBnd( X, Xlb, Xub, f0, fOpt );
	}
	writef64( "X", X ); writeScalar( "f0", f0 ); writeScalar( "fOpt", fOpt );
	writeScalar( "stream_pos", (double) shimStreamPosition() );
	return 0;
}

static int cmdGAStage( const string & which )
{
	int n = (int) argi("n", 1);
	vector<double> flat = readf64( arg("xpop") );
	size_t Npop = flat.size()/n;
	vector<vector<double> > Xpop( Npop, vector<double>(n) );
	for( size_t i = 0; i < Npop; i++ ) for( int j = 0; j < n; j++ ) Xpop[i][j] = flat[i*n + j];
	if( which == "popsort" )
	{
		vector<double> F = readf64( arg("F") );
		popSort( Xpop, F );
		writef64( "F", F );
	}
	else
	{
		vector<double> Xlb = argvec( "xlb", n, -5 ), Xub = argvec( "xub", n, 5 );
		vector<bool> ind( Npop, false );
		setStreamFromArgs();
		if( which == "checkbounds" ) checkPopulationBoundsAndReplace( Xpop, Xlb, Xub, ind );
		else checkIndenticalChildAndReplace( Xpop, Xlb, Xub, ind );
		vector<double> indf( Npop );
		for( size_t i = 0; i < Npop; i++ ) indf[i] = ind[i] ? 1.0 : 0.0;
		writef64( "ind", indf );
		writeScalar( "stream_pos", (double) shimStreamPosition() );
	}
	flatten( Xpop, flat ); writef64( "xpop", flat );
	return 0;
}

static int cmdBox()
{
	vector<double> X = readf64( arg("x") );
	int n = (int) X.size();
	vector<double> Xlb = argvec( "xlb", n, -5 ), Xub = argvec( "xub", n, 5 );
	if( !arg("p").empty() )
	{
		vector<double> p = readf64( arg("p") );
		writeScalar( "alphabnd", computeAlphaBnd( X, Xlb, Xub, p ) );
	}
	// silence the reference's warning prints
	std::streambuf * old = cout.rdbuf(0);
	checkBoxBounds( X, Xlb, Xub );
	cout.rdbuf(old);
	writef64( "X", X );
	return 0;
}

// The alpha-pool evaluation of the pooled line searches through the reference's own class:
// BFGS_Bnd_MPI_SW::evaluateAlphaPoolAndDerivatives (Source/BFGS_bnd_linesearch_MPI_SW.cpp:599-699; phi and its forward-difference
// slope per entry, dealt over the ranks, 1e10 sentinel for NaN / inf) with an optional active set (constx / ind of the full length)
static int cmdAlphaPool()
{
	Objective * obj = makeScalar( arg("obj") );
	vector<double> X = readf64( arg("x") ), p = readf64( arg("p") ), alphaPool = readf64( arg("alpha") );
	int n = (int) X.size();
	vector<double> constantX( n, 0 );
	vector<bool> constantIndicator( n, false );
	if( !arg("constx").empty() )
	{
		constantX = readf64( arg("constx") );
		vector<double> ind = readf64( arg("ind") );
		constantIndicator.assign( constantX.size(), false );
		for( size_t i = 0; i < ind.size(); i++ ) constantIndicator[i] = ind[i] != 0;
	}
	vector<double> phiPool( alphaPool.size(), 0 ), dphiPool( alphaPool.size(), 0 );
	BFGS_Bnd_MPI_SW b; b.setObjPtr( *obj );
	b.setParams( 1e-4, 0.9, argd("dalpha", 1e-6), 1, 1e-20, 2, 50, 1e-5, 1e-6, 1e-3, 100, 1e-5, 1e-5, 0, 0 );
	b.evaluateAlphaPoolAndDerivatives( alphaPool, X, p, constantX, constantIndicator, phiPool, dphiPool );
	writef64( "phi", phiPool ); writef64( "dphi", dphiPool );
	return 0;
}

// CPU baseline: one LM iteration's hot path (Jacobian + normal equations + solve + trial residual) on the
// Lorentz-sum model through the reference's own code, timed with MPI_Wtime on rank 0.
static int cmdBenchLM()
{
	long long m = 0;
	MultiObjective * obj = makeMulti( "lorentz", m );
	vector<double> X = readf64( arg("x") );
	int n = (int) X.size();
	int steps = (int) argi("steps", 2), warmup = (int) argi("warmup", 1);
	LevMarqMPI lm; lm.setObjPtr( *obj );
	vector<double> F0( m, 0 ), F( m, 0 );
	// warm-up + timed run use findMin with maxIter = k; xMinDiff = 0 so that it never stops early
	vector<double> Xw(X);
	lm.setParams( argd("lambda0", 0.001), argd("factor", 10), argd("dxgrad", 1e-6), (double) warmup, 0.0, -1 );
	if( warmup > 0 ) lm.findMin( Xw, F0, F );
	MPI_Barrier( MPI_COMM_WORLD );
	double t0 = MPI_Wtime();
	vector<double> Xt(X);
	lm.setParams( argd("lambda0", 0.001), argd("factor", 10), argd("dxgrad", 1e-6), (double) steps, 0.0, -1 );
	lm.findMin( Xt, F0, F );
	MPI_Barrier( MPI_COMM_WORLD );
	double t1 = MPI_Wtime();
	int P; MPI_Comm_size( MPI_COMM_WORLD, &P );
	if( gRank == 0 )
		printf( "{\"bench\": \"lm\", \"m\": %lld, \"n\": %d, \"steps\": %d, \"ranks\": %d, \"seconds\": %.6f, \"s_per_iter\": %.6f}\n",
				m, n, steps, P, t1 - t0, (t1 - t0)/steps );
	writef64( "X", Xt );
	return 0;
}

// CPU baseline: GA fitness sweep through GeneticAlgorithmMPI::evaluatePopulationParallel
static int cmdBenchGAEval()
{
	Objective * obj = makeScalar( arg("obj", "rastrigin") );
	int n = (int) argi("n", 32), Npop = (int) argi("npop", 20000), reps = (int) argi("reps", 3);
	setStreamFromArgs();
	vector<vector<double> > Xpop( Npop, vector<double>(n) );
	for( int i = 0; i < Npop; i++ ) for( int j = 0; j < n; j++ ) Xpop[i][j] = -5.12 + 10.24*timeRand();
	vector<double> F( Npop, 0 ); vector<bool> ind( Npop, true );
	GeneticAlgorithmMPI ga; ga.setObjPtr( *obj );
	ga.setGAParams( Npop, 1, 0.1, 0.3, 0.2, 0.5, 0.01, 0.5, 50, false );
	MPI_Barrier( MPI_COMM_WORLD );
	double t0 = MPI_Wtime();
	for( int r = 0; r < reps; r++ ){ for( int i = 0; i < Npop; i++ ) ind[i] = true; ga.evaluatePopulationParallel( Xpop, F, ind ); }
	MPI_Barrier( MPI_COMM_WORLD );
	double t1 = MPI_Wtime();
	int P; MPI_Comm_size( MPI_COMM_WORLD, &P );
	if( gRank == 0 )
		printf( "{\"bench\": \"ga_eval\", \"npop\": %d, \"n\": %d, \"reps\": %d, \"ranks\": %d, \"seconds\": %.6f, \"evals_per_s\": %.1f}\n",
				Npop, n, reps, P, t1 - t0, (double) Npop*reps/(t1 - t0) );
	writef64( "F", F );
	return 0;
}

int main( int argc, char ** argv )
{
	if( argc < 2 ){ fprintf(stderr, "usage: pnol_ref_cli <command> key=value ...\n"); return 2; }
	string cmd = argv[1];
	for( int i = 2; i < argc; i++ )
	{
		string a = argv[i];
		size_t eq = a.find('=');
		if( eq == string::npos ) gArgs[a] = "1"; else gArgs[a.substr(0, eq)] = a.substr(eq + 1);
	}
	MPI_Init( &argc, &argv );
	MPI_Comm_rank( MPI_COMM_WORLD, &gRank );
	// the reference prints progress unconditionally in places; keep stdout of non-root ranks quiet
	std::streambuf * keep = cout.rdbuf();
	gKeepCout = keep;
	if( arg("quiet", "1") == "1" && cmd.compare(0, 5, "bench") != 0 ) cout.rdbuf(0);
	int rc = 0;
	if( cmd == "fdgrad" ) rc = cmdFdGrad();
	else if( cmd == "recur" ) rc = cmdRecur();
	else if( cmd == "hessian" ) rc = cmdHessian();
	else if( cmd == "eval" ) rc = cmdEval();
	else if( cmd == "fdjac" ) rc = cmdFdJac();
	else if( cmd == "lm" ) rc = cmdLM();
	else if( cmd == "updhinv" ) rc = cmdUpdHinv();
	else if( cmd == "bfgs" || cmd == "bfgs_mpi" || cmd == "bfgs_bnd" || cmd == "bfgs_bnd_sw" || cmd == "bfgsbnd_mpi" ) rc = cmdBfgs(cmd);
	else if( cmd == "ga" ) rc = cmdGA();
	else if( cmd == "simplex" ) rc = cmdSimplex();
	else if( cmd == "popsort" || cmd == "checkbounds" || cmd == "checkidentical" ) rc = cmdGAStage(cmd);
	else if( cmd == "box" ) rc = cmdBox();
	else if( cmd == "alphapool" ) rc = cmdAlphaPool();
	else if( cmd == "example" ) rc = cmdExample();
	else if( cmd == "bench_lm" ){ cout.rdbuf(0); rc = cmdBenchLM(); }
	else if( cmd == "bench_ga_eval" ){ cout.rdbuf(0); rc = cmdBenchGAEval(); }
	else { fprintf(stderr, "ref_cli: unknown command %s\n", cmd.c_str()); rc = 2; }
	cout.rdbuf(keep);
	MPI_Finalize();
	return rc;
}
