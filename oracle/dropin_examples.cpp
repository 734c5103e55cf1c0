// dropin_examples.cpp -- TEST INFRASTRUCTURE (not product code): the reference's OWN example drivers
// (/root/reference/Source/Examples.cpp, compiled from where it lies, unmodified) built against THIS repo's plugin headers
// (include/pnol/*.hpp) and linked to libpnol_b200_host.so / libpnol_b200.so -- the drop-in claim of INTEGRATION.md (A) as a
// program: same user source, other include path and link line, every objective evaluation and dense update on the B200.
//
// How: the repo's headers are included first; then the include guards of the reference's headers are defined, so that the
// `#include "X.hpp"` lines of Examples.hpp (which resolve next to Examples.cpp, i.e. to the reference's own headers) are no-ops.
// Nothing of the reference is copied; the binary goes to oracle/_ref/ (git-ignored) like the verbatim build.
//
//   oracle/_ref/pnol_examples_dropin <driver> [pool_width]      driver = testBFGS, testBFGSBnd_MPI, testLMExpMPI, ...
#include <cstring>
#include <iostream>
#include <vector>
#include <cmath>
#include <mpi.h>            // include/pnol/nompi/mpi.h

using namespace std;

#include "SimplexSearch.hpp"
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

// the reference's guards (Source/*.hpp:9)
#define SIMPLEXSEARCH_HPP_
#define PNOL_ALGORITHM_HPP_
#define PNOL_OBJECTIVE_HPP_
#define EXAMPLEOBJECTIVES_HPP_
#define LEVENBERGMARQUARDT_HPP_
#define LEVENBERGMARQUARDTMPI_HPP_
#define GENETICALGORITHM_HPP_
#define GENETICALGORITHMMPI_HPP_
#define BFGS_WITH_LINESEARCH_HPP_
#define BFGS_WITH_LINESEARCH_MPI_HPP_
#define BFGS_WITH_BND_LINESEARCH_MPI_HPP_
#define BFGS_BND_LINESEARCH_HPP_
#define BFGS_BND_LINESEARCH_MPI_SW_HPP_
#define BOX_BOUNDARY_FUNCTIONS_HPP_

#include "Examples.cpp"     // found through -I$(REFSRC)

int main( int argc, char ** argv )
{
	if( argc < 2 ){ cerr << "usage: pnol_examples_dropin <driver> [pool_width] [seed]" << endl; return 2; }
	string d = argv[1];
	try
	{
		pnol::Runtime & rt = pnol::Runtime::instance();
		if( argc > 2 ) rt.setPoolWidth( atoi( argv[2] ) );
		pnol_stream_desc s;
		memset( &s, 0, sizeof s );
		s.seed = argc > 3 ? strtoull( argv[3], nullptr, 10 ) : 12345ULL;
		s.scale = 1.0 - 1.0/1048576.0;
		rt.setRandomStream( s );
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
		else { cerr << "unknown driver " << d << endl; return 2; }
	}
	catch( const std::exception & e )
	{
		cerr << "pnol_examples_dropin: " << e.what() << endl;
		return 1;
	}
	return 0;
}
