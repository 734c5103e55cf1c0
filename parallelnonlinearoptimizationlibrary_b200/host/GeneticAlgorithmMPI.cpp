/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
// GeneticAlgorithmMPI.cpp -- host face of the device GA (Source/GeneticAlgorithmMPI.cpp, Source/GeneticAlgorithm.cpp).
#include "pnol/GeneticAlgorithm.hpp"

#include <chrono>
#include <cmath>
#include <iostream>

namespace {

pnol_stream_desc currentStream( int Npop )
{
	// scale keeps round(u * Npop) < Npop, where the reference would index one past the end of its arrays (:138-140).
	return pnol::Runtime::instance().defaultStream( 1.0 - 1.0/(double) Npop );
}

void flatten( const vector<vector<double> > & A, vector<double> & flat )
{
	flat.clear();
	for( size_t i = 0; i < A.size(); i++ ) flat.insert( flat.end(), A[i].begin(), A[i].end() );
}
void unflatten( const vector<double> & flat, vector<vector<double> > & A )
{
	size_t k = 0;
	for( size_t i = 0; i < A.size(); i++ ) for( size_t j = 0; j < A[i].size(); j++ ) A[i][j] = flat[k++];
}

// stream position shared by the stand-alone host-vector stage functions below
uint64_t gStagePos = 0;

}

namespace pnol {

void gaFindMinBnd( Objective * objPtr, int Npop, int maxGenerations, double eliteFrac, double crossFrac, double eliteMutationFrac,
		double mutationSize, double eliteMutationSize, double NstaticGenerations, bool verbose,
		std::vector <double> & X, std::vector <double> & Xlb, std::vector <double> & Xub, double & f0, double & fOpt, GAReport & report )
{
	Runtime & rt = Runtime::instance();
	pnol_ctx * ctx = rt.ctx();
	pnol_functor * f = objPtr->deviceFunctor();
	if( !f ) throw Error( PNOL_ERR_NO_FUNCTOR, "GeneticAlgorithm: the objective has no device functor (no CPU fallback)" );
	int Nparam = (int) X.size();

	pnol_ga_params prm;
	prm.npop = Npop; prm.max_generations = maxGenerations; prm.elite_frac = eliteFrac; prm.cross_frac = crossFrac;
	prm.elite_mutation_frac = eliteMutationFrac; prm.mutation_size = mutationSize; prm.elite_mutation_size = eliteMutationSize;
	prm.n_static_generations = NstaticGenerations;
	pnol_stream_desc stream = currentStream( Npop );

	pnol_ga * ga = nullptr;
	// invalid fractions: the reference prints "666 GA fractions set incorrectly..." and calls exit(0)
	// (Source/GeneticAlgorithmMPI.cpp:37-44); here the status comes back as a pnol::Error
	rt.check( pnol_ga_create( ctx, f, &prm, Nparam, Xlb.data(), Xub.data(), &stream, &ga ) );
	try
	{
		rt.check( pnol_ga_init( ga, X.data(), &f0 ) );                         // (:55-81)
		pnol_ga_status st;
		rt.check( pnol_ga_status_get( ga, &st ) );
		objPtr->noteDeviceEvaluations( Npop );                                // initial population (:76)
		if( verbose )
			cout << "Computing genetic algorithm with population Nelite = " << st.n_elite << ", NeliteMut = " << st.n_elite_mut
			     << ", Ncross = " << st.n_cross << ", Nrand = " << st.n_rand << endl;
		while( !st.stopped && st.generation < maxGenerations )                // (:87)
		{
			rt.check( pnol_ga_generation( ga ) );
			rt.check( pnol_ga_status_get( ga, &st ) );
			objPtr->noteDeviceEvaluations( Npop - st.n_elite );               // elites keep their value (:108-118, :220)
			if( verbose ) cout << "At generation = " << st.generation << " minimum of f = " << st.f_best << endl;
		}
		// store result (:255-259): the best individual is row 0 of the sorted population
		{
			vector<double> pop( (size_t) Npop*Nparam ), Fall( Npop );
			rt.check( pnol_ga_get_population( ga, pop.data(), Fall.data() ) );
			for( int i = 0; i < Nparam; i++ ) X[i] = pop[i];
			fOpt = Fall[0];
		}
		report.generations = st.generation; report.stoppedStatic = st.stopped; report.streamPos = st.stream_pos;
		report.Nelite = st.n_elite; report.NeliteMut = st.n_elite_mut; report.Ncross = st.n_cross; report.Nrand = st.n_rand;
	}
	catch( ... )
	{
		pnol_ga_destroy( ga );
		throw;
	}
	pnol_ga_destroy( ga );
	if( verbose ) { cout << "Completed genetic algorithm. minimum of f = " << fOpt << " at params: "; print1DVector( X ); }
}

} // namespace pnol

static void evaluateRows( Objective * objPtr, vector<vector<double> > & Xpop, vector <double> & F, vector <bool> & evaluateIndicator )
{
	pnol::Runtime & rt = pnol::Runtime::instance();
	pnol_functor * f = objPtr->deviceFunctor();
	if( !f ) throw pnol::Error( PNOL_ERR_NO_FUNCTOR, "evaluatePopulation: the objective has no device functor (no CPU fallback)" );
	vector<double> flat;
	flatten( Xpop, flat );
	vector<unsigned char> ind( evaluateIndicator.size() );
	for( size_t i = 0; i < ind.size(); i++ ) ind[i] = evaluateIndicator[i] ? 1 : 0;
	int n = Xpop.empty() ? 0 : (int) Xpop[0].size();
	rt.check( pnol_eval_batch( rt.ctx(), f, flat.data(), (long long) Xpop.size(), n, n, ind.data(), F.data() ) );
	long long evaluated = 0;
	for( size_t i = 0; i < ind.size(); i++ ) evaluated += ind[i];
	objPtr->noteDeviceEvaluations( evaluated );
}

// Source/GeneticAlgorithmMPI.cpp:283-414
void GeneticAlgorithmMPI::evaluatePopulationParallel( vector<vector<double> > & Xpop, vector <double> & F, vector <bool> & evaluateIndicator )
{
	evaluateRows( objPtr, Xpop, F, evaluateIndicator );
}

// Source/GeneticAlgorithm.cpp:301-311
void GeneticAlgorithm::evaluatePopulation( vector<vector<double> > & Xpop, vector <double> & F, vector <bool> & evaluateIndicator )
{
	evaluateRows( objPtr, Xpop, F, evaluateIndicator );
}

// Source/GeneticAlgorithm.cpp:370-412
void popSort( vector<vector<double> > & Xpop, vector <double> & F )
{
	pnol::Runtime & rt = pnol::Runtime::instance();
	vector<double> flat;
	flatten( Xpop, flat );
	int n = Xpop.empty() ? 0 : (int) Xpop[0].size();
	rt.check( pnol_ga_pop_sort( rt.ctx(), flat.data(), F.data(), (long long) Xpop.size(), n ) );
	unflatten( flat, Xpop );
}

static void stageCall( bool identical, vector<vector<double> > & Xpop, std::vector <double> & Xlb, std::vector <double> & Xub, vector <bool> & evaluateIndicator )
{
	pnol::Runtime & rt = pnol::Runtime::instance();
	vector<double> flat;
	flatten( Xpop, flat );
	int n = Xpop.empty() ? 0 : (int) Xpop[0].size();
	vector<unsigned char> ind( evaluateIndicator.size() );
	for( size_t i = 0; i < ind.size(); i++ ) ind[i] = evaluateIndicator[i] ? 1 : 0;
	pnol_stream_desc stream = currentStream( (int) Xpop.size() );
	if( identical ) rt.check( pnol_ga_check_identical( rt.ctx(), flat.data(), (long long) Xpop.size(), n, Xlb.data(), Xub.data(), ind.data(), &stream, &gStagePos ) );
	else rt.check( pnol_ga_check_bounds( rt.ctx(), flat.data(), (long long) Xpop.size(), n, Xlb.data(), Xub.data(), ind.data(), &stream, &gStagePos ) );
	unflatten( flat, Xpop );
	for( size_t i = 0; i < ind.size(); i++ ) evaluateIndicator[i] = ind[i] != 0;
}

// Source/GeneticAlgorithm.cpp:347-365
void checkPopulationBoundsAndReplace( vector<vector<double> > & Xpop, std::vector <double> & Xlb, std::vector <double> & Xub, vector <bool> & evaluateIndicator )
{
	stageCall( false, Xpop, Xlb, Xub, evaluateIndicator );
}

// Source/GeneticAlgorithm.cpp:313-344
void checkIndenticalChildAndReplace( vector<vector<double> > & Xpop, std::vector <double> & Xlb, std::vector <double> & Xub, vector <bool> & evaluateIndicator )
{
	stageCall( true, Xpop, Xlb, Xub, evaluateIndicator );
}
