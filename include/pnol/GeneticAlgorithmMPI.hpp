/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * GeneticAlgorithmMPI.hpp -- GeneticAlgorithmMPI, interface of /root/reference/Source/GeneticAlgorithmMPI.hpp:35-82.
 * The whole generation loop runs on the device (pnol_ga_*): population, fitness sweep, selection, crossover,
 * mutation, repair and the sort. Random numbers come from pnol::Runtime's stream (the reference seeds rand() with
 * time(0), Source/GeneticAlgorithmMPI.cpp:57); without one a counter stream seeded from the clock is used.
 */
#ifndef PNOL_GENETICALGORITHMMPI_HPP_
#define PNOL_GENETICALGORITHMMPI_HPP_

#include <vector>

#include "UtilityFunctions.hpp"
#include "PNOL_Algorithm.hpp"

using namespace std;

// local functions (Source/GeneticAlgorithm.hpp:32-34), host-vector forms running the device stages
void checkPopulationBoundsAndReplace( vector<vector<double> > & Xpop, std::vector <double> & Xlb, std::vector <double> & Xub, vector <bool> & evaluateIndicator );
void checkIndenticalChildAndReplace( vector<vector<double> > & Xpop, std::vector <double> & Xlb, std::vector <double> & Xub, vector <bool> & evaluateIndicator );
void popSort( vector<vector<double> > & Xpop, vector <double> & F );

namespace pnol {
struct GAReport {
	int generations = 0;
	int stoppedStatic = 0;
	unsigned long long streamPos = 0;
	int Nelite = 0, NeliteMut = 0, Ncross = 0, Nrand = 0;
};
void gaFindMinBnd( Objective * obj, int Npop, int maxGenerations, double eliteFrac, double crossFrac, double eliteMutationFrac,
		double mutationSize, double eliteMutationSize, double NstaticGenerations, bool verbose,
		std::vector <double> & X, std::vector <double> & Xlb, std::vector <double> & Xub, double & f0, double & fOpt, GAReport & report );
}

class GeneticAlgorithmMPI : public AlgorithmBnd {
  private:
	int Npop;
	int maxGenerations;
	double eliteFrac, crossFrac, eliteMutationFrac;
	double mutationSize, eliteMutationSize;
	double initialPopScaling;
	double NstaticGenerations;
	bool verbose;
	pnol::GAReport report;

  public:
	void findMinBnd( std::vector <double> & X, std::vector <double> & Xlb, std::vector <double> & Xub, double & f0 , double & fOpt )
	{
		pnol::gaFindMinBnd( objPtr, Npop, maxGenerations, eliteFrac, crossFrac, eliteMutationFrac, mutationSize, eliteMutationSize,
				NstaticGenerations, verbose, X, Xlb, Xub, f0, fOpt, report );
	}

	void setGAParams( int NpopIn, int maxGenerationsIn, double eliteFracIn, double crossFracIn, double eliteMutationFracIn,
			double mutationSizeIn, double eliteMutationSizeIn, double initialPopScalingIn,
			double NstaticGenerationsIn, bool verboseIn )
	{ Npop = NpopIn; eliteFrac = eliteFracIn; crossFrac = crossFracIn; eliteMutationFrac = eliteMutationFracIn;
	maxGenerations = maxGenerationsIn; mutationSize = mutationSizeIn; eliteMutationSize = eliteMutationSizeIn;
	NstaticGenerations = NstaticGenerationsIn; verbose = verboseIn; initialPopScaling = initialPopScalingIn;
	}

	// evaluates the rows whose indicator is set (Source/GeneticAlgorithmMPI.cpp:283-414) -- one fitness sweep kernel
	void evaluatePopulationParallel( vector<vector<double> > & Xpop, vector <double> & F, vector <bool> & evaluateIndicator );

	const pnol::GAReport & lastReport() const { return report; }

	GeneticAlgorithmMPI()
	{
		Npop = 100; maxGenerations = 1000; eliteFrac = 0.1; crossFrac = 0.3; eliteMutationFrac = 0.2; mutationSize = 0.5;
		eliteMutationSize = 0.01; initialPopScaling = 0.5; NstaticGenerations = 50; verbose = 0;
	}
	~GeneticAlgorithmMPI(){}
};

#endif
