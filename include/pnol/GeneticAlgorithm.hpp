/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * GeneticAlgorithm.hpp -- GeneticAlgorithm (the reference's serial class, /root/reference/Source/GeneticAlgorithm.hpp:37-87).
 * Same device path as GeneticAlgorithmMPI (the two reference classes produce identical populations for one stream);
 * the optional VTK live plot (graph flag) is accepted and ignored.
 */
#ifndef PNOL_GENETICALGORITHM_HPP_
#define PNOL_GENETICALGORITHM_HPP_

#include "GeneticAlgorithmMPI.hpp"

class GeneticAlgorithm : public AlgorithmBnd {
  private:
	int Npop;
	int maxGenerations;
	double eliteFrac, crossFrac, eliteMutationFrac;
	double mutationSize, eliteMutationSize;
	double initialPopScaling;
	double NstaticGenerations;
	bool verbose;
	bool graph;
	pnol::GAReport report;

  public:
	void findMinBnd( std::vector <double> & X, std::vector <double> & Xlb, std::vector <double> & Xub, double & f0 , double & fOpt )
	{
		pnol::LocalScope serial;         // the serial class never touches the communicator (Source/GeneticAlgorithm.cpp)
		pnol::gaFindMinBnd( objPtr, Npop, maxGenerations, eliteFrac, crossFrac, eliteMutationFrac, mutationSize, eliteMutationSize,
				NstaticGenerations, verbose, X, Xlb, Xub, f0, fOpt, report );
	}

	void setGAParams( int NpopIn, int maxGenerationsIn, double eliteFracIn, double crossFracIn, double eliteMutationFracIn,
			double mutationSizeIn, double eliteMutationSizeIn, double initialPopScalingIn,
			double NstaticGenerationsIn, bool verboseIn, bool graphIn )
	{ Npop = NpopIn; eliteFrac = eliteFracIn; crossFrac = crossFracIn; eliteMutationFrac = eliteMutationFracIn;
	maxGenerations = maxGenerationsIn; mutationSize = mutationSizeIn; eliteMutationSize = eliteMutationSizeIn;
	NstaticGenerations = NstaticGenerationsIn; verbose = verboseIn; initialPopScaling = initialPopScalingIn;
	graph = graphIn;
	}

	void evaluatePopulation( vector<vector<double> > & Xpop, vector <double> & F, vector <bool> & evaluateIndicator );

	const pnol::GAReport & lastReport() const { return report; }

	GeneticAlgorithm()
	{
		Npop = 100; maxGenerations = 1000; eliteFrac = 0.1; crossFrac = 0.3; eliteMutationFrac = 0.2; mutationSize = 0.5;
		eliteMutationSize = 0.01; initialPopScaling = 0.5; NstaticGenerations = 50; verbose = 0; graph = 0;
	}
	~GeneticAlgorithm(){}
};

#endif
