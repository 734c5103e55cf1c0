/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * PNOL_Algorithm.hpp -- algorithm base classes, identical in shape to the reference
 * (/root/reference/Source/PNOL_Algorithm.hpp:22-65): non-owning objective pointers set with setObjPtr(), pure
 * virtual findMin / findMinBnd with in/out X.
 */
#ifndef PNOL_ALGORITHM_HPP_
#define PNOL_ALGORITHM_HPP_

#ifndef ROOT_ID
#define ROOT_ID 0 // id of root process
#endif

#include <vector>

#include "PNOL_Objective.hpp"

class AlgorithmBnd {
  protected:
	Objective * objPtr;
  public:
	virtual ~AlgorithmBnd() {}
	virtual void findMinBnd( std::vector <double> & X, std::vector <double> & Xlb, std::vector <double> & Xub, double & f0 , double & fOpt ) = 0;
	void setObjPtr( Objective & obj ){ objPtr = &obj; }
};

class Algorithm {
  protected:
	Objective * objPtr;
  public:
	virtual ~Algorithm() {}
	virtual void findMin( std::vector <double> & X, double & f0 , double & fOpt ) = 0;
	void setObjPtr( Objective & obj ){ objPtr = &obj; }
};

class MultiAlgorithm {
  protected:
	MultiObjective * mObjPtr;
  public:
	virtual ~MultiAlgorithm() {}
	virtual void findMin( std::vector <double> & X, std::vector <double> & F0, std::vector <double> & F ) = 0;
	void setObjPtr( MultiObjective & mObj ){ mObjPtr = &mObj; }
};

#endif /* PNOL_ALGORITHM_HPP_ */
