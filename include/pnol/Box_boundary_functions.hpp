/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * Box_boundary_functions.hpp -- box-bound helpers, same names as
 * /root/reference/Source/Box_boundary_functions.hpp:28-32 plus computeAlphaBnd / checkAlphaPoolBnd which the
 * reference keeps in BFGS_with_bnd_linsearch_MPI.cpp:665-743. O(n) host arithmetic.
 */
#ifndef PNOL_BOX_BOUNDARY_FUNCTIONS_HPP_
#define PNOL_BOX_BOUNDARY_FUNCTIONS_HPP_

#include <vector>

#include "UtilityFunctions.hpp"

using namespace std;

void checkBoxBounds( vector <double> & X, vector <double> & Xlb, vector <double> & Xub );
void setHardRandValues( vector <double> & X, vector <double> & Xlb, vector <double> & Xub );
double computeAlphaBnd( vector <double> & X, vector <double> & Xlb, vector <double> & Xub, vector <double> & p );
void checkAlphaPoolBnd( bool & bndIndicator, vector <double> & alphaPool, vector <double> & X, vector <double> & Xlb, vector <double> & Xub,
		vector <double> & p, vector<double> & constantX, vector<bool> & constantIndicator );

#endif
