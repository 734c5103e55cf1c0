/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * ExampleObjectives.hpp -- the reference's example objectives (/root/reference/Source/ExampleObjectives.hpp) with
 * device twins, plus the two synthetic objectives of BASELINE.json (Rastrigin, sum of Lorentzians).
 *
 * Every class keeps its reference name and members (getEvals(), setPower(), getDataSize()); objEval evaluates the
 * SAME __host__ __device__ functor (include/pnol/functors.hpp) the kernels run, so host and device values agree to
 * the last bit. Differences from the reference, all forced by device reproducibility:
 *   - pow(x, 2) is written x*x (what GCC emits for the reference at -O2/-O3 anyway);
 *   - PowerObject multiplies repeatedly instead of calling pow(x, power) with a run-time exponent;
 *   - ExpCurve* use pnol::exp_hd (include/pnol/pnol_math.h) instead of libm exp; their DATA are still generated with
 *     libm exp/pow on the host exactly as the reference constructors do.
 * PowerObjectSlow (:236-265) keeps its name and values; its deliberately slow busy-loop is not reproduced.
 */
#ifndef PNOL_EXAMPLEOBJECTIVES_HPP_
#define PNOL_EXAMPLEOBJECTIVES_HPP_

#include <math.h>

#include "PNOL_Objective.hpp"
#include "UtilityFunctions.hpp"
#include "functors.hpp"

namespace pnol {
// scalar objective backed by functor F (no data columns)
template <class F> class FunctorObjective : public Objective {
  protected:
	int evals;
	FunctorParams P;
	DeviceFunctor dev;
  public:
	FunctorObjective() : evals(0) { memset( &P, 0, sizeof P ); }
	double objEval( vector <double> & X )
	{
		evals++;
		PtrAcc acc{ X.data() };
		return F::eval( P, acc, (int) X.size() );
	}
	pnol_functor * deviceFunctor()
	{
		return dev.get( F::kKind, vector<double>( P.scalars, P.scalars + PNOL_MAX_SCALARS ), vector<long long>( P.ints, P.ints + PNOL_MAX_INTS ) );
	}
	void noteDeviceEvaluations( long long count ) { evals += (int) count; }
	// host objEval calls + points evaluated on the device twin: for a serial algorithm the number the reference prints
	// ("Optimization used ... function evaluations"); the reference's MPI classes count per rank, this is the total
	double getEvals(){ return evals; }
};
}

class GoldsteinFunction : public pnol::FunctorObjective<pnol::GoldsteinFunctor> {};
class BoothFunction : public pnol::FunctorObjective<pnol::BoothFunctor> {};
class RosenbrockObject : public pnol::FunctorObjective<pnol::RosenbrockFunctor> {};
class RastriginObject : public pnol::FunctorObjective<pnol::RastriginFunctor> {};

class PowerObject : public pnol::FunctorObjective<pnol::PowerFunctor> {
  public:
	void setPower( int powIn ){ P.ints[0] = powIn; dev.release(); }
	int getPower(){ return (int) P.ints[0]; }
	PowerObject(){ P.ints[0] = 2; }
};

// Source/ExampleObjectives.hpp:238-271: PowerObject behind a busy loop ("slow computation for no reason") that the reference uses to
// make testGAParallel's evaluations expensive. The values are PowerObject's; a device functor has nothing to wait for.
class PowerObjectSlow : public PowerObject {};

namespace pnol {
// residual model backed by functor R with host-owned data columns
template <class R> class FunctorMultiObjective : public MultiObjective {
  protected:
	FunctorParams P;
	vector<vector<double> > cols;
	DeviceFunctor dev;
	void bindColumns()
	{
		for( size_t c = 0; c < cols.size(); c++ ) P.col[c] = cols[c].data();
		P.m = cols.empty() ? 0 : (long long) cols[0].size();
		dev.release();
	}
  public:
	FunctorMultiObjective(){ memset( &P, 0, sizeof P ); }
	void objEval( vector <double> & X, vector <double> & F )
	{
		PtrAcc acc{ X.data() };
		for( long long k = 0; k < P.m; k++ ) F[k] = R::residual( P, acc, (int) X.size(), k );
	}
	pnol_functor * deviceFunctor()
	{
		vector<const double *> ptrs;
		for( size_t c = 0; c < cols.size(); c++ ) ptrs.push_back( cols[c].data() );
		return dev.get( R::kKind, vector<double>( P.scalars, P.scalars + PNOL_MAX_SCALARS ), vector<long long>( P.ints, P.ints + PNOL_MAX_INTS ), ptrs, P.m );
	}
	// drop the device twin (the data columns are uploaded again by the next deviceFunctor() call)
	void releaseDeviceFunctor(){ dev.release(); }
	int getDataSize(){ return (int) P.m; }
};
}

// Source/ExampleObjectives.hpp:106-154
class ExpCurveObjective : public pnol::FunctorMultiObjective<pnol::ExpCurveFunctor> {
  public:
	ExpCurveObjective()
	{
		cols.resize(2);
		linspace( 0, 5, 100, cols[0] );
		cols[1].resize( cols[0].size() );
		for( size_t k = 0; k < cols[0].size(); k++ ) cols[1][k] = 10.2*exp( 0.4*cols[0][k] ) + 0.1;
		bindColumns();
	}
};

// Source/ExampleObjectives.hpp:160-201; the x^3 column is what the reference's pow(xData[k],3) returns
class CubicObjective : public pnol::FunctorMultiObjective<pnol::CubicFunctor> {
  public:
	CubicObjective()
	{
		vector<double> xData;
		linspace( -5, 5, 100, xData );
		cols.resize(3);
		cols[0].resize( xData.size() ); cols[1] = xData; cols[2].resize( xData.size() );
		for( size_t k = 0; k < xData.size(); k++ )
		{
			cols[0][k] = pow( xData[k], 3 );
			cols[2][k] = 0.3*pow( xData[k], 3 ) + 1.1*pow( xData[k], 2 ) - 4.3*xData[k] + 7.3;
		}
		bindColumns();
	}
};

// ours: y - tree-sum_k a_k / (1 + w (t - c_k)^2); data supplied by the caller (host arrays are copied)
class LorentzSumObjective : public pnol::FunctorMultiObjective<pnol::LorentzSumFunctor> {
  public:
	LorentzSumObjective( const vector<double> & t, const vector<double> & y, double w )
	{
		cols.resize(2);
		cols[0] = t; cols[1] = y;
		P.scalars[0] = w;
		bindColumns();
	}
};

// Source/ExampleObjectives.hpp:270-320
class ExpCurveObjectiveSingle : public Objective {
  private:
	pnol::FunctorParams P;
	vector<double> xData, yData;
	pnol::DeviceFunctor dev;
  public:
	double objEval( vector <double> & X )
	{
		pnol::PtrAcc acc{ X.data() };
		return pnol::ExpCurveSingleFunctor::eval( P, acc, (int) X.size() );
	}
	pnol_functor * deviceFunctor()
	{
		vector<const double *> ptrs; ptrs.push_back( xData.data() ); ptrs.push_back( yData.data() );
		return dev.get( PNOL_F_EXPCURVE_SINGLE, vector<double>(), vector<long long>(), ptrs, (long long) xData.size() );
	}
	int getDataSize(){ return (int) xData.size(); }
	ExpCurveObjectiveSingle()
	{
		memset( &P, 0, sizeof P );
		linspace( 0, 5, 100, xData );
		yData.resize( xData.size() );
		for( size_t k = 0; k < xData.size(); k++ ) yData[k] = 10.2*exp( 0.4*xData[k] ) + 0.1;
		P.col[0] = xData.data(); P.col[1] = yData.data(); P.m = (long long) xData.size();
	}
};

#endif /* PNOL_EXAMPLEOBJECTIVES_HPP_ */
