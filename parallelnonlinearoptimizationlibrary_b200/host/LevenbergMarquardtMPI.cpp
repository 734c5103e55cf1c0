/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
// LevenbergMarquardtMPI.cpp -- LevMarqMPI::findMin / LevMarq::findMin: the control flow of
// Source/LevenbergMarquardtMPI.cpp:12-173 (serial twin Source/LevenbergMarquardt.cpp:11-177) on the host, every
// O(m) or O(n^2) statement as a kernel through the C-ABI. J, F and the trial F never leave the device.
#include "pnol/LevenbergMarquardtMPI.hpp"

#include <cmath>
#include <iostream>

namespace pnol {

void lmFindMin( MultiObjective * mObjPtr, double lambda0, double lambdaFactor, double dXGrad, int maxIter, double xMinDiff,
		int verbose, vector <double> & X, vector <double> & F0, vector <double> & FOpt, LMReport & report )
{
	Runtime & rt = Runtime::instance();
	pnol_ctx * ctx = rt.ctx();
	pnol_functor * f = mObjPtr->requireFunctor( "LevMarq::findMin" );
	const bool root = pnol_comm_rank( ctx ) == ROOT_ID;

	double chiSq;
	double lambda = lambda0;

	// Problem size (Source/LevenbergMarquardtMPI.cpp:23-24); Ndata is this rank's row block
	const int Nparam = (int) X.size();
	const long long Ndata = pnol_functor_rows( f );
	if( (long long) F0.size() != Ndata || (long long) FOpt.size() != Ndata )
		throw Error( PNOL_ERR_INVALID, "LevMarq::findMin: F0 / FOpt must be pre-sized to the number of residuals" );

	// Storage (:27-38): J and the residual vectors live on the device; there is no JT copy. JTJ holds J^T J followed by -J^T F.
	// With Runtime::setStoreJacobian(false) no J is allocated: pnol_lm_step sums the normal equations over row blocks.
	DeviceArray J( rt.storeJacobian() ? (size_t) Ndata*Nparam : 0 ), F( Ndata ), Ftrial( Ndata ), JTJ( (size_t) Nparam*Nparam + Nparam );
	DeviceArray dXdev( Nparam );
	vector <double> dX( Nparam, dXGrad );
	vector <double> sigma( Nparam, 0 );
	vector <double> Xprev( Nparam, 0 ), Xtrial( Nparam, 0 );
	dXdev.upload( dX.data(), Nparam );

	// whether the while loop runs on device-resident state (see below)
	const bool deviceLoop = verbose < 1 && !rt.jacobianCache() && maxIter > 0;

	// Initial f values (:42-49)
	double sumsq = 0;
	rt.check( pnol_residual_eval( ctx, f, X.data(), Nparam, F.data(), &sumsq ) );
	// F0 goes back to the host; with the device loop the read-back runs beside the iterations (from a copy: F changes with the
	// first accepted step), and findMin waits for it before it returns -- also when it leaves through an exception
	DeviceArray F0dev( deviceLoop ? (size_t) Ndata : 0 );
	struct CopyGuard { pnol_ctx * c; bool on; ~CopyGuard() { if( on ) pnol_copy_wait( c ); } } f0Guard{ ctx, false };
	if( deviceLoop )
	{
		rt.check( pnol_memcpy( ctx, F0dev.data(), F.data(), (size_t) Ndata*sizeof(double) ) );
		rt.check( pnol_copy_start( ctx, F0.data(), F0dev.data(), (size_t) Ndata*sizeof(double) ) );
		f0Guard.on = true;
	}
	else
		F.download( F0.data(), Ndata );
	for( int k = 0; k < Nparam; k++ ) Xprev[k] = X[k];
	chiSq = pow( sqrt(sumsq), 2 );                               // pow(vector2Norm(F),2)  (:51)

	int iter = 0;
	double xdiff2Norm = xMinDiff*2;
	bool jacobianCurrent = false;      // J^T J is unchanged after a rejected step (X was restored): SURVEY.md 3.1
	report = LMReport();

	// Nobody watches the iterations (verbose < 1) and J is recomputed in every pass as in the reference (jacobianCache off): the whole
	// while loop (:55-141) runs on device-resident state, accept / reject rule included (pnol_lm_iterate: same arithmetic, same
	// decisions, X / lambda / chi^2 bit for bit those of the loop below -- tests/test_gpu_host_api.py), with one synchronisation per
	// batch of iterations instead of one per iteration.
	if( deviceLoop )
	{
		int accepted = 0, rejected = 0, swapped = 0, stopped = 0;
		double xdiff = 0;
		rt.check( pnol_lm_iterate( ctx, f, X.data(), dXdev.data(), Nparam, rt.storeJacobian() ? J.data() : nullptr, F.data(), Ftrial.data(),
				JTJ.data(), &lambda, &chiSq, lambdaFactor, xMinDiff, maxIter, rt.jacobianMode(), &accepted, &rejected, &swapped ) );
		rt.check( pnol_lm_last_run( ctx, &stopped, &xdiff ) );
		if( swapped ) F.swap( Ftrial );
		report.accepted = accepted;
		report.rejected = rejected;
		if( accepted > 0 ) xdiff2Norm = xdiff;
		iter = accepted + rejected - ( stopped ? 1 : 0 );            // the pass that meets the stopping rule leaves through `break` (:138-140)
		maxIter = 0;                                                 // (skips the host loop below)
		f0Guard.on = false;
		rt.check( pnol_copy_wait( ctx ) );                           // F0 has landed
	}

	while( iter < maxIter )
	{
		// Gradient, normal equations, damped solve, parameter update and the trial residuals (:60-103) are one device call with
		// one synchronisation. After a rejected step the reference recomputes the identical J (:60 after :120-129); re-damping
		// the stored J^T J gives the same A, so that work is skipped unless the runtime asks for it (jacobianCache off).
		// A zero / negative pivot (e.g. a zero column of J) comes back as a NaN step: the reference's luSolve would return
		// inf/NaN there and the step would be rejected through the NaN test at :110. Same outcome here.
		int info = 0;
		rt.check( pnol_lm_step( ctx, f, X.data(), dXdev.data(), Nparam, rt.storeJacobian() ? J.data() : nullptr, F.data(), Ftrial.data(), lambda, rt.jacobianMode(),
				jacobianCurrent ? 1 : 0, JTJ.data(), sigma.data(), Xtrial.data(), &sumsq, &info ) );
		jacobianCurrent = rt.jacobianCache();

		// store previous, update parameters (:91-100): Xtrial[i] is X[i] + sigma[i], added on the device
		for( int k = 0; k < Nparam; k++ ) Xprev[k] = X[k];
		for( int i = 0; i < Nparam; i++ ) X[i] = Xtrial[i];

		// chi (:107-108)
		double chiSqPrev = chiSq;
		chiSq = pow( sqrt(sumsq), 2 );

		if( chiSq >= chiSqPrev || chiSq != chiSq )
		{
			if( verbose >= 1 && root )
			{
				cout << "Step " << iter << " failed with chiSq = " << chiSq << ", chiSqPrev = " << chiSqPrev;
				cout << ",  increasing lambda: " << lambda/lambdaFactor << " --> " << lambda << endl;
			}
			// reset X (F on the device was never overwritten) and increase lambda (:118-129)
			chiSq = chiSqPrev;
			for( int i = 0; i < Nparam; i++ ) X[i] = Xprev[i];
			lambda = lambda*lambdaFactor;
			report.rejected++;
		}
		else
		{
			// keep the new step: the trial residuals become F (pointer swap instead of the copy at :91-94)
			lambda = lambda/lambdaFactor;
			F.swap( Ftrial );
			jacobianCurrent = false;
			report.accepted++;

			// Check stopping criterion (:138-140)
			xdiff2Norm = vector2Norm( sigma );
			if( xdiff2Norm < xMinDiff )
				break;
		}

		if( verbose >= 1 && mod(iter,10) == 0 && root )
		{
			cout << "At iter = " << iter << " the xdiff 2Norm = " << xdiff2Norm << ", chi^2 = " << chiSq << ", and params: ";
			print1DVector( X );
		}

		iter++;
	}

	F.download( FOpt.data(), Ndata );                            // (:159-162)
	report.iterations = iter;
	report.chiSq = chiSq;
	report.lambda = lambda;
	report.xdiff2Norm = xdiff2Norm;
	if( verbose >= 0 && root )
	{
		cout << endl << "-----------------------------------------------------------------------------------" << endl;
		cout << "Completed Levenberg Marquardt." << endl;
		cout << "At iter = " << iter << " the xdiff 2Norm = " << xdiff2Norm << ", chi^2 = " << chiSq << ", and  optimal params: " << endl;
		print1DVector( X );
		cout << "-----------------------------------------------------------------------------------" << endl << endl;
	}
}

// The reference's findMin hands the FULL F0 / FOpt to every rank; under row sharding a rank only has its block (see the header).
// This rebuilds the full vector: block lengths travel in one small all-reduce, the blocks in one all-gather of padded slots.
void gatherResiduals( const vector <double> & local, vector <double> & full )
{
	Runtime & rt = Runtime::instance();
	pnol_ctx * ctx = rt.ctx();
	const int R = pnol_comm_size( ctx ), r = pnol_comm_rank( ctx );
	if( R <= 1 ) { full = local; return; }
	vector <double> lens( R, 0.0 );
	lens[r] = (double) local.size();
	rt.check( pnol_comm_allreduce_sum( ctx, lens.data(), (size_t) R ) );
	size_t slot = 0, total = 0;
	for( int k = 0; k < R; k++ ) { if( (size_t) lens[k] > slot ) slot = (size_t) lens[k]; total += (size_t) lens[k]; }
	if( slot == 0 ) { full.clear(); return; }
	vector <double> send( slot, 0.0 ), recv( slot*R, 0.0 );
	for( size_t i = 0; i < local.size(); i++ ) send[i] = local[i];
	rt.check( pnol_comm_allgather( ctx, send.data(), recv.data(), slot ) );
	full.resize( total );
	size_t at = 0;
	for( int k = 0; k < R; k++ )
		for( size_t i = 0; i < (size_t) lens[k]; i++ ) full[at++] = recv[(size_t) k*slot + i];
}

} // namespace pnol

void LevMarqMPI::findMin( vector <double> & X, vector <double> & F0, vector <double> & FOpt )
{
	pnol::lmFindMin( mObjPtr, lambda0, lambdaFactor, dXGrad, maxIter, xMinDiff, verbose, X, F0, FOpt, report );
}
