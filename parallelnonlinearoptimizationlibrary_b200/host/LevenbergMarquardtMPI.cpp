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

	// Storage (:27-38): J and the residual vectors live on the device; there is no JT copy
	DeviceArray J( (size_t) Ndata*Nparam ), F( Ndata ), Ftrial( Ndata ), A( (size_t) Nparam*Nparam ), JTJ( (size_t) Nparam*Nparam );
	DeviceArray rhs( Nparam ), dXdev( Nparam ), Xdev( Nparam ), sigmaDev( Nparam );
	vector <double> dX( Nparam, dXGrad );
	vector <double> sigma( Nparam, 0 );
	vector <double> Xprev( Nparam, 0 );
	dXdev.upload( dX.data(), Nparam );

	// Initial f values (:42-49)
	double sumsq = 0;
	rt.check( pnol_residual_eval( ctx, f, X.data(), Nparam, F.data(), &sumsq ) );
	F.download( F0.data(), Ndata );
	for( int k = 0; k < Nparam; k++ ) Xprev[k] = X[k];
	chiSq = pow( sqrt(sumsq), 2 );                               // pow(vector2Norm(F),2)  (:51)

	int iter = 0;
	double xdiff2Norm = xMinDiff*2;
	bool jacobianCurrent = false;      // J^T J is unchanged after a rejected step (X was restored): SURVEY.md 3.1
	report = LMReport();
	while( iter < maxIter )
	{
		// Update the gradient and the normal equations (:60-85). After a rejected step the reference recomputes the
		// identical J; re-damping the stored J^T J gives the same A.
		if( !jacobianCurrent )
		{
			Xdev.upload( X.data(), Nparam );
			rt.check( pnol_fd_jacobian( ctx, f, Xdev.data(), dXdev.data(), Nparam, J.data(), nullptr, rt.jacobianMode() ) );
			rt.check( pnol_lm_normal_eq( ctx, J.data(), F.data(), Ndata, Nparam, lambda, JTJ.data(), A.data(), rhs.data() ) );
			jacobianCurrent = rt.jacobianCache();
		}
		else
		{
			rt.check( pnol_lm_damp( ctx, JTJ.data(), Nparam, lambda, A.data() ) );
		}

		// Solve for sigma (:88)
		int info = 0;
		int st = pnol_spd_solve( ctx, A.data(), rhs.data(), Nparam, sigma.data(), &info );
		if( st == PNOL_ERR_NOT_SPD )
		{
			// a zero / negative pivot (e.g. a zero column of J): the reference's luSolve would return inf/NaN and the
			// step would be rejected through the NaN test at :110. Same outcome here.
			for( int i = 0; i < Nparam; i++ ) sigma[i] = NAN;
		}
		else rt.check( st );

		// store previous, update parameters (:91-100)
		for( int k = 0; k < Nparam; k++ ) Xprev[k] = X[k];
		for( int i = 0; i < Nparam; i++ ) X[i] = X[i] + sigma[i];

		// Update F and chi (:103-108)
		rt.check( pnol_residual_eval( ctx, f, X.data(), Nparam, Ftrial.data(), &sumsq ) );
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

} // namespace pnol

void LevMarqMPI::findMin( vector <double> & X, vector <double> & F0, vector <double> & FOpt )
{
	pnol::lmFindMin( mObjPtr, lambda0, lambdaFactor, dXGrad, maxIter, xMinDiff, verbose, X, F0, FOpt, report );
}
