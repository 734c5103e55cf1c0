// TEST INFRASTRUCTURE -- the Levenberg-Marquardt restatement of oracle/pnol_oracle.cpp with the loops re-nested and
// dealt to host threads so that it finishes at the FULL BASELINE shapes (cfg5: m = 4M, n = 256; cfg2: m = 100k, n = 16).
// Same pinning statement and the same access rules as oracle/pnol_oracle.cpp (tests/, smoke() and bench.py's CPU legs only).
//
// Every floating-point value is produced by the same operations in the same order as in oracle_lm / oracle_fd_jacobian /
// oracle_lm_normal_eq (and therefore as in Source/LevenbergMarquardtMPI.cpp:12-173 + Source/PNOL_Objective.cpp:165-197 with
// the shim's sequential sums):
//   * residuals and FD Jacobian entries are independent per row -> rows are dealt to threads;
//   * JTJ[i][j] = sum_k J[k][i] J[k][j] and rhs[i] = -sum_k J[k][i] F[k] add their terms in ascending k. Here k is the OUTER
//     loop (one pass over J instead of n^2 column-strided passes) and threads own bands of i: each accumulator still sees its
//     terms in ascending k, one rounding per multiply and one per add (-ffp-contract=off) -- bit-identical sums;
//   * sqrt-of-sum-of-squares (vector2Norm) stays one sequential loop.
// tests/test_golden_oracle.py checks oracle_lm_mt == oracle_lm bit for bit at sizes both finish.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>

using std::vector;

#include "oracle_objectives.h"

extern "C" void oracle_lu_solve(const double * A, const double * b, int n, double * x);

namespace {

template <class Fn> void parallel_ranges(long long count, int threads, Fn fn)
{
	if (threads < 1) threads = 1;
	if ((long long) threads > count) threads = count > 0 ? (int) count : 1;
	vector<std::thread> th;
	for (int t = 0; t < threads; t++) {
		long long lo = count * t / threads, hi = count * (t + 1) / threads;
		th.emplace_back([=]() { fn(lo, hi); });
	}
	for (auto & x : th) x.join();
}

// F[lo, hi) of the residual model: the functor restricted to those rows (data columns advanced) is the same arithmetic
void residual_rows(const OFunctor & f, const double * X, int n, double * F, int threads)
{
	parallel_ranges(f.m, threads, [&](long long lo, long long hi) {
		const double * cols[8];
		for (int c = 0; c < 8; c++) cols[c] = f.cols[c] ? f.cols[c] + lo : nullptr;
		OFunctor view = {f.kind, f.scalars, f.ints, cols, hi - lo};
		o_residual(view, X, n, F + lo);
	});
}

// MultiObjective::gradientApproximation (Source/PNOL_Objective.cpp:165-197): J[i][j] = (FdX_j[i] - F[i]) / dX[j]
void jacobian_rows(const OFunctor & f, const double * X, const double * dX, int N, const double * F, double * J, int threads)
{
	parallel_ranges(f.m, threads, [&](long long lo, long long hi) {
		const long long blk = 4096;                    // rows per pass: FdX of a block stays in cache
		vector<double> XdX(N), FdX(blk);
		for (long long r0 = lo; r0 < hi; r0 += blk) {
			const long long nr = hi - r0 < blk ? hi - r0 : blk;
			const double * cols[8];
			for (int c = 0; c < 8; c++) cols[c] = f.cols[c] ? f.cols[c] + r0 : nullptr;
			OFunctor view = {f.kind, f.scalars, f.ints, cols, nr};
			for (int j = 0; j < N; j++) {
				for (int k = 0; k < N; k++) XdX[k] = X[k];
				XdX[j] = XdX[j] + dX[j];
				o_residual(view, XdX.data(), N, FdX.data());
				for (long long i = 0; i < nr; i++) J[(size_t) (r0 + i) * N + j] = (FdX[i] - F[r0 + i]) / dX[j];
			}
		}
	});
}

// Source/LevenbergMarquardtMPI.cpp:64-85 with the shim's sequential-k products, k outermost
void normal_eq(const double * J, const double * F, long long m, int n, double lambda, double * A, double * rhs, int threads)
{
	vector<double> S((size_t) n * n, 0.0), R(n, 0.0);
	parallel_ranges(n, threads, [&](long long i0, long long i1) {
		for (long long k = 0; k < m; k++) {
			const double * row = J + (size_t) k * n;
			const double fk = F[k];
			for (long long i = i0; i < i1; i++) {
				const double a = row[i];
				double * s = S.data() + (size_t) i * n;
				for (int j = 0; j < n; j++) s[j] = s[j] + a * row[j];
				R[i] = R[i] + a * fk;
			}
		}
	});
	for (int i = 0; i < n; i++) {
		for (int j = 0; j < n; j++) A[(size_t) i * n + j] = (i == j) ? (1 + lambda) * S[(size_t) i * n + j] : S[(size_t) i * n + j];
		rhs[i] = -R[i];
	}
}

double norm2(const double * v, long long n)
{
	double s = 0;
	for (long long i = 0; i < n; i++) s = s + v[i] * v[i];
	return sqrt(s);
}

} // namespace

extern "C" {

// LevMarqMPI::findMin (Source/LevenbergMarquardtMPI.cpp:12-173); arguments as oracle_lm plus the thread count.
// trace rows: [X after the pass | chiSq | lambda]
int oracle_lm_mt(int kind, const double * scalars, const long long * ints, const double * const * cols, long long m,
                 double * X, int Nparam, double lambda0, double lambdaFactor, double dXGrad, int maxIter, double xMinDiff,
                 double * F0, double * FOpt, double * chisq_out, double * lambda_out, double * trace, int threads)
{
	OFunctor f = {kind, scalars, ints, cols, m};
	const long long Ndata = m;
	double lambda = lambda0;
	vector<double> A((size_t) Nparam * Nparam), J((size_t) Ndata * Nparam), rhs(Nparam, 0.0), dX(Nparam, dXGrad);
	vector<double> F(Ndata, 0), Fprev(Ndata, 0), sigma(Nparam, 0), Xprev(Nparam, 0);

	residual_rows(f, X, Nparam, F.data(), threads);                          // :42
	if (F0) memcpy(F0, F.data(), (size_t) Ndata * sizeof(double));
	double chiSq = pow(norm2(F.data(), Ndata), 2);                           // :51
	int iter = 0;
	while (iter < maxIter) {                                                 // :55
		jacobian_rows(f, X, dX.data(), Nparam, F.data(), J.data(), threads); // :60 (F(X) is what the stencil re-evaluates)
		normal_eq(J.data(), F.data(), Ndata, Nparam, lambda, A.data(), rhs.data(), threads);   // :64-85
		oracle_lu_solve(A.data(), rhs.data(), Nparam, sigma.data());         // :88
		for (int k = 0; k < Nparam; k++) Xprev[k] = X[k];                    // :91-94
		Fprev.swap(F);
		for (int i = 0; i < Nparam; i++) X[i] = X[i] + sigma[i];             // :97-100
		residual_rows(f, X, Nparam, F.data(), threads);                      // :103
		const double chiSqPrev = chiSq;                                      // :107
		chiSq = pow(norm2(F.data(), Ndata), 2);                              // :108
		bool stop = false;
		if (chiSq >= chiSqPrev || chiSq != chiSq) {                          // :110
			chiSq = chiSqPrev;
			for (int i = 0; i < Nparam; i++) X[i] = Xprev[i];
			F.swap(Fprev);
			lambda = lambda * lambdaFactor;                                  // :129
		} else {
			lambda = lambda / lambdaFactor;                                  // :135
			if (norm2(sigma.data(), Nparam) < xMinDiff) stop = true;         // :138-140
		}
		if (trace) {
			double * row = trace + (size_t) iter * (Nparam + 2);
			for (int i = 0; i < Nparam; i++) row[i] = X[i];
			row[Nparam] = chiSq; row[Nparam + 1] = lambda;
		}
		if (stop) break;
		iter++;
	}
	if (FOpt) memcpy(FOpt, F.data(), (size_t) Ndata * sizeof(double));       // :159-162
	if (chisq_out) *chisq_out = chiSq;
	if (lambda_out) *lambda_out = lambda;
	return iter;
}

}
