// dense.cu -- the small dense FP64 pieces around the DMMA kernels: Marquardt damping epilogue, Cholesky solve,
// search-direction matvec and the BFGS inverse-Hessian update (rank-2 and literal forms).
#include "common.cuh"

namespace pnol {

// ---------------------------------------------------------------------------------------------------
// a9: A = JTJ, A_ii = (1 + lambda) JTJ_ii ; rhs = -J^T F     (Source/LevenbergMarquardtMPI.cpp:66-85)
// ---------------------------------------------------------------------------------------------------
__global__ void lm_damp_kernel(const double * __restrict__ packed, int n, double lambda, double * __restrict__ JTJ,
                               double * __restrict__ A, double * __restrict__ rhs)
{
	long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	long long total = (long long) n * n;
	if (idx < total) {
		double v = packed[idx];
		if (JTJ) JTJ[idx] = v;
		if (A) {
			int i = (int) (idx / n), j = (int) (idx - (long long) i * n);
			A[idx] = (i == j) ? (1 + lambda) * v : v;
		}
	} else if (idx < total + n && rhs) {
		rhs[idx - total] = -packed[idx];
	}
}

int launch_lm_damp(pnol_ctx * ctx, const double * packed, int n, double lambda, double * JTJ, double * A, double * rhs)
{
	long long total = (long long) n * n + n;
	PNOL_LAUNCH(ctx, lm_damp_kernel, (unsigned) ((total + 255) / 256), 256, 0, packed, n, lambda, JTJ, A, rhs);
	return PNOL_OK;
}

// ---------------------------------------------------------------------------------------------------
// a9: SPD solve by blocked Cholesky in ONE CTA (n <= 704). Replaces luSolve (LevenbergMarquardtMPI.cpp:88).
// The factor lives in a global n x n scratch (L2 resident); the 32-wide diagonal block and the panel below it
// are staged in shared memory.
// ---------------------------------------------------------------------------------------------------
constexpr int kCholNB = 32;
constexpr int kCholThreads = 512;
constexpr int kCholMaxN = 704;

__global__ void __launch_bounds__(kCholThreads, 1)
spd_solve_kernel(const double * __restrict__ A, const double * __restrict__ rhs, int n, double * __restrict__ L,
                 double * __restrict__ x, int * __restrict__ info)
{
	extern __shared__ double sm[];
	double * Dk = sm;                          // 32 x 33 diagonal block
	double * Pn = sm + kCholNB * (kCholNB + 1); // (n - 32) x 33 panel
	double * yv = Pn + (size_t) (n > kCholNB ? n - kCholNB : 0) * (kCholNB + 1);   // n : rhs / solution
	__shared__ int bad;
	const int tid = threadIdx.x;
	constexpr int P = kCholNB + 1;
	if (tid == 0) bad = 0;
	// copy lower triangle
	for (long long e = tid; e < (long long) n * n; e += kCholThreads) {
		int i = (int) (e / n), j = (int) (e - (long long) i * n);
		L[e] = (j <= i) ? A[e] : 0.0;
	}
	for (int i = tid; i < n; i += kCholThreads) yv[i] = rhs[i];
	__syncthreads();

	for (int kb = 0; kb < n; kb += kCholNB) {
		const int nbk = min(kCholNB, n - kb);
		const int below = n - (kb + nbk);
		// 1. diagonal block -> smem
		for (int e = tid; e < nbk * nbk; e += kCholThreads) {
			int r = e / nbk, c = e - r * nbk;
			Dk[r * P + c] = L[(long long) (kb + r) * n + kb + c];
		}
		__syncthreads();
		// 2. unblocked Cholesky of the diagonal block by warp 0 (lane = row)
		if (tid < 32) {
			const int r = tid;
			for (int j = 0; j < nbk; j++) {
				double d = Dk[j * P + j];
				if (!(d > 0.0)) { if (r == 0 && bad == 0) bad = kb + j + 1; d = 1.0; }
				double sd = sqrt(d);
				__syncwarp();
				if (r == j) Dk[j * P + j] = sd;
				if (r > j && r < nbk) Dk[r * P + j] = Dk[r * P + j] / sd;
				__syncwarp();
				if (r > j && r < nbk) {
					double lrj = Dk[r * P + j];
					for (int c = j + 1; c <= r; c++) Dk[r * P + c] = fma(-lrj, Dk[c * P + j], Dk[r * P + c]);
				}
				__syncwarp();
			}
		}
		__syncthreads();
		// write the factored block back
		for (int e = tid; e < nbk * nbk; e += kCholThreads) {
			int r = e / nbk, c = e - r * nbk;
			if (c <= r) L[(long long) (kb + r) * n + kb + c] = Dk[r * P + c];
		}
		// 3. panel solve: row i of the panel, X L_kk^T = A_ik
		for (int ri = tid; ri < below; ri += kCholThreads) {
			const long long grow = (long long) (kb + nbk + ri) * n + kb;
			double row[kCholNB];
#pragma unroll
			for (int c = 0; c < kCholNB; c++) row[c] = c < nbk ? L[grow + c] : 0.0;
#pragma unroll
			for (int c = 0; c < kCholNB; c++) {
				if (c < nbk) {
					double s = row[c];
#pragma unroll
					for (int t = 0; t < kCholNB; t++)
						if (t < c) s = fma(-row[t], Dk[c * P + t], s);
					row[c] = s / Dk[c * P + c];
				}
			}
#pragma unroll
			for (int c = 0; c < kCholNB; c++) {
				if (c < nbk) { L[grow + c] = row[c]; Pn[ri * P + c] = row[c]; }
			}
		}
		__syncthreads();
		// 4. trailing update of the lower triangle: L[i][j] -= sum_t P[i][t] P[j][t]
		{
			const int base = kb + nbk;
			// enumerate (i, j), j <= i, over a below x below square; threads along j for coalescing
			for (long long e = tid; e < (long long) below * below; e += kCholThreads) {
				int i = (int) (e / below), j = (int) (e - (long long) i * below);
				if (j > i) continue;
				double s = 0;
#pragma unroll 8
				for (int t = 0; t < kCholNB; t++)
					if (t < nbk) s = fma(Pn[i * P + t], Pn[j * P + t], s);
				L[(long long) (base + i) * n + base + j] -= s;
			}
		}
		__syncthreads();
	}

	// forward substitution L y = b (blocked; the diagonal block is staged in shared memory)
	for (int kb = 0; kb < n; kb += kCholNB) {
		const int nbk = min(kCholNB, n - kb);
		for (int e = tid; e < nbk * nbk; e += kCholThreads) {
			int r = e / nbk, c = e - r * nbk;
			Dk[r * P + c] = L[(long long) (kb + r) * n + kb + c];
		}
		__syncthreads();
		if (tid < 32) {
			for (int k = 0; k < nbk; k++) {
				double yk = yv[kb + k] / Dk[k * P + k];
				__syncwarp();
				if (tid == 0) yv[kb + k] = yk;
				int r = k + 1 + tid;
				if (r < nbk) yv[kb + r] = fma(-Dk[r * P + k], yk, yv[kb + r]);
				__syncwarp();
			}
		}
		__syncthreads();
		for (int i = kb + nbk + tid; i < n; i += kCholThreads) {
			double s = yv[i];
			const double * Lr = L + (long long) i * n + kb;
			for (int t = 0; t < nbk; t++) s = fma(-Lr[t], yv[kb + t], s);
			yv[i] = s;
		}
		__syncthreads();
	}
	// back substitution L^T x = y (blocked, descending)
	for (int kb = ((n - 1) / kCholNB) * kCholNB; kb >= 0; kb -= kCholNB) {
		const int nbk = min(kCholNB, n - kb);
		for (int e = tid; e < nbk * nbk; e += kCholThreads) {
			int r = e / nbk, c = e - r * nbk;
			Dk[r * P + c] = L[(long long) (kb + r) * n + kb + c];
		}
		__syncthreads();
		if (tid < 32) {
			for (int k = nbk - 1; k >= 0; k--) {
				double xk = yv[kb + k] / Dk[k * P + k];
				__syncwarp();
				if (tid == 0) yv[kb + k] = xk;
				int r = tid;
				if (r < k) yv[kb + r] = fma(-Dk[k * P + r], xk, yv[kb + r]);
				__syncwarp();
			}
		}
		__syncthreads();
		for (int i = tid; i < kb; i += kCholThreads) {
			double s = yv[i];
			for (int t = 0; t < nbk; t++) s = fma(-L[(long long) (kb + t) * n + i], yv[kb + t], s);
			yv[i] = s;
		}
		__syncthreads();
	}
	for (int i = tid; i < n; i += kCholThreads) x[i] = yv[i];
	if (tid == 0) *info = bad;
}

int launch_spd_solve(pnol_ctx * ctx, const double * A, const double * rhs, int n, double * x, int * info_dev)
{
	PNOL_REQUIRE(ctx, n >= 1 && n <= kCholMaxN, "spd_solve: n = %d outside [1, %d]", n, kCholMaxN);
	TimerScope ts(ctx, "spd_solve");
	PNOL_CHECK(ws_reserve(ctx, 1, (size_t) n * n * sizeof(double)));
	size_t smem = ((size_t) kCholNB * (kCholNB + 1) + (size_t) (n > kCholNB ? n - kCholNB : 0) * (kCholNB + 1) + n) * sizeof(double);
	PNOL_REQUIRE(ctx, smem <= ctx->smem_optin, "spd_solve: n = %d needs %zu bytes of shared memory", n, smem);
	PNOL_CUDA(ctx, cudaFuncSetAttribute(spd_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	PNOL_LAUNCH(ctx, spd_solve_kernel, 1, kCholThreads, smem, A, rhs, n, (double *) ctx->ws[1], x, info_dev);
	return PNOL_OK;
}

// ---------------------------------------------------------------------------------------------------
// a12: p = -D g    (Source/BFGS_bnd_linesearch_MPI_SW.cpp:143-144). One warp per row, coalesced row reads.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
matvec_kernel(const double * __restrict__ D, const double * __restrict__ g, int n, double scale, double * __restrict__ out)
{
	int row = (int) (((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	int lane = threadIdx.x & 31;
	if (row >= n) return;
	const double * Dr = D + (long long) row * n;
	double s = 0;
	for (int k = lane; k < n; k += 32) s = fma(Dr[k], g[k], s);
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if (lane == 0) out[row] = scale * s;
}

int launch_matvec_neg(pnol_ctx * ctx, const double * D, const double * g, int n, double * p)
{
	TimerScope ts(ctx, "matvec_neg");
	long long threads = (long long) n * 32;
	PNOL_LAUNCH(ctx, matvec_kernel, (unsigned) ((threads + 255) / 256), 256, 0, D, g, n, -1.0, p);
	return PNOL_OK;
}

// ---------------------------------------------------------------------------------------------------
// a11: updateHessianInv(D, g, s)   (Source/BFGS_with_linesearch.cpp:389-432)
// ---------------------------------------------------------------------------------------------------
// rank-2 form. pass 1: u = D g (row sums), v = D^T g (column sums), both from ONE read of D.
constexpr int kR2Rows = 32;     // rows per block in pass 1
constexpr int kR2Cols = 512;    // columns per block in pass 1 (8 warps x 512 x 8 B = 32 KB static smem)

__global__ void __launch_bounds__(256)
hinv_pass1_kernel(const double * __restrict__ D, const double * __restrict__ g, int n, double * __restrict__ upart /* gridDim.y x n */,
                  double * __restrict__ vpart /* gridDim.x x n */)
{
	// block = 8 warps, rows [r0, r0 + 32) x columns [c0, c0 + 1024): warp w handles rows r0 + w, r0 + w + 8, ...;
	// column partials are kept per warp in shared memory and combined at the end (fixed order).
	__shared__ double sm[8 * kR2Cols];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int r0 = blockIdx.x * kR2Rows;
	const int c0 = blockIdx.y * kR2Cols;
	const int c1 = min(n, c0 + kR2Cols);
	double * vcol = sm + warp * kR2Cols;
	for (int k = lane; k < kR2Cols; k += 32) vcol[k] = 0;
	__syncwarp();
	for (int rr = warp; rr < kR2Rows; rr += 8) {
		int row = r0 + rr;
		if (row >= n) break;
		const double * Dr = D + (long long) row * n;
		const double gi = g[row];
		double s = 0;
		for (int k = c0 + lane; k < c1; k += 32) {
			double d = Dr[k];
			s = fma(d, g[k], s);
			vcol[k - c0] = fma(gi, d, vcol[k - c0]);
		}
		for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
		if (lane == 0) upart[(size_t) blockIdx.y * n + row] = s;
	}
	__syncthreads();
	for (int k = threadIdx.x; k < c1 - c0; k += blockDim.x) {
		double s = 0;
		for (int w = 0; w < 8; w++) s = s + sm[w * kR2Cols + k];
		vpart[(size_t) blockIdx.x * n + c0 + k] = s;
	}
}

// out[k] = sum over parts of part[b][k] (fixed order)
__global__ void __launch_bounds__(256)
hinv_pass1_finish_kernel(const double * __restrict__ part, int nparts, int n, double * __restrict__ out)
{
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	double s = 0;
	for (int b = 0; b < nparts; b++) s = s + part[(size_t) b * n + k];
	out[k] = s;
}

// scal[0] = g.s , scal[1] = g.u   (single block, fixed-order tree)
__global__ void __launch_bounds__(1024)
dot2_kernel(const double * __restrict__ g, const double * __restrict__ s, const double * __restrict__ u, int n,
            double * __restrict__ scal)
{
	__shared__ double r0[1024], r1[1024];
	double a = 0, b = 0;
	for (int k = threadIdx.x; k < n; k += 1024) { a = fma(g[k], s[k], a); if (u) b = fma(g[k], u[k], b); }
	r0[threadIdx.x] = a; r1[threadIdx.x] = b;
	__syncthreads();
	for (int o = 512; o > 0; o >>= 1) {
		if (threadIdx.x < o) { r0[threadIdx.x] += r0[threadIdx.x + o]; r1[threadIdx.x] += r1[threadIdx.x + o]; }
		__syncthreads();
	}
	if (threadIdx.x == 0) { scal[0] = r0[0]; scal[1] = r1[0]; }
}

// pass 2: D_ij += -rho s_i v_j - rho u_i s_j + (rho^2 gamma + rho) s_i s_j
__global__ void __launch_bounds__(256)
hinv_pass2_kernel(double * __restrict__ D, const double * __restrict__ s, const double * __restrict__ u,
                  const double * __restrict__ v, const double * __restrict__ scal, int n)
{
	const double rho = 1.0 / scal[0];
	const double cc = rho * rho * scal[1] + rho;
	long long idx2 = (long long) blockIdx.x * blockDim.x + threadIdx.x;     // index of a double2
	long long total2 = (long long) n * n / 2;
	for (; idx2 < total2; idx2 += (long long) gridDim.x * blockDim.x) {
		long long e = idx2 * 2;
		int i = (int) (e / n), j = (int) (e - (long long) i * n);
		double2 d = reinterpret_cast<double2 *>(D)[idx2];
		double si = s[i], ui = u[i];
		d.x = d.x + (-rho * si * v[j] - rho * ui * s[j] + cc * si * s[j]);
		d.y = d.y + (-rho * si * v[j + 1] - rho * ui * s[j + 1] + cc * si * s[j + 1]);
		reinterpret_cast<double2 *>(D)[idx2] = d;
	}
}
__global__ void __launch_bounds__(256)
hinv_pass2_scalar_kernel(double * __restrict__ D, const double * __restrict__ s, const double * __restrict__ u,
                         const double * __restrict__ v, const double * __restrict__ scal, int n)
{
	const double rho = 1.0 / scal[0];
	const double cc = rho * rho * scal[1] + rho;
	long long total = (long long) n * n;
	for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long) gridDim.x * blockDim.x) {
		int i = (int) (e / n), j = (int) (e - (long long) i * n);
		D[e] = D[e] + (-rho * s[i] * v[j] - rho * u[i] * s[j] + cc * s[i] * s[j]);
	}
}

int launch_hinv_rank2(pnol_ctx * ctx, double * D, const double * g, const double * s, int n)
{
	TimerScope ts(ctx, "hinv_rank2");
	int nrb = (n + kR2Rows - 1) / kR2Rows;
	int ncb = (n + kR2Cols - 1) / kR2Cols;
	size_t need = ((size_t) 2 * n + 2 + (size_t) nrb * n + (size_t) ncb * n) * sizeof(double);
	PNOL_CHECK(ws_reserve(ctx, 1, need));
	double * u = (double *) ctx->ws[1];
	double * v = u + n;
	double * scal = v + n;
	double * vpart = scal + 2;
	double * upart = vpart + (size_t) nrb * n;
	PNOL_LAUNCH(ctx, hinv_pass1_kernel, dim3(nrb, ncb), 256, 0, D, g, n, upart, vpart);
	PNOL_LAUNCH(ctx, hinv_pass1_finish_kernel, (n + 255) / 256, 256, 0, vpart, nrb, n, v);
	PNOL_LAUNCH(ctx, hinv_pass1_finish_kernel, (n + 255) / 256, 256, 0, upart, ncb, n, u);
	PNOL_LAUNCH(ctx, dot2_kernel, 1, 1024, 0, g, s, u, n, scal);
	int grid = ctx->sm_count * 8;
	if (n % 2 == 0 && (((size_t) D) & 15) == 0)
		PNOL_LAUNCH(ctx, hinv_pass2_kernel, grid, 256, 0, D, s, u, v, scal, n);
	else
		PNOL_LAUNCH(ctx, hinv_pass2_scalar_kernel, grid, 256, 0, D, s, u, v, scal, n);
	return PNOL_OK;
}

// literal form: M1 = I - rho s g^T, M2 = I - rho g s^T, A = M1 D, D = A M2, D += rho s s^T
__global__ void __launch_bounds__(256)
hinv_build_m_kernel(const double * __restrict__ g, const double * __restrict__ s, const double * __restrict__ scal, int n,
                    double * __restrict__ M1, double * __restrict__ M2)
{
	const double rho = 1.0 / scal[0];
	long long total = (long long) n * n;
	for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long) gridDim.x * blockDim.x) {
		int i = (int) (e / n), j = (int) (e - (long long) i * n);
		double id = (i == j) ? 1.0 : 0.0;
		M1[e] = id - rho * s[i] * g[j];      // (Source/BFGS_with_linesearch.cpp:412)
		M2[e] = id - rho * g[i] * s[j];      // (:413)
	}
}
__global__ void __launch_bounds__(256)
hinv_add_m3_kernel(double * __restrict__ D, const double * __restrict__ s, const double * __restrict__ scal, int n)
{
	const double rho = 1.0 / scal[0];
	long long total = (long long) n * n;
	for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long) gridDim.x * blockDim.x) {
		int i = (int) (e / n), j = (int) (e - (long long) i * n);
		D[e] = D[e] + rho * s[i] * s[j];     // (:414, :428)
	}
}

int launch_hinv_literal(pnol_ctx * ctx, double * D, const double * g, const double * s, int n)
{
	TimerScope ts(ctx, "hinv_literal");
	size_t nn = (size_t) n * n;
	PNOL_CHECK(ws_reserve(ctx, 1, (3 * nn + 2) * sizeof(double)));
	double * M1 = (double *) ctx->ws[1];
	double * M2 = M1 + nn;
	double * A = M2 + nn;
	double * scal = A + nn;
	PNOL_LAUNCH(ctx, dot2_kernel, 1, 1024, 0, g, s, (const double *) nullptr, n, scal);
	int grid = ctx->sm_count * 8;
	PNOL_LAUNCH(ctx, hinv_build_m_kernel, grid, 256, 0, g, s, scal, n, M1, M2);
	PNOL_CHECK(launch_dgemm_nn(ctx, M1, D, A, n, n, n));
	PNOL_CHECK(launch_dgemm_nn(ctx, A, M2, D, n, n, n));
	PNOL_LAUNCH(ctx, hinv_add_m3_kernel, grid, 256, 0, D, s, scal, n);
	return PNOL_OK;
}

} // namespace pnol
