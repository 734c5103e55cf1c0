// dense.cu -- the small dense FP64 pieces around the DMMA kernels: Marquardt damping epilogue, Cholesky solve,
// search-direction matvec and the BFGS inverse-Hessian update (rank-2 and literal forms).
#include "common.cuh"

namespace pnol {

// 256-bit global loads (one full 32-byte sector per lane and load). D is streamed (134 MB at n = 4096: larger than L2), so it
// bypasses L1; the vectors (g, s, v: 32 KB each) are re-read by every warp and stay in L1.
__device__ __forceinline__ void ldg256_stream(const double * p, double (&v)[4])
{
	asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ldg256_cached(const double * p, double (&v)[4])
{
	asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}


// ---------------------------------------------------------------------------------------------------
// a9: A = JTJ, A_ii = (1 + lambda) JTJ_ii ; rhs = -J^T F     (Source/LevenbergMarquardtMPI.cpp:66-85)
// ---------------------------------------------------------------------------------------------------
__global__ void lm_damp_kernel(const double * __restrict__ packed, int n, double lambda, const double * __restrict__ lambda_dev,
                               double * __restrict__ JTJ, double * __restrict__ A, double * __restrict__ rhs, int jtj_tail)
{
	long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	long long total = (long long) n * n;
	if (lambda_dev) lambda = *lambda_dev;           // device-resident LM loop (pnol_lm_iterate): lambda never visits the host
	if (idx < total) {
		double v = packed[idx];
		if (JTJ) JTJ[idx] = v;
		if (A) {
			int i = (int) (idx / n), j = (int) (idx - (long long) i * n);
			A[idx] = (i == j) ? (1 + lambda) * v : v;
		}
	} else if (idx < total + n) {
		const double v = -packed[idx];
		if (rhs) rhs[idx - total] = v;
		if (jtj_tail && JTJ) JTJ[idx] = v;      // the LM step keeps the right-hand side behind J^T J (a re-damped step reuses it)
	}
}

int launch_lm_damp(pnol_ctx * ctx, const double * packed, int n, double lambda, double * JTJ, double * A, double * rhs,
                   const double * lambda_dev, bool rhs_behind_jtj)
{
	long long total = (long long) n * n + n;
	PNOL_LAUNCH(ctx, lm_damp_kernel, (unsigned) ((total + 255) / 256), 256, 0, packed, n, lambda, lambda_dev, JTJ, A, rhs, rhs_behind_jtj ? 1 : 0);
	return PNOL_OK;
}

// ---------------------------------------------------------------------------------------------------
// a9: SPD solve  A x = b  by right-looking blocked Cholesky on ONE thread-block cluster of 8 CTAs.
// Replaces luSolve (Source/LevenbergMarquardtMPI.cpp:88).
//
// The work is 5.6 MFLOP at n = 256 -- nothing; the cost is the length of the dependency chain and the number of L2 round trips
// on it. History: one CTA 0.56 ms; 8-CTA cluster with a shuffle Cholesky of the diagonal block 0.183 ms, of which (phase
// timestamps, tools/solve_stamps.py) 65 us were the eight 32 x 32 diagonal factorisations (485 cycles per pivot: MUFU.RSQ64H,
// three Newton steps, a correction of the root and 31 64-bit shuffles, all behind one another), 31 us trailing updates and 13 us
// the copy into the work matrix (both waiting for L2 once per loop trip), 26 us the back substitution (ditto). Now:
//   * the matrix W is (n+1) x n in global memory (L2 resident): rows 0..n-1 the lower triangle of A, row n = b. Carrying b
//     as an extra row makes the forward substitution L y = b fall out of the factorisation (the panel solve of that row IS
//     y_k = L_kk^-1 (b_k - ...), its trailing update IS the forward-substitution update), so only L^T x = y remains;
//   * per 32-column block step every CTA factors the 32 x 32 diagonal block redundantly (one warp, one matrix row per lane in
//     REGISTERS) as L' D L'^T: the chain per pivot is one broadcast, one reciprocal (MUFU.RCP64H + 5 FMA) and two FMA; the
//     column values the update needs are shuffled BEFORE the reciprocal is known. 1/sqrt(d) is taken for all 32 pivots at once
//     at the end (L = L' sqrt(D));
//   * panel rows: one thread per row, right-looking unit-triangular substitution (one FMA per step on the chain) with L' read as
//     shared-memory broadcasts; trailing update: one warp per row (fixed owner: row mod 64), lanes along the columns, the whole
//     panel staged in shared memory; every global load of a phase is issued before its first use (256-bit where aligned);
//   * two cluster barriers per step order the phases;
//   * CTA 0 finishes with the blocked back substitution; the block row of L the next step needs is copied into shared memory
//     (cp.async) while the current step computes.
// ---------------------------------------------------------------------------------------------------
constexpr int kCholNB = 32;
constexpr int kCholThreads = 256;
constexpr int kCholCluster = 8;
constexpr int kCholMaxN = 704;

__device__ __forceinline__ unsigned cluster_ctarank()
{
	unsigned r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
	// release/acquire at cluster scope: global-memory writes of one phase are visible to every CTA of the cluster in the next
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// 1/sqrt(d) for a positive normal d without branches: MUFU.RSQ64H seed and three Newton steps (quadratic convergence from
// >= 20 good bits: the last step only polishes), accurate to about an ulp. The factor needs no more: it is compared with the
// reference's LU solve at 1e-9 and never bit for bit (the reference's luSolve lives in an un-vendored library).
__device__ __forceinline__ double rsqrt_nr(double d)
{
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#pragma unroll
	for (int it = 0; it < 3; it++) {
		const double e = fma(-d * y, y, 1.0);        // 1 - d y^2
		y = fma(y, 0.5 * e, y);
	}
	return y;
}

// 1/d for a positive normal d: MUFU.RCP64H seed, one cubic and one quadratic step (the sequence nvcc itself emits in front of a
// division), five dependent FMA
__device__ __forceinline__ double rcp_nr(double d)
{
	double y;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
	double e = fma(-d, y, 1.0);
	e = fma(e, e, e);
	y = fma(y, e, y);
	e = fma(-d, y, 1.0);
	return fma(y, e, y);
}

// L' D L'^T of a 32 x 32 block by one warp: lane r holds row r (lower part) in registers. On return row[c], c < r, holds L'[r][c]
// (unit lower triangle) and d_mine the pivot d_r of this lane; bad_col is the first non-positive pivot (1-based inside the
// block, 0 = none; the arithmetic goes on with it, the caller discards the result). Rows / columns past the block's real size
// must have been padded with the identity. Entries above the diagonal (row[c], c > lane) take meaningless updates and are
// never read by a valid lane.
// Column J's values a'[.][J] travel through row J of a 32 x 32 shared-memory array (one STS, one __syncwarp, broadcast LDS)
// rather than through shuffles: the warp runs under `if (warp == 0)`, where the compiler cannot prove convergence and brackets
// every __shfl_sync with WARPSYNC.COLLECTIVE / ENDCOLLECTIVE (measured: 485 cycles per pivot with 31 shuffles per column; all
// eight warps factoring redundantly at the top level, plain SHFL: 326 cycles, issue-bound).
// Two levels: inside a group of 8 columns a pivot is followed only by the updates of the group's own columns, the next column
// first and out to shared memory at once (the chain: LDS, MUFU.RCP64H + 5 FMA, one multiply, one FMA, STS); the columns right of
// the group take the group's eight updates in one bulk pass. (With all 31 - J updates behind every pivot ptxas put the
// store that the next pivot waits for behind them: 330 cycles per pivot.)
// (template recursion: a doubly nested `#pragma unroll` is only partially honoured at 496 bodies, and a dynamic index would
// push row[] into local memory)
constexpr int kLdlGroup = 8;

template <int J> struct LdlColumn {
	// on entry row J of bc holds column J (a'[r][J] of every lane r), published by the previous column / bulk pass / caller.
	// A column is published twice, at bc[J][r] and, shifted by one, at bs[J][r + 1]: the pivot a'[J][J] and a'[J + 1][J], which the
	// next pivot waits for, then sit in one aligned 16-byte pair for even J in the first and for odd J in the second copy
	// (one LDS.128 at the head of the chain instead of an LDS.64 behind the reciprocal).
	__device__ __forceinline__ static void publish(double * bc, int lane, double v)
	{
		bc[J * kCholNB + lane] = v;
		bc[kCholNB * kCholNB + J * (kCholNB + 2) + lane + 1] = v;
		__syncwarp();
	}
	__device__ __forceinline__ static void run(double (&row)[kCholNB], int lane, int & bad_col, double & d_mine, double * bc)
	{
		constexpr int kEnd = (J / kLdlGroup + 1) * kLdlGroup;      // first column right of J's group
		const double * col = bc + J * kCholNB;
		const double2 head = (J & 1) ? *reinterpret_cast<const double2 *>(bc + kCholNB * kCholNB + J * (kCholNB + 2) + J + 1)
		                             : *reinterpret_cast<const double2 *>(col + J);
		const double d = head.x;                                   // the pivot
		if (!(d > 0.0 && d < 0x1p1000) && bad_col == 0) bad_col = J + 1;
		if (lane == J) d_mine = d;
		const double l = row[J] * rcp_nr(d);                       // L'[r][J] (meaningful for lane > J)
		if (J + 1 < kEnd) {
			constexpr int N = J + 1 < kCholNB ? J + 1 : J;
			row[N] = fma(-l, head.y, row[N]);
			LdlColumn<N>::publish(bc, lane, row[N]);
		}
#pragma unroll
		for (int c = J + 2; c < kEnd; c++) row[c] = fma(-l, col[c], row[c]);
		row[J] = l;
		if (J + 1 == kEnd && kEnd < kCholNB) {
			// bulk pass: the columns right of the group take its kLdlGroup updates (independent chains of 8 FMA)
#pragma unroll
			for (int c = kEnd; c < kCholNB; c++) {
#pragma unroll
				for (int t = kEnd - kLdlGroup; t < kEnd; t++) row[c] = fma(-row[t], bc[t * kCholNB + c], row[c]);
				if (c == kEnd) LdlColumn<(kEnd < kCholNB ? kEnd : 0)>::publish(bc, lane, row[c]);
			}
		}
		LdlColumn<J + 1>::run(row, lane, bad_col, d_mine, bc);
	}
};
template <> struct LdlColumn<kCholNB> {
	__device__ __forceinline__ static void run(double (&)[kCholNB], int, int &, double &, double *) {}
};

__device__ __forceinline__ void dmma_8x8x4_acc(double & d0, double & d1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// tuning aid (-DPNOL_SOLVE_STAMPS, tools/build_variants.sh): CTA 0 / thread 0 writes %globaltimer at the phase boundaries into x
#ifdef PNOL_SOLVE_STAMPS
#define SOLVE_STAMP() do { if (cta == 0 && tid == 0 && nstamp < 200) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); stamps[nstamp++] = t_; } } while (0)
#else
#define SOLVE_STAMP() do { } while (0)
#endif

__device__ __forceinline__ void ld256(const double * p, double (&v)[4])       // coherent (not .nc): W is rewritten between phases
{
	asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void cp_async16(void * smem_dst, const void * gsrc)
{
	const unsigned d = (unsigned) __cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

constexpr int kBsP = kCholNB + 2;      // pitch of the staged diagonal block in the back substitution (rows 16-byte aligned)

// back substitution: stage block row kb of L (diagonal block -> Dd, pitch kBsP; rows kb..kb+nbk-1, columns [0, kb) -> Pd, pitch `pitch`)
__device__ __forceinline__ void backsub_stage(const double * W, int n, int kb, int nbk, double * Dd, double * Pd, int pitch, bool async_ok, int tid)
{
	if (async_ok && nbk == kCholNB) {
		for (int e = tid; e < kCholNB * (kCholNB / 2); e += kCholThreads) {
			const int r = e >> 4, c2 = e & 15;
			cp_async16(Dd + r * kBsP + 2 * c2, W + (long long) (kb + r) * n + kb + 2 * c2);
		}
		const int per_row = kb >> 1;                            // 16-byte pieces per row (kb is a multiple of 32)
		for (int e = tid; e < nbk * per_row; e += kCholThreads) {
			const int t = e / per_row, c2 = e - t * per_row;
			cp_async16(Pd + (size_t) t * pitch + 2 * c2, W + (long long) (kb + t) * n + 2 * c2);
		}
	} else {
		for (int e = tid; e < nbk * kCholNB; e += kCholThreads) {
			const int r = e >> 5, c = e & 31;
			if (c <= r) Dd[r * kBsP + c] = W[(long long) (kb + r) * n + kb + c];
		}
		for (int e = tid; e < nbk * kb; e += kCholThreads) {
			const int t = e / kb, c = e - t * kb;
			Pd[(size_t) t * pitch + c] = W[(long long) (kb + t) * n + c];
		}
	}
}

constexpr int kPnP = kCholNB + 4;      // pitch of the staged panel: 8 rows x 4 doubles of a DMMA fragment load hit 32 distinct bank pairs

__global__ void __cluster_dims__(kCholCluster, 1, 1) __launch_bounds__(kCholThreads, 1)
spd_solve_kernel(const double * __restrict__ A, const double * __restrict__ rhs, int n, double * __restrict__ W,
                 double * __restrict__ x, int * __restrict__ info, int nbuf,
                 const double * __restrict__ xbase, double * __restrict__ xtrial, double * __restrict__ step_out)
{
	extern __shared__ double sm[];
	__shared__ int s_bad;
	constexpr int P = kCholNB + 1;
	double * LpT = sm;                              // 32 x 32: LpT[t * 32 + c] = L'[c][t], c > t (unit lower factor of the diagonal block)
	double * Rsd = sm + kCholNB * kCholNB;          // 32: 1 / sqrt(d)
	double * Bc = Rsd + kCholNB;                    // 32 x 32 + 32 x 34: the columns of the diagonal factorisation and their shifted copy (LdlColumn)
	double * Pn = Bc + kCholNB * kCholNB + kCholNB * (kCholNB + 2);           // up to round8(n + 1 - 32) x kPnP: the panel below it (incl. the b row)
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	// the grid is exactly one cluster: blockIdx.x IS the CTA's rank in it
	const int cta = (int) blockIdx.x;
	const int gwarp = warp * kCholCluster + cta, gwarps = (kCholThreads / 32) * kCholCluster;
	const bool vec = (n & 3) == 0 && ((((size_t) A) | ((size_t) W) | ((size_t) rhs)) & 31) == 0;
	int bad = 0;
#ifdef PNOL_SOLVE_STAMPS
	__shared__ unsigned long long stamps[200];
	int nstamp = 0;
#endif
	SOLVE_STAMP();

	// The work matrix W (global, L2-resident) is (n + 1) x n: rows 0..n-1 the lower triangle, row n the right-hand side. Step 0
	// reads A and rhs themselves (no copy pass): every entry of W that a later step reads has been written by then.
	for (int kb = 0; kb < n; kb += kCholNB) {
		const int nbk = min(kCholNB, n - kb);
		const int r0 = kb + nbk;                    // first row below the diagonal block
		const int below = n + 1 - r0;               // rows below, b row included
		const int cbelow = n - r0;                  // columns to the right
		const bool first = kb == 0;
		const bool vblk = vec && nbk == kCholNB;
		auto rowsrc = [&](int i) -> const double * { return first ? (i < n ? A + (long long) i * n : rhs) : W + (long long) i * n; };

		// ---- panel rows: warps 1..7, one thread per row; their loads are on the way while warp 0 factors the diagonal block ----
		const int pri = tid >= 32 ? (tid - 32) * kCholCluster + cta : -1;
		double v[kCholNB];
		static_assert(kCholMaxN + 1 <= (kCholThreads - 32) * kCholCluster, "one panel row per thread of warps 1..7");
		if (pri >= 0 && pri < below) {
			const double * wsrc = rowsrc(r0 + pri) + kb;
			if (vblk) {
#pragma unroll
				for (int c4 = 0; c4 < kCholNB / 4; c4++) {
					double t4[4];
					ld256(wsrc + 4 * c4, t4);
#pragma unroll
					for (int e = 0; e < 4; e++) v[4 * c4 + e] = t4[e];
				}
			} else {
#pragma unroll
				for (int c = 0; c < kCholNB; c++) v[c] = c < nbk ? wsrc[c] : 0.0;
			}
		}
		// ---- A: diagonal block, factored by warp 0 of EVERY CTA (no cluster barrier needed before the panel solve) ----
		if (warp == 0) {
			double row[kCholNB];
			const double * wr = rowsrc(kb + min(lane, nbk - 1)) + kb;
			if (vblk) {
#pragma unroll
				for (int c4 = 0; c4 < kCholNB / 4; c4++) {
					double t4[4];
					ld256(wr + 4 * c4, t4);
#pragma unroll
					for (int e = 0; e < 4; e++) row[4 * c4 + e] = (4 * c4 + e <= lane) ? t4[e] : 0.0;
				}
			} else {
#pragma unroll
				for (int c = 0; c < kCholNB; c++)
					row[c] = (lane < nbk && c < nbk) ? (c <= lane ? wr[c] : 0.0) : (c == lane ? 1.0 : 0.0);
			}
			int bc = 0;
			double dm = 1.0;
#ifdef PNOL_SOLVE_STAMPS
			if (row[0] == 123.456) bad = 7;      // the loads have landed
			SOLVE_STAMP();
#endif
			LdlColumn<0>::publish(Bc, lane, row[0]);
			LdlColumn<0>::run(row, lane, bc, dm, Bc);
#ifdef PNOL_SOLVE_STAMPS
			if (row[31] == 123.456) bad = 7;
			SOLVE_STAMP();
#endif
			if (bc != 0 && bc <= nbk && bad == 0) bad = kb + bc;
#pragma unroll
			for (int c = 0; c < kCholNB; c++) LpT[c * kCholNB + lane] = (c < lane) ? row[c] : 0.0;
			Rsd[lane] = rsqrt_nr(dm);
		}
		__syncthreads();
		SOLVE_STAMP();
		// ---- B: panel solve  X L_kk^T = W[row][kb .. kb+nbk),  L_kk = L' sqrt(D) ----
		if (pri >= 0 && pri < below) {
			// right-looking: once x_t is final every later entry takes its contribution (one FMA on the chain per step)
#pragma unroll
			for (int t = 0; t < kCholNB - 1; t++) {
				const double xt = v[t];
#pragma unroll
				for (int c = t + 1; c < kCholNB; c++) v[c] = fma(-xt, LpT[t * kCholNB + c], v[c]);
			}
#pragma unroll
			for (int c = 0; c < kCholNB; c++) v[c] = v[c] * Rsd[c];
			double * wrow = W + (long long) (r0 + pri) * n + kb;
			if (vblk) {
#pragma unroll
				for (int c4 = 0; c4 < kCholNB / 4; c4++)
					*reinterpret_cast<double4 *>(wrow + 4 * c4) = make_double4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
			} else {
#pragma unroll
				for (int c = 0; c < kCholNB; c++)
					if (c < nbk) wrow[c] = v[c];
			}
		}
		SOLVE_STAMP();
		cluster_sync_all();
		SOLVE_STAMP();
		// CTA 0 stores the factor of the diagonal block for the back substitution (strict lower triangle: L'; diagonal: 1 / sqrt(d))
		if (cta == 0) {
			for (int e = tid; e < nbk * kCholNB; e += kCholThreads) {
				const int r = e >> 5, c = e & 31;
				if (c < r) W[(long long) (kb + r) * n + kb + c] = LpT[c * kCholNB + r];
				else if (c == r) W[(long long) (kb + r) * n + kb + c] = Rsd[r];
			}
		}
		// ---- C: trailing update  W[i][j] -= sum_t L[i][t] L[j][t]  for r0 <= j <= i (i runs over the b row too), on the FP64
		// tensor cores: an item is 8 rows x 32 columns (four m8n8k4 accumulators, the A fragment shared), items dealt round-robin to
		// the 64 warps of the cluster. The old values of a warp's items are requested before the panel is staged / while the previous
		// item computes, so that the L2 round trips overlap. (One thread per entry with DFMA: 32 LDS.64 per 32 FMA, shared-memory
		// bound -- 52 us of a 140 us solve.)
		const int TA = (below + 7) >> 3;                        // 8-row tiles
		const int Gf = TA >> 2;
		const int n_items = 2 * Gf * (Gf + 1) + (TA & 3) * (Gf + 1);
		const int fr = lane >> 2, fk = lane & 3;                // fragment row / k index (A, B); C: row fr, columns 2 fk, 2 fk + 1
		auto decode = [&](int id, int & ta, int & c4) {
			int g = 0;
			while (2 * (g + 1) * (g + 2) <= id) g++;
			const int rem = id - 2 * g * (g + 1);
			ta = 4 * g + rem / (g + 1);
			c4 = rem - (rem / (g + 1)) * (g + 1);
		};
		auto load_c = [&](int id, double (&c)[4][2]) {
			int ta, c4;
			decode(id, ta, c4);
			const int ri = 8 * ta + fr;
			const int jmax = min(ri, cbelow - 1);
			const double * wsrc = rowsrc(r0 + min(ri, below - 1)) + r0;
#pragma unroll
			for (int nt = 0; nt < 4; nt++) {
				const int j = 32 * c4 + 8 * nt + 2 * fk;
				c[nt][0] = c[nt][1] = 0.0;
				if (ri < below && j <= jmax) {
					if (vec && j + 1 <= jmax) {
						const double2 t2 = *reinterpret_cast<const double2 *>(wsrc + j);
						c[nt][0] = t2.x; c[nt][1] = t2.y;
					} else {
						c[nt][0] = wsrc[j];
						if (j + 1 <= jmax) c[nt][1] = wsrc[j + 1];
					}
				}
			}
		};
		double ccur[4][2], cnxt[4][2];
		int item = gwarp;
		if (item < n_items) load_c(item, ccur);
		if (vblk) {
			const int pieces = below * 8;               // 32-byte pieces of the panel
			for (int e0 = tid; e0 < pieces; e0 += kCholThreads * 8) {
				double t8[8][4];
#pragma unroll
				for (int q = 0; q < 8; q++) {
					const int e = e0 + q * kCholThreads;
					if (e < pieces) ld256(W + (long long) (r0 + (e >> 3)) * n + kb + 4 * (e & 7), t8[q]);
				}
#pragma unroll
				for (int q = 0; q < 8; q++) {
					const int e = e0 + q * kCholThreads;
					if (e < pieces)
						*reinterpret_cast<double4 *>(Pn + (e >> 3) * kPnP + 4 * (e & 7)) = make_double4(t8[q][0], t8[q][1], t8[q][2], t8[q][3]);
				}
			}
		} else {
			for (int e = tid; e < below * kCholNB; e += kCholThreads) {
				const int ri = e >> 5, c = e & 31;
				Pn[ri * kPnP + c] = c < nbk ? W[(long long) (r0 + ri) * n + kb + c] : 0.0;
			}
		}
		__syncthreads();
		for (; item < n_items; item += gwarps) {
			const bool more = item + gwarps < n_items;
			if (more) load_c(item + gwarps, cnxt);
			int ta, c4;
			decode(item, ta, c4);
			const int ri = 8 * ta + fr;
			const double * pa = Pn + ri * kPnP + fk;
			const double * pb = Pn + (32 * c4 + fr) * kPnP + fk;
			const int ntmax = min(3, (8 * ta + 7 - 32 * c4) >> 3);       // column tiles at or below the diagonal (warp-uniform)
#pragma unroll
			for (int t0 = 0; t0 < kCholNB; t0 += 4) {
				const double av = -pa[t0];
#pragma unroll
				for (int nt = 0; nt < 4; nt++)
					if (nt <= ntmax) dmma_8x8x4_acc(ccur[nt][0], ccur[nt][1], av, pb[nt * 8 * kPnP + t0]);
			}
			const int jmax = min(ri, cbelow - 1);
			double * wrow = W + (long long) (r0 + ri) * n + r0;
#pragma unroll
			for (int nt = 0; nt < 4; nt++) {
				const int j = 32 * c4 + 8 * nt + 2 * fk;
				if (ri < below && j <= jmax) {
					if (vec && j + 1 <= jmax) *reinterpret_cast<double2 *>(wrow + j) = make_double2(ccur[nt][0], ccur[nt][1]);
					else {
						wrow[j] = ccur[nt][0];
						if (j + 1 <= jmax) wrow[j + 1] = ccur[nt][1];
					}
				}
			}
			if (more) {
#pragma unroll
				for (int nt = 0; nt < 4; nt++) { ccur[nt][0] = cnxt[nt][0]; ccur[nt][1] = cnxt[nt][1]; }
			}
		}
		SOLVE_STAMP();
		cluster_sync_all();
		SOLVE_STAMP();
	}

	// ---- back substitution L^T x = y (y = row n of W), CTA 0.  L_kk = L' sqrt(D): L'^T x_k = (y_k - ...) / sqrt(d) ----
	if (cta == 0) {
		// shared memory: (32 unused) | yv (n) | nbuf x { Dd 32 x kBsP | Pd 32 x pitch }
		const int last = ((n - 1) / kCholNB) * kCholNB;
		const int pitch = last > 0 ? last : 2;
		double * yv = sm + kCholNB;
		const size_t buf_doubles = (size_t) kCholNB * kBsP + (size_t) kCholNB * pitch;
		double * buf0 = yv + (((size_t) n + 1) & ~(size_t) 1);
		const bool async_ok = (n & 1) == 0 && (((size_t) W) & 15) == 0;
		__syncthreads();
		for (int i = tid; i < n; i += kCholThreads) yv[i] = W[(long long) n * n + i];
		backsub_stage(W, n, last, n - last, buf0, buf0 + kCholNB * kBsP, pitch, async_ok, tid);
		int cur = 0;
		for (int kb = last; kb >= 0; kb -= kCholNB) {
			const int nbk = min(kCholNB, n - kb);
			double * Dd = buf0 + (size_t) cur * buf_doubles;
			double * Pd = Dd + kCholNB * kBsP;
			cp_async_wait_all();
			__syncthreads();
			// next block row on its way while this one computes
			if (nbuf == 2 && kb > 0) {
				double * Dn = buf0 + (size_t) (cur ^ 1) * buf_doubles;
				backsub_stage(W, n, kb - kCholNB, kCholNB, Dn, Dn + kCholNB * kBsP, pitch, async_ok, tid);
			}
			{
				// lane r holds z_r = y_r / sqrt(d_r); unit upper-triangular solve from the last unknown up. EVERY warp does it (the
				// branch around this block is uniform for the compiler -- blockIdx -- so the shuffles are plain SHFL, 24 cycles per
				// step; under `if (warp == 0)` each would be a WARPSYNC.COLLECTIVE bracket); warp 0 stores
				double z = lane < nbk ? yv[kb + lane] * Dd[lane * kBsP + lane] : 0.0;
				double lk[kCholNB];
#pragma unroll
				for (int k = 1; k < kCholNB; k++) lk[k] = (k < nbk && lane < k) ? Dd[k * kBsP + lane] : 0.0;
#pragma unroll
				for (int k = kCholNB - 1; k >= 1; k--) z = fma(-lk[k], __shfl_sync(0xffffffffu, z, k), z);
				__syncthreads();           // every warp has read y_k
				if (warp == 0 && lane < nbk) yv[kb + lane] = z;
			}
			__syncthreads();
			for (int i = tid; i < kb; i += kCholThreads) {
				double s0 = yv[i], s1 = 0, s2 = 0, s3 = 0;
				if (nbk == kCholNB) {
#pragma unroll
					for (int t = 0; t < kCholNB; t += 4) {
						s0 = fma(-Pd[(size_t) t * pitch + i], yv[kb + t], s0);
						s1 = fma(-Pd[(size_t) (t + 1) * pitch + i], yv[kb + t + 1], s1);
						s2 = fma(-Pd[(size_t) (t + 2) * pitch + i], yv[kb + t + 2], s2);
						s3 = fma(-Pd[(size_t) (t + 3) * pitch + i], yv[kb + t + 3], s3);
					}
				} else {
					for (int t = 0; t < nbk; t++) s0 = fma(-Pd[(size_t) t * pitch + i], yv[kb + t], s0);
				}
				yv[i] = (s0 + s1) + (s2 + s3);
			}
			__syncthreads();
			if (nbuf == 2) cur ^= 1;
			else if (kb > 0) backsub_stage(W, n, kb - kCholNB, kCholNB, Dd, Pd, pitch, async_ok, tid);
		}
		for (int i = tid; i < n; i += kCholThreads) x[i] = yv[i];
		if (xtrial) {
			// the LM trial point rides along (Source/LevenbergMarquardtMPI.cpp:97-100: X[i] = X[i] + sigma[i]). A non-positive pivot:
			// the reference's luSolve would have produced inf / NaN and the step would be rejected by the NaN test of its chi^2 (:110);
			// hand back a NaN step for the same outcome
			if (tid == 0) s_bad = bad;
			__syncthreads();
			const int isbad = s_bad;
			for (int i = tid; i < n; i += kCholThreads) {
				const double s = isbad ? __longlong_as_double(0x7ff8000000000000LL) : yv[i];
				step_out[i] = s;
				xtrial[i] = xbase[i] + s;
			}
		}
#ifdef PNOL_SOLVE_STAMPS
		__syncthreads();
		SOLVE_STAMP();
		if (tid == 0) for (int q = 0; q < nstamp && q < n; q++) x[q] = (double) (stamps[q] - stamps[0]);
#endif
		if (tid == 0) *info = bad;         // warp 0 of every CTA saw the same pivots
	}
}

int launch_spd_solve(pnol_ctx * ctx, const double * A, const double * rhs, int n, double * x, int * info_dev,
                     const double * xbase, double * xtrial, double * step_out)
{
	PNOL_REQUIRE(ctx, n >= 1 && n <= kCholMaxN, "spd_solve: n = %d outside [1, %d]", n, kCholMaxN);
	TimerScope ts(ctx, "spd_solve");
	PNOL_CHECK(ws_reserve(ctx, 1, ((size_t) n + 1) * n * sizeof(double)));
	const int prows = ((n + 1 > kCholNB ? n + 1 - kCholNB : 1) + 7) & ~7;                          // whole 8-row DMMA tiles
	const size_t fact = (size_t) 3 * kCholNB * kCholNB + 3 * kCholNB + (size_t) (prows + 32) * kPnP;   // LpT | Rsd | Bc | Pn (B fragments read up to 31 rows past a tile row)
	const int last = ((n - 1) / kCholNB) * kCholNB;
	const size_t buf = (size_t) kCholNB * kBsP + (size_t) kCholNB * (last > 0 ? last : 2);         // one staged block row of L
	const size_t yv = kCholNB + (((size_t) n + 1) & ~(size_t) 1);
	int nbuf = 2;
	size_t doubles = fact > yv + 2 * buf ? fact : yv + 2 * buf;
	if (doubles * sizeof(double) > ctx->smem_optin) { nbuf = 1; doubles = fact > yv + buf ? fact : yv + buf; }
	const size_t smem = doubles * sizeof(double);
	PNOL_REQUIRE(ctx, smem <= ctx->smem_optin, "spd_solve: n = %d needs %zu bytes of shared memory", n, smem);
	PNOL_CUDA(ctx, cudaFuncSetAttribute(spd_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	if (!(xbase && xtrial && step_out)) { xbase = nullptr; xtrial = nullptr; step_out = nullptr; }
	PNOL_LAUNCH(ctx, spd_solve_kernel, kCholCluster, kCholThreads, smem, A, rhs, n, (double *) ctx->ws[1], x, info_dev, nbuf, xbase, xtrial, step_out);
	return PNOL_OK;
}

// ---------------------------------------------------------------------------------------------------
// General inverse by LU with partial pivoting: the reference inverts the forward-difference Hessian with `matrixInverse`
// (Source/BFGS_bnd_linesearch_MPI_SW.cpp:51-59, BFGS_with_linesearch.cpp:35-41) and carries on with whatever comes back, also
// when the Hessian is indefinite (away from a minimum that is the normal case). matrixInverse lives in the un-vendored utility
// library; oracle/shim defines it as Doolittle LU with partial pivoting (first largest |entry| on ties) followed by one pair of
// substitutions per unit vector. This kernel performs exactly those operations, each entry's updates in the same order
// (k ascending, one rounding per multiply and per subtract, -fmad=false), so the result is the shim's bit for bit.
// One CTA: n is small wherever an FD Hessian is affordable (3 n^2 / 2 objective evaluations).
//   LU  (n x n, global, L2-resident) in/out: A on entry, the factors on exit;  piv (n ints);  Y (n x n scratch);  Ainv (n x n)
//   info: 0, or k+1 when pivot k is exactly zero (the reference would divide by zero and go on with inf / NaN -- so do we)
// ---------------------------------------------------------------------------------------------------
constexpr int kLuThreads = 1024;

__global__ void __launch_bounds__(kLuThreads, 1)
lu_inverse_kernel(double * __restrict__ LU, int n, int * __restrict__ piv, double * __restrict__ Y, double * __restrict__ Ainv,
                  int * __restrict__ info)
{
	extern __shared__ double lu_sm[];
	double * rowk = lu_sm;                                   // n: row k right of the diagonal
	__shared__ double red_v[kLuThreads / 32];
	__shared__ int red_i[kLuThreads / 32];
	__shared__ int s_p;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	int first_zero = 0;
	for (int i = tid; i < n; i += kLuThreads) piv[i] = i;
	__syncthreads();
	for (int k = 0; k < n; k++) {
		// pivot: first row i >= k with the largest |LU[i][k]| (strict > in the reference's scan keeps the first)
		double best = -1.0;
		int bi = n;
		for (int i = k + tid; i < n; i += kLuThreads) {
			const double v = fabs(LU[(size_t) i * n + k]);
			if (v > best) { best = v; bi = i; }              // i ascending per thread: the first of equal values is kept
		}
		// NaN entries never win a `>` comparison in the reference either (best starts at |LU[k][k]|; handled below)
		for (int o = 16; o > 0; o >>= 1) {
			const double ov = __shfl_xor_sync(0xffffffffu, best, o);
			const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
			if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
		}
		if (lane == 0) { red_v[warp] = best; red_i[warp] = bi; }
		__syncthreads();
		if (warp == 0) {
			best = lane < kLuThreads / 32 ? red_v[lane] : -1.0;
			bi = lane < kLuThreads / 32 ? red_i[lane] : n;
			for (int o = 16; o > 0; o >>= 1) {
				const double ov = __shfl_xor_sync(0xffffffffu, best, o);
				const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
				if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
			}
			if (lane == 0) {
				// the reference starts from p = k, best = |LU[k][k]| and moves only on a strictly larger value: a NaN diagonal keeps p = k
				const double dk = fabs(LU[(size_t) k * n + k]);
				s_p = (dk != dk || bi >= n || !(best > dk)) ? k : bi;
			}
		}
		__syncthreads();
		const int p = s_p;
		if (p != k) {
			for (int j = tid; j < n; j += kLuThreads) {
				const double a = LU[(size_t) p * n + j], b = LU[(size_t) k * n + j];
				LU[(size_t) p * n + j] = b; LU[(size_t) k * n + j] = a;
			}
			if (tid == 0) { const int t = piv[p]; piv[p] = piv[k]; piv[k] = t; }
		}
		__syncthreads();
		for (int j = k + tid; j < n; j += kLuThreads) rowk[j] = LU[(size_t) k * n + j];
		__syncthreads();
		const double d = rowk[k];
		if (d == 0.0 && first_zero == 0) first_zero = k + 1;
		// rows below: warp per row, lanes along the columns
		for (int i = k + 1 + warp; i < n; i += kLuThreads / 32) {
			double * r = LU + (size_t) i * n;
			const double l = r[k] / d;
			for (int j = k + 1 + lane; j < n; j += 32) r[j] = r[j] - l * rowk[j];
			__syncwarp();
			if (lane == 0) r[k] = l;
		}
		__syncthreads();
	}
	// columns of the inverse: thread c solves L y = P e_c, U x = y (sums in ascending k, as luBackSub)
	for (int c = tid; c < n; c += kLuThreads) {
		for (int i = 0; i < n; i++) {
			double s = (piv[i] == c) ? 1.0 : 0.0;
			const double * r = LU + (size_t) i * n;
			for (int k = 0; k < i; k++) s = s - r[k] * Y[(size_t) k * n + c];
			Y[(size_t) i * n + c] = s;
		}
		for (int i = n - 1; i >= 0; i--) {
			double s = Y[(size_t) i * n + c];
			const double * r = LU + (size_t) i * n;
			for (int k = i + 1; k < n; k++) s = s - r[k] * Ainv[(size_t) k * n + c];
			Ainv[(size_t) i * n + c] = s / r[i];
		}
	}
	if (tid == 0) *info = first_zero;
}

int launch_lu_inverse(pnol_ctx * ctx, const double * A, int n, double * Ainv, int * info_dev)
{
	PNOL_REQUIRE(ctx, n >= 1 && n <= 8192, "lu_inverse: n = %d outside [1, 8192]", n);
	TimerScope ts(ctx, "lu_inverse");
	const size_t nn = (size_t) n * n;
	PNOL_CHECK(ws_reserve(ctx, 1, (2 * nn + (size_t) n + 8) * sizeof(double)));
	double * LU = (double *) ctx->ws[1];
	double * Y = LU + nn;
	int * piv = (int *) (Y + nn);
	PNOL_CUDA(ctx, cudaMemcpyAsync(LU, A, nn * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
	const size_t smem = (size_t) n * sizeof(double);
	PNOL_CUDA(ctx, cudaFuncSetAttribute(lu_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	PNOL_LAUNCH(ctx, lu_inverse_kernel, 1, kLuThreads, smem, LU, n, piv, Y, Ainv, info_dev);
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

// n a multiple of 4, D 32-byte aligned: one warp per row, every lane keeps kMvUnroll 256-bit loads of the row in flight (4 KB per
// warp and batch) and four independent FMA chains. The first version read one double per lane and load (256 B per warp in
// flight per dependent step) and ran at 0.53 of the HBM roofline at n = 4096; blocks of 64 threads so that the 4096 rows spread
// evenly over the 148 SMs (512 blocks of 8 warps left 4 : 3 blocks per SM).
constexpr int kMvUnroll = 4;
constexpr int kMvThreads = 64;

__global__ void __launch_bounds__(kMvThreads)
matvec_v4_kernel(const double * __restrict__ D, const double * __restrict__ g, int n, double scale, double * __restrict__ out)
{
	const int row = (int) (((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	const int lane = threadIdx.x & 31;
	if (row >= n) return;
	const double * Dr = D + (long long) row * n;
	double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
	for (int k0 = lane * 4; k0 < n; k0 += 128 * kMvUnroll) {
		double d[kMvUnroll][4], gv[kMvUnroll][4];
#pragma unroll
		for (int q = 0; q < kMvUnroll; q++) {
			const int k = k0 + 128 * q;
			if (k < n) ldg256_stream(Dr + k, d[q]);
		}
#pragma unroll
		for (int q = 0; q < kMvUnroll; q++) {
			const int k = k0 + 128 * q;
			if (k < n) ldg256_cached(g + k, gv[q]);
		}
#pragma unroll
		for (int q = 0; q < kMvUnroll; q++) {
			const int k = k0 + 128 * q;
			if (k < n) {
				s0 = fma(d[q][0], gv[q][0], s0); s1 = fma(d[q][1], gv[q][1], s1);
				s2 = fma(d[q][2], gv[q][2], s2); s3 = fma(d[q][3], gv[q][3], s3);
			}
		}
	}
	double s = (s0 + s1) + (s2 + s3);
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if (lane == 0) out[row] = scale * s;
}

int launch_matvec_neg(pnol_ctx * ctx, const double * D, const double * g, int n, double * p)
{
	TimerScope ts(ctx, "matvec_neg");
	long long threads = (long long) n * 32;
	if (n % 4 == 0 && (((size_t) D) & 31) == 0 && (((size_t) g) & 31) == 0)
		PNOL_LAUNCH(ctx, matvec_v4_kernel, (unsigned) ((threads + kMvThreads - 1) / kMvThreads), kMvThreads, 0, D, g, n, -1.0, p);
	else
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

// ---- fast path of the rank-2 form (n a multiple of 4, D 32-byte aligned): three launches, 256-bit loads -------------------------
// pass 1: tile = R rows x 256 columns per block of 8 warps (R a multiple of 4, chosen by the launcher so that the grid is a whole
// number of waves). A lane owns 8 columns (two groups of four) and keeps their column sums v in REGISTERS over the rows of its warp
// (the first version kept them in shared memory: one LDS + one STS per element of D); a warp takes four rows per batch, i.e. eight
// 256-bit loads per lane in flight. The warps' column sums are added in warp order in shared memory: one partial row per block.
// Partials: upart[column block][row], vpart[row block][column].
constexpr int kR4Cols = 256;

__global__ void __launch_bounds__(256, 2)
hinv_pass1_v4_kernel(const double * __restrict__ D, const double * __restrict__ g, int n, int R, double * __restrict__ upart,
                     double * __restrict__ vpart)
{
	__shared__ double vs[8][kR4Cols];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int r0 = blockIdx.x * R, r1 = min(n, r0 + R);
	const int c0 = blockIdx.y * kR4Cols + lane * 4;
	double v[2][4], gc[2][4];
#pragma unroll
	for (int q = 0; q < 2; q++) {
#pragma unroll
		for (int e = 0; e < 4; e++) { v[q][e] = 0; gc[q][e] = 0; }
		if (c0 + 128 * q < n) ldg256_cached(g + c0 + 128 * q, gc[q]);
	}
	for (int row = r0 + warp * 4; row < r1; row += 32) {
		const double * Dr = D + (long long) row * n + c0;
		double d[4][2][4];
#pragma unroll
		for (int h = 0; h < 4; h++)
#pragma unroll
			for (int q = 0; q < 2; q++)
				if (row + h < r1 && c0 + 128 * q < n) ldg256_stream(Dr + (long long) h * n + 128 * q, d[h][q]);
#pragma unroll
		for (int h = 0; h < 4; h++) {
			if (row + h >= r1) break;
			const double gi = g[row + h];
			double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
			for (int q = 0; q < 2; q++) {
				if (c0 + 128 * q < n) {
					s0 = fma(d[h][q][0], gc[q][0], s0); s1 = fma(d[h][q][1], gc[q][1], s1);
					s2 = fma(d[h][q][2], gc[q][2], s2); s3 = fma(d[h][q][3], gc[q][3], s3);
#pragma unroll
					for (int e = 0; e < 4; e++) v[q][e] = fma(gi, d[h][q][e], v[q][e]);
				}
			}
			double s = (s0 + s1) + (s2 + s3);
			for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
			if (lane == 0) upart[(size_t) blockIdx.y * n + row + h] = s;
		}
	}
#pragma unroll
	for (int q = 0; q < 2; q++)
		*reinterpret_cast<double4 *>(&vs[warp][lane * 4 + 128 * q]) = make_double4(v[q][0], v[q][1], v[q][2], v[q][3]);
	__syncthreads();
	const int c = blockIdx.y * kR4Cols + threadIdx.x;
	if (c < n) {
		double t = 0;
#pragma unroll
		for (int w = 0; w < 8; w++) t = t + vs[w][threadIdx.x];
		vpart[(size_t) blockIdx.x * n + c] = t;
	}
}

// finish: out[k] = sum over parts, fixed order. A block takes 32 columns x 8 part groups (all loads of a thread independent), the
// eight group sums of a column are added in order. blockIdx.y = 0: v from vpart, 1: u from upart.
__global__ void __launch_bounds__(256)
hinv_finish_v4_kernel(const double * __restrict__ vpart, int nvparts, const double * __restrict__ upart, int nuparts, int n,
                      double * __restrict__ v, double * __restrict__ u)
{
	__shared__ double red[8][33];
	const double * part = blockIdx.y ? upart : vpart;
	const int nparts = blockIdx.y ? nuparts : nvparts;
	double * out = blockIdx.y ? u : v;
	const int c = threadIdx.x & 31, pg = threadIdx.x >> 5;
	const int k = blockIdx.x * 32 + c;
	const int per = (nparts + 7) / 8;
	double s = 0;
	if (k < n) {
		const int p1 = min(nparts, (pg + 1) * per);
#pragma unroll 8
		for (int p = pg * per; p < p1; p++) s = s + part[(size_t) p * n + k];
	}
	red[pg][c] = s;
	__syncthreads();
	if (pg == 0 && k < n) {
		double t = 0;
#pragma unroll
		for (int q = 0; q < 8; q++) t = t + red[q][c];
		out[k] = t;
	}
}

// pass 2: D_ij += -rho s_i v_j - rho u_i s_j + (rho^2 gamma + rho) s_i s_j, products associated as in hinv_pass2_kernel (same bits).
// Every block first computes g.s and g.u itself (fixed-order tree over 256 threads: identical in every block; saves the one-block
// dot-product launch), then walks (row, 1024-column chunk) items, four items' 256-bit loads in flight per thread.
__global__ void __launch_bounds__(256)
hinv_pass2_v4_kernel(double * __restrict__ D, const double * __restrict__ g, const double * __restrict__ s, const double * __restrict__ u,
                     const double * __restrict__ v, int n, double * __restrict__ scal_out)
{
	__shared__ double r0[256], r1[256];
	{
		double a = 0, b = 0;
		for (int k = threadIdx.x; k < n; k += 256) { a = fma(g[k], s[k], a); b = fma(g[k], u[k], b); }
		r0[threadIdx.x] = a; r1[threadIdx.x] = b;
		__syncthreads();
		for (int o = 128; o > 0; o >>= 1) {
			if (threadIdx.x < o) { r0[threadIdx.x] += r0[threadIdx.x + o]; r1[threadIdx.x] += r1[threadIdx.x + o]; }
			__syncthreads();
		}
	}
	const double gs = r0[0], gu = r1[0];
	if (blockIdx.x == 0 && threadIdx.x == 0 && scal_out) { scal_out[0] = gs; scal_out[1] = gu; }
	const double rho = 1.0 / gs;
	const double cc = rho * rho * gu + rho;
	const int chunks = (n + 1023) / 1024;
	const long long items = (long long) n * chunks;
	const int col = threadIdx.x * 4;
	for (long long it0 = blockIdx.x; it0 < items; it0 += (long long) gridDim.x * 4) {
		double d[4][4];
		int row[4], cj[4];
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const long long it = it0 + (long long) q * gridDim.x;
			row[q] = (int) (it / chunks);
			cj[q] = (int) (it - (long long) row[q] * chunks) * 1024 + col;
			if (it >= items || cj[q] >= n) row[q] = -1;
			if (row[q] >= 0) ldg256_stream(D + (long long) row[q] * n + cj[q], d[q]);
		}
#pragma unroll
		for (int q = 0; q < 4; q++) {
			if (row[q] < 0) continue;
			const double si = s[row[q]], ui = u[row[q]];
			const double a = -rho * si, b = rho * ui, c = cc * si;
			double sv[4], vv[4];
			ldg256_cached(s + cj[q], sv);
			ldg256_cached(v + cj[q], vv);
			double o[4];
#pragma unroll
			for (int e = 0; e < 4; e++) o[e] = d[q][e] + (a * vv[e] - b * sv[e] + c * sv[e]);
			*reinterpret_cast<double4 *>(D + (long long) row[q] * n + cj[q]) = make_double4(o[0], o[1], o[2], o[3]);
		}
	}
}

int launch_hinv_rank2(pnol_ctx * ctx, double * D, const double * g, const double * s, int n)
{
	TimerScope ts(ctx, "hinv_rank2");
	if (n % 4 == 0 && (((size_t) D) & 31) == 0 && (((size_t) g) & 31) == 0 && (((size_t) s) & 31) == 0) {
		// row blocks: about 112 rows each, rounded so that the grid is a whole number of waves of 2 resident blocks per SM
		const int ncb = (n + kR4Cols - 1) / kR4Cols;
		const int wave = ctx->sm_count * 2;
		int nrb = (n + 111) / 112;
		const long long waves = ((long long) nrb * ncb + wave / 2) / wave;
		if (waves >= 1) nrb = (int) ((waves * wave) / ncb);
		if (nrb < 1) nrb = 1;
		int R = ((n + nrb - 1) / nrb + 3) & ~3;
		nrb = (n + R - 1) / R;
		const int nvparts = nrb;
		const size_t n4 = ((size_t) n + 3) & ~(size_t) 3;
		PNOL_CHECK(ws_reserve(ctx, 1, (2 * n4 + 4 + (size_t) nvparts * n + (size_t) ncb * n) * sizeof(double)));
		double * u = (double *) ctx->ws[1];
		double * v = u + n4;
		double * scal = v + n4;
		double * vpart = scal + 4;
		double * upart = vpart + (size_t) nvparts * n;
		PNOL_LAUNCH(ctx, hinv_pass1_v4_kernel, dim3(nrb, ncb), 256, 0, D, g, n, R, upart, vpart);
		PNOL_LAUNCH(ctx, hinv_finish_v4_kernel, dim3((n + 31) / 32, 2), 256, 0, vpart, nvparts, upart, ncb, n, v, u);
		const long long items = (long long) n * ((n + 1023) / 1024);
		long long grid = (long long) ctx->sm_count * 2 * 4;      // 2 resident blocks per SM (86 registers): four whole waves
		if (grid > items) grid = items;
		PNOL_LAUNCH(ctx, hinv_pass2_v4_kernel, (unsigned) grid, 256, 0, D, g, s, u, v, n, scal);
		return PNOL_OK;
	}
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
