// eval_kernels.cu -- scalar-objective kernels: population / batch sweep, forward-difference point
// evaluation (gradient, Recur gradient, Hessian) and the alpha-pool evaluation of the pooled line searches.
//
// Mapping (north_star): one evaluation point per THREAD, the shared base point staged in shared memory;
// population tiles are read from HBM with coalesced 16-byte loads and parked in shared memory with an odd row
// pitch so that thread r walking row r is bank-conflict free. Functor code is compiled with -fmad=false.
#include "common.cuh"

#include <stdlib.h>

namespace pnol {

// ---------------------------------------------------------------------------------------------------
// a1 / a15: batch sweep. Replaces GeneticAlgorithmMPI::evaluatePopulationParallel
// (Source/GeneticAlgorithmMPI.cpp:283-414): F[i] = objEval(Xpop[i]) for rows with evaluateIndicator[i].
// ---------------------------------------------------------------------------------------------------
constexpr int kSweepThreads = 128;

template <class F>
__global__ void __launch_bounds__(kSweepThreads)
eval_batch_tile_kernel(FunctorParams P, const double * __restrict__ pts, long long B, int n, long long ld,
                       const unsigned char * __restrict__ indicator, double * __restrict__ f_out, int pitch)
{
	extern __shared__ double tile[];   // kSweepThreads rows x pitch
	const int tid = threadIdx.x;
	for (long long row0 = (long long) blockIdx.x * kSweepThreads; row0 < B; row0 += (long long) gridDim.x * kSweepThreads) {
		const int rows = (int) min((long long) kSweepThreads, B - row0);
		const bool mine = tid < rows && (indicator == nullptr || indicator[row0 + tid] != 0);
		// skip tiles with nothing to evaluate (the elite block of a GA generation)
		if (!__syncthreads_or(mine)) continue;

		if (ld == n && (n & 1) == 0 && ((((size_t) (pts + row0 * ld)) & 15) == 0)) {
			// contiguous tile: 16-byte coalesced loads
			const double2 * src = reinterpret_cast<const double2 *>(pts + row0 * ld);
			const int n2 = n >> 1;
			const int total2 = rows * n2;
			for (int e = tid; e < total2; e += kSweepThreads) {
				double2 v = __ldg(src + e);
				int r = e / n2, c = (e - r * n2) * 2;
				tile[r * pitch + c] = v.x;
				tile[r * pitch + c + 1] = v.y;
			}
		} else {
			const int total = rows * n;
			for (int e = tid; e < total; e += kSweepThreads) {
				int r = e / n, c = e - r * n;
				tile[r * pitch + c] = pts[(row0 + r) * ld + c];
			}
		}
		__syncthreads();
		if (mine) {
			PtrAcc acc{tile + tid * pitch};
			f_out[row0 + tid] = F::eval(P, acc, n);
		}
		__syncthreads();
	}
}

// Separable objectives (functors.hpp: kSeparable): f = init + sum_k term(x_k) in index order. One WARP takes 32 individuals:
// the terms are computed one gene per lane directly from coalesced row loads (no staging of the inputs, no block-wide
// barrier, so the warps of an SM drift apart and loads overlap arithmetic), parked in the warp's private shared-memory tile
// (odd pitch), and lane r then adds up the terms of individual r in index order -- the same operations in the same order as
// F::eval, hence the same bits. The tile kernel below spends its time in lock-step load / compute phases instead.
template <class F> struct is_separable {
	template <class T> static constexpr bool test(decltype(T::kSeparable) *) { return T::kSeparable; }
	template <class T> static constexpr bool test(...) { return false; }
	static constexpr bool value = test<F>(nullptr);
};

constexpr int kSepThreads = 128;

template <class F, int G>
__global__ void __launch_bounds__(kSepThreads)
eval_batch_separable_kernel(FunctorParams P, const double * __restrict__ pts, long long B, int n, long long ld,
                            const unsigned char * __restrict__ indicator, double * __restrict__ f_out, int pitch)
{
	extern __shared__ double sm[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	double * tile = sm + (size_t) warp * 32 * pitch;
	const long long nbatch = (B + 31) / 32;
	const long long gwarp = (long long) blockIdx.x * (kSepThreads / 32) + warp, gwarps = (long long) gridDim.x * (kSepThreads / 32);
	for (long long b = gwarp; b < nbatch; b += gwarps) {
		const long long row0 = b * 32;
		const int rows = (int) min((long long) 32, B - row0);
		const bool mine = lane < rows && (indicator == nullptr || indicator[row0 + lane] != 0);
		if (!__any_sync(0xffffffffu, mine)) continue;       // e.g. the elite block of a GA generation
		const double * src = pts + row0 * ld;
		if (n == 32 && rows == 32) {
			// software pipeline: the loads of the next 8 rows are in flight while the terms of these 8 are computed
			double cur[G], nxt[G];
#pragma unroll
			for (int q = 0; q < G; q++) cur[q] = __ldg(src + q * ld + lane);
#pragma unroll
			for (int r0 = 0; r0 < 32; r0 += G) {
				if (r0 + G < 32) {
#pragma unroll
					for (int q = 0; q < G; q++) nxt[q] = __ldg(src + (r0 + G + q) * ld + lane);
				}
#pragma unroll
				for (int q = 0; q < G; q++) tile[(r0 + q) * pitch + lane] = F::sep_term(P, cur[q]);
#pragma unroll
				for (int q = 0; q < G; q++) cur[q] = nxt[q];
			}
		} else if (n == 32) {
			for (int r = 0; r < rows; r++) tile[r * pitch + lane] = F::sep_term(P, __ldg(src + r * ld + lane));
		} else {
			for (int r = 0; r < rows; r++)
				for (int k = lane; k < n; k += 32) tile[r * pitch + k] = F::sep_term(P, __ldg(src + r * ld + k));
		}
		__syncwarp();
		if (mine) {
			double v = F::sep_init(P, n);
			const double * mt = tile + lane * pitch;
#pragma unroll 8
			for (int k = 0; k < n; k++) v = v + mt[k];
			f_out[row0 + lane] = v;
		}
		__syncwarp();
	}
}

// Variant without shared memory: one individual per THREAD, its genes read with 256-bit loads (one full 32-byte sector per
// load, so the uncoalesced row walk still moves only the bytes it needs), four independent term chains in flight per thread.
__device__ __forceinline__ void ldg_f64x4(const double * p, double (&v)[4])
{
	asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

__device__ __forceinline__ void ldg_f64x4p(const double * p, double * v)
{
	asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

template <class F>
__global__ void __launch_bounds__(128)
eval_batch_rowwise_kernel(FunctorParams P, const double * __restrict__ pts, long long B, int n, long long ld,
                          const unsigned char * __restrict__ indicator, double * __restrict__ f_out)
{
	const long long b = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= B) return;
	if (indicator && !indicator[b]) return;
	const double * row = pts + b * ld;
	double v = F::sep_init(P, n);
	double cur[4], nxt[4];
	ldg_f64x4(row, cur);
	for (int k = 0; k < n; k += 4) {
		if (k + 4 < n) ldg_f64x4(row + k + 4, nxt);
		const double t0 = F::sep_term(P, cur[0]), t1 = F::sep_term(P, cur[1]), t2 = F::sep_term(P, cur[2]), t3 = F::sep_term(P, cur[3]);
		v = v + t0; v = v + t1; v = v + t2; v = v + t3;
#pragma unroll
		for (int q = 0; q < 4; q++) cur[q] = nxt[q];
	}
	f_out[b] = v;
}

// same, n a multiple of 16: four 256-bit loads (16 genes) in flight per thread ahead of the arithmetic, so that the memory system
// keeps working through the FP64-heavy terms (the kernel's FP64 issue time and its HBM time are about equal; with one load ahead
// they overlapped badly). Same order of additions, hence the same bits.
template <class F>
__global__ void __launch_bounds__(128)
eval_batch_rowwise16_kernel(FunctorParams P, const double * __restrict__ pts, long long B, int n, long long ld,
                            const unsigned char * __restrict__ indicator, double * __restrict__ f_out)
{
	const long long b = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= B) return;
	if (indicator && !indicator[b]) return;
	const double * row = pts + b * ld;
	double v = F::sep_init(P, n);
	double cur[16], nxt[16];
#pragma unroll
	for (int q = 0; q < 4; q++) ldg_f64x4p(row + 4 * q, cur + 4 * q);
	for (int k = 0; k < n; k += 16) {
		if (k + 16 < n) {
#pragma unroll
			for (int q = 0; q < 4; q++) ldg_f64x4p(row + k + 16 + 4 * q, nxt + 4 * q);
		}
#pragma unroll
		for (int q = 0; q < 16; q++) v = v + F::sep_term(P, cur[q]);
#pragma unroll
		for (int q = 0; q < 16; q++) cur[q] = nxt[q];
	}
	f_out[b] = v;
}

template <class F, bool kSep = is_separable<F>::value> struct SeparableLaunch {
	static int run(pnol_ctx *, const pnol_functor *, const double *, long long, int, long long, const unsigned char *, double *, bool * done)
	{
		*done = false;
		return PNOL_OK;
	}
};
template <class F> struct SeparableLaunch<F, true> {
	static int run(pnol_ctx * ctx, const pnol_functor * f, const double * pts, long long B, int n, long long ld,
	               const unsigned char * indicator, double * f_out, bool * done)
	{
		*done = false;
		const int pitch = n | 1;
		const size_t smem = (size_t) (kSepThreads / 32) * 32 * pitch * sizeof(double);
		const bool rowwise_ok = n % 4 == 0 && ld % 4 == 0 && (((size_t) pts) & 31) == 0;
		if (!rowwise_ok && (smem > ctx->smem_optin / 2 || B < 64 || n < 24)) return PNOL_OK;   // long / very short genomes, tiny batches: the generic kernels
		// default: the row-wise kernel (0.059 ms at 1M x 32, 68 % of the HBM roofline; the warp-tile kernel below it 0.078 ms).
		// PNOL_SWEEP_G = 4 / 8 / 16 forces the warp-tile kernel with that prefetch depth (tuning runs).
		static const int g = [] { const char * e = getenv("PNOL_SWEEP_G"); return e ? atoi(e) : 0; }();
		if (g == 0 && n % 4 == 0 && ld % 4 == 0 && (((size_t) pts) & 31) == 0) {
			static const int deep = [] { const char * e = getenv("PNOL_SWEEP_DEEP"); return e ? atoi(e) : 1; }();      // 0: one load ahead (A/B runs)
			if (deep && n % 16 == 0)
				PNOL_LAUNCH(ctx, eval_batch_rowwise16_kernel<F>, (unsigned) ((B + 127) / 128), 128, 0, f->params, pts, B, n, ld, indicator, f_out);
			else
				PNOL_LAUNCH(ctx, eval_batch_rowwise_kernel<F>, (unsigned) ((B + 127) / 128), 128, 0, f->params, pts, B, n, ld, indicator, f_out);
			*done = true;
			return PNOL_OK;
		}
		auto go = [&](auto kern) -> int {
			PNOL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			int per_sm = 1;
			PNOL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSepThreads, smem));
			if (per_sm < 1) per_sm = 1;
			const long long blocks = ((B + 31) / 32 + kSepThreads / 32 - 1) / (kSepThreads / 32);
			static const int waves = [] { const char * e = getenv("PNOL_SWEEP_WAVES"); return e ? atoi(e) : 1; }();   // tuning override
			const long long grid = min(blocks, (long long) ctx->sm_count * per_sm * waves);
			PNOL_LAUNCH(ctx, kern, (unsigned) grid, kSepThreads, smem, f->params, pts, B, n, ld, indicator, f_out, pitch);
			return PNOL_OK;
		};
		*done = true;
		if (g == 4) return go(eval_batch_separable_kernel<F, 4>);
		if (g == 16) return go(eval_batch_separable_kernel<F, 16>);
		return go(eval_batch_separable_kernel<F, 8>);
	}
};

// large-n fallback: one thread per row straight from global memory
template <class F>
__global__ void eval_batch_direct_kernel(FunctorParams P, const double * __restrict__ pts, long long B, int n, long long ld,
                                         const unsigned char * __restrict__ indicator, double * __restrict__ f_out)
{
	long long b = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= B) return;
	if (indicator && !indicator[b]) return;
	PtrAcc acc{pts + b * ld};
	f_out[b] = F::eval(P, acc, n);
}

int launch_eval_batch(pnol_ctx * ctx, const pnol_functor * f, const double * pts, long long B, int n, long long ld,
                      const unsigned char * indicator, double * f_out)
{
	if (B <= 0) return PNOL_OK;
	TimerScope ts(ctx, "eval_batch");
	return dispatch_scalar(ctx, f->kind, [&](auto tag) -> int {
		using F = decltype(tag);
		bool done = false;
		PNOL_CHECK((SeparableLaunch<F>::run(ctx, f, pts, B, n, ld, indicator, f_out, &done)));
		if (done) return PNOL_OK;
		int pitch = n | 1;
		size_t smem = (size_t) kSweepThreads * pitch * sizeof(double);
		if (smem <= ctx->smem_optin) {
			auto kern = eval_batch_tile_kernel<F>;
			PNOL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			long long tiles = (B + kSweepThreads - 1) / kSweepThreads;
			int per_sm = (int) max((size_t) 1, min((size_t) 8, ctx->smem_optin / max(smem, (size_t) 1)));
			long long grid = min(tiles, (long long) ctx->sm_count * per_sm);
			PNOL_LAUNCH(ctx, kern, (unsigned) grid, kSweepThreads, smem, f->params, pts, B, n, ld, indicator, f_out, pitch);
		} else {
			auto kern = eval_batch_direct_kernel<F>;
			PNOL_LAUNCH(ctx, kern, (unsigned) ((B + 127) / 128), 128, 0, f->params, pts, B, n, ld, indicator, f_out);
		}
		return PNOL_OK;
	});
}

// ---------------------------------------------------------------------------------------------------
// a3 / a4 / a5: forward-difference stencil points. Point i in [i0, i1) is the base point with coordinate
// pos[i] (pos == nullptr: i itself) incremented by dx[i]; the base value f0 is produced by one extra thread
// when f0_out != nullptr. Replaces the evaluation loops of Objective::gradientApproximation[MPI]
// (Source/PNOL_Objective.cpp:19-32, 125-145) and the Recur variants (:345-358, :399-420).
// ---------------------------------------------------------------------------------------------------
constexpr int kFdThreads = 32;

template <class F>
__global__ void __launch_bounds__(kFdThreads)
fd_points_kernel(FunctorParams P, const double * __restrict__ xfull, int nfull, const int * __restrict__ pos,
                 const double * __restrict__ dx, int i0, int i1, double * __restrict__ fdx_out, double * __restrict__ f0_out)
{
	extern __shared__ double xs[];
	for (int j = threadIdx.x; j < nfull; j += blockDim.x) xs[j] = xfull[j];
	__syncthreads();
	int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
	if (i < i1) {
		// pos[i] < 0: reduced variable i has no slot in the full point (more reduced entries than free variables: the reference's
		// objEvalRecur never reads it, Source/PNOL_Objective.cpp:311-323), so its stencil point is the base point itself
		int pi = pos ? pos[i] : i;
		PerturbAcc acc{xs, pi, pi >= 0 ? xs[pi] + dx[i] : 0.0};   // XdX[i] = XdX[i] + dX[i]  (Source/PNOL_Objective.cpp:27)
		fdx_out[i] = F::eval(P, acc, nfull);
	} else if (i == i1 && f0_out) {
		PtrAcc acc{xs};
		*f0_out = F::eval(P, acc, nfull);
	}
}

int launch_fd_points(pnol_ctx * ctx, const pnol_functor * f, const double * xfull, int nfull, const int * pos,
                     const double * dx, int i0, int i1, double * fdx_out, double * f0_out)
{
	TimerScope ts(ctx, "fd_points");
	return dispatch_scalar(ctx, f->kind, [&](auto tag) -> int {
		using F = decltype(tag);
		int npts = (i1 - i0) + (f0_out ? 1 : 0);
		if (npts <= 0) return PNOL_OK;
		size_t smem = (size_t) nfull * sizeof(double);
		PNOL_REQUIRE(ctx, smem <= ctx->smem_optin, "fd stencil: n = %d does not fit in shared memory", nfull);
		auto kern = fd_points_kernel<F>;
		PNOL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		PNOL_LAUNCH(ctx, kern, (unsigned) ((npts + kFdThreads - 1) / kFdThreads), kFdThreads, smem, f->params, xfull, nfull,
		            pos, dx, i0, i1, fdx_out, f0_out);
		return PNOL_OK;
	});
}

// g[i] = (fdx[i] - f0) / dx[i]     (Source/PNOL_Objective.cpp:31, :150-153, :451-454)
__global__ void fd_quotient_kernel(const double * __restrict__ fdx, const double * __restrict__ f0,
                                   const double * __restrict__ dx, int n, double * __restrict__ g)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) g[i] = (fdx[i] - *f0) / dx[i];
}

int launch_fd_quotient(pnol_ctx * ctx, const double * fdx, const double * f0, const double * dx, int n, double * g)
{
	if (n <= 0) return PNOL_OK;
	PNOL_LAUNCH(ctx, fd_quotient_kernel, (n + 255) / 256, 256, 0, fdx, f0, dx, n, g);
	return PNOL_OK;
}

// Active-set assembly (Objective::objEvalRecur, Source/PNOL_Objective.cpp:303-323): the full point takes
// const_x[j] where const_ind[j], else the next reduced variable. Also emits pos[i] (full index of reduced
// variable i; the caller presets pos to -1, which stays for reduced entries beyond the number of free variables). Single block; a serial scan over chunks of blockDim entries.
__global__ void assemble_recur_kernel(const double * __restrict__ xr, int nr,
                                      const double * __restrict__ const_x, const unsigned char * __restrict__ const_ind,
                                      int nfull, double * __restrict__ xfull, int * __restrict__ pos,
                                      int * __restrict__ nr_found)
{
	__shared__ int warp_tot[32];
	__shared__ int base;
	if (threadIdx.x == 0) base = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	for (int j0 = 0; j0 < nfull; j0 += blockDim.x) {
		int j = j0 + threadIdx.x;
		int freev = (j < nfull && !(const_ind && const_ind[j])) ? 1 : 0;
		int incl = freev;
		for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
		if (lane == 31) warp_tot[wid] = incl;
		__syncthreads();
		int woff = 0;
		for (int w = 0; w < wid; w++) woff += warp_tot[w];
		int tot = 0;
		for (int w = 0; w < nw; w++) tot += warp_tot[w];
		int idx = base + woff + incl - freev;   // exclusive rank among free variables
		if (j < nfull) {
			if (freev) {
				double v = idx < nr ? xr[idx] : 0.0;
				xfull[j] = v;
				if (idx < nr && pos) pos[idx] = j;
			} else {
				xfull[j] = const_x[j];
			}
		}
		__syncthreads();
		if (threadIdx.x == 0) base += tot;
		__syncthreads();
	}
	if (threadIdx.x == 0 && nr_found) *nr_found = base;
}

int launch_assemble_recur(pnol_ctx * ctx, const double * xr, int nr, const double * const_x,
                          const unsigned char * const_ind, int nfull, double * xfull, int * pos, int * nr_found)
{
	PNOL_LAUNCH(ctx, assemble_recur_kernel, 1, 256, 0, xr, nr, const_x, const_ind, nfull, xfull, pos, nr_found);
	return PNOL_OK;
}

// ---------------------------------------------------------------------------------------------------
// a6: forward-difference Hessian (Objective::hessianApproximation, Source/PNOL_Objective.cpp:38-85).
// One thread per pair (i <= j): B_ij = (f_ij - f_i - f_j + f) / (dx_i dx_j), mirrored. f_i are the n stencil
// values already produced by fd_points_kernel (the reference recomputes the same value for every pair).
// ---------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(kFdThreads)
fd_hessian_kernel(FunctorParams P, const double * __restrict__ x, const double * __restrict__ dx, int n,
                  const double * __restrict__ fdx, const double * __restrict__ f0, double * __restrict__ Bout)
{
	extern __shared__ double xs[];
	for (int j = threadIdx.x; j < n; j += blockDim.x) xs[j] = x[j];
	__syncthreads();
	long long pair = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	long long npairs = (long long) n * (n + 1) / 2;
	if (pair >= npairs) return;
	// unrank pair -> (i, j), i <= j, row-major over the upper triangle
	int i = 0;
	{
		// row i starts at s(i) = i*n - i*(i-1)/2 ; solve by floating estimate then fix up
		double nn = (double) n;
		double est = (2.0 * nn + 1.0 - sqrt((2.0 * nn + 1.0) * (2.0 * nn + 1.0) - 8.0 * (double) pair)) * 0.5;
		i = (int) est;
		if (i < 0) i = 0;
		if (i > n - 1) i = n - 1;
		while (i > 0 && (long long) i * n - (long long) i * (i - 1) / 2 > pair) i--;
		while ((long long) (i + 1) * n - (long long) (i + 1) * i / 2 <= pair) i++;
	}
	int j = i + (int) (pair - ((long long) i * n - (long long) i * (i - 1) / 2));
	double fij;
	if (i == j) {
		// XdXij[i] = (X[i] + dX[i]) + dX[i]   (Source/PNOL_Objective.cpp:61-62 with i == j)
		PerturbAcc acc{xs, i, (xs[i] + dx[i]) + dx[i]};
		fij = F::eval(P, acc, n);
	} else {
		Perturb2Acc acc{xs, i, xs[i] + dx[i], j, xs[j] + dx[j]};
		fij = F::eval(P, acc, n);
	}
	double b = (fij - fdx[i] - fdx[j] + *f0) / (dx[i] * dx[j]);
	Bout[(long long) i * n + j] = b;
	Bout[(long long) j * n + i] = b;
}

int launch_fd_hessian(pnol_ctx * ctx, const pnol_functor * f, const double * x, const double * dx, int n,
                      const double * fdx, const double * f0, double * B)
{
	TimerScope ts(ctx, "fd_hessian");
	return dispatch_scalar(ctx, f->kind, [&](auto tag) -> int {
		using F = decltype(tag);
		size_t smem = (size_t) n * sizeof(double);
		PNOL_REQUIRE(ctx, smem <= ctx->smem_optin, "fd hessian: n = %d does not fit in shared memory", n);
		auto kern = fd_hessian_kernel<F>;
		PNOL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		long long npairs = (long long) n * (n + 1) / 2;
		PNOL_LAUNCH(ctx, kern, (unsigned) ((npairs + kFdThreads - 1) / kFdThreads), kFdThreads, smem, f->params, x, dx, n, fdx, f0, B);
		return PNOL_OK;
	});
}

// ---------------------------------------------------------------------------------------------------
// a13: alpha pool. Point k evaluates phi = f(x + alpha_k p) and (optionally) the forward-difference slope
// (f(x + (alpha_k + dalpha) p) - phi) / dalpha. Mirrors lineSearchObj / lineSearchFDDerivative
// (Source/BFGS_bnd_linesearch_MPI_SW.cpp:703-734: Xtemp[i] = X[i] + alpha*p[i], no contraction) and the
// NaN/inf -> 1e10 sentinel of evaluateAlphaPoolAndDerivatives (:657-668).
// ---------------------------------------------------------------------------------------------------
struct LineAcc {
	const double * x; const double * p; const unsigned char * is_const; double alpha;
	__device__ __forceinline__ double operator[](int j) const
	{
		if (is_const && is_const[j]) return x[j];
		return x[j] + alpha * p[j];
	}
};

template <class F>
__global__ void __launch_bounds__(32)
alpha_pool_kernel(FunctorParams P, const double * __restrict__ xfull, const double * __restrict__ pfull,
                  const unsigned char * __restrict__ is_const, int nfull, const double * __restrict__ alpha, int npool,
                  double dalpha, const unsigned char * __restrict__ eval_ind, double * __restrict__ vals /* 2*npool */)
{
	extern __shared__ double sm[];
	double * xs = sm;
	double * ps = sm + nfull;
	unsigned char * cs = reinterpret_cast<unsigned char *>(sm + 2 * nfull);
	for (int j = threadIdx.x; j < nfull; j += blockDim.x) {
		xs[j] = xfull[j]; ps[j] = pfull[j];
		cs[j] = is_const ? is_const[j] : 0;
	}
	__syncthreads();
	int t = blockIdx.x;   // one point per block (t < npool: phi, t >= npool: shifted point); the warp stages, lane 0 evaluates
	if (threadIdx.x != 0 || t >= 2 * npool) return;
	int k = t < npool ? t : t - npool;
	if (eval_ind && !eval_ind[k]) return;
	double a = t < npool ? alpha[k] : alpha[k] + dalpha;
	LineAcc acc{xs, ps, is_const ? cs : nullptr, a};
	vals[t] = F::eval(P, acc, nfull);
}

__global__ void alpha_pool_finish_kernel(const double * __restrict__ vals, int npool, double dalpha,
                                         const unsigned char * __restrict__ eval_ind, bool want_dphi,
                                         double * __restrict__ phi, double * __restrict__ dphi, int * __restrict__ bad)
{
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= npool) return;
	if (eval_ind && !eval_ind[k]) return;
	double ph = vals[k];
	int nbad = 0;
	if (want_dphi) dphi[k] = (vals[npool + k] - ph) / dalpha;   // slope keeps NaN/inf: the reference only guards phi
	if (ph != ph || isinf(ph)) { ph = 1e10; nbad++; }
	phi[k] = ph;
	if (nbad) atomicAdd(bad, nbad);
}

int launch_alpha_pool(pnol_ctx * ctx, const pnol_functor * f, const double * xfull, const double * pfull,
                      const unsigned char * is_const, int nfull, const double * alpha, int npool, double dalpha,
                      const unsigned char * eval_ind, double * phi, double * dphi, int * bad_dev)
{
	TimerScope ts(ctx, "alpha_pool");
	PNOL_CHECK(ws_reserve(ctx, 3, (size_t) 2 * npool * sizeof(double)));
	double * vals = (double *) ctx->ws[3];
	int st = dispatch_scalar(ctx, f->kind, [&](auto tag) -> int {
		using F = decltype(tag);
		size_t smem = (size_t) nfull * (2 * sizeof(double) + 1) + 16;
		PNOL_REQUIRE(ctx, smem <= ctx->smem_optin, "alpha pool: n = %d does not fit in shared memory", nfull);
		auto kern = alpha_pool_kernel<F>;
		PNOL_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		int npts = dphi ? 2 * npool : npool;
		// one point per block of one warp: the pool is tiny and each point is a long dependent chain
		PNOL_LAUNCH(ctx, kern, (unsigned) npts, 32, smem, f->params, xfull, pfull, is_const, nfull, alpha, npool,
		            dalpha, eval_ind, vals);
		return PNOL_OK;
	});
	PNOL_CHECK(st);
	PNOL_LAUNCH(ctx, alpha_pool_finish_kernel, (npool + 63) / 64, 64, 0, vals, npool, dalpha, eval_ind, dphi != nullptr, phi, dphi, bad_dev);
	return PNOL_OK;
}

} // namespace pnol
