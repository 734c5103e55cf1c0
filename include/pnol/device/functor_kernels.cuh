/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root). */
/*
 * functor_kernels.cuh -- the library's generic kernels and launchers as templates over a device functor, so that an objective
 * that is NOT one of the built-ins gets the same device path WITHOUT rebuilding libpnol_b200.so.
 *
 * The reference lets a user plug any objective into every stencil and algorithm by subclassing Objective / MultiObjective and
 * implementing objEval (Source/PNOL_Objective.hpp:29, :57). Here the device side of that plug-in point is:
 *
 *   1. write the functor once, __host__ __device__, to the concept of include/pnol/functors.hpp (scalar: eval(P, X, n);
 *      residual model: residual(P, X, n, i); optionally kSeparable / sep_init / sep_term for the fast population sweep);
 *   2. in ONE .cu file of your own, compiled with
 *        nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -ffp-contract=off -I<repo>/include ...
 *      write   PNOL_REGISTER_SCALAR_FUNCTOR(kind, MyFunctor, n_columns)   or   PNOL_REGISTER_RESIDUAL_FUNCTOR(kind, MyModel, n_columns)
 *      with a kind in [PNOL_F_USER_SCALAR_BASE, +1000) resp. [PNOL_F_USER_RESIDUAL_BASE, +1000);
 *   3. link that object into your program / shared library next to libpnol_b200.so. The macro registers a launch table
 *      (pnol_functor_vtable, include/pnol_b200.h) at load time; pnol_functor_create(kind) and every entry point that takes a
 *      functor -- pnol_eval_batch, pnol_fd_gradient[_recur], pnol_fd_hessian, pnol_alpha_pool, the GA, pnol_residual_eval,
 *      pnol_fd_jacobian, pnol_lm_step / pnol_lm_iterate -- then launch YOUR instantiations of the kernels below on the
 *      context's stream. The built-in objectives run through exactly these templates (csrc/eval_kernels.cu, residual_kernels.cu).
 *
 * Everything here is header-only device / launch code; it needs nothing from the library but pnol_register_functor.
 */
#ifndef PNOL_DEVICE_FUNCTOR_KERNELS_CUH_
#define PNOL_DEVICE_FUNCTOR_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../pnol_b200.h"
#include "../functors.hpp"
#include "exact_div.cuh"

namespace pnol {
namespace dev {

/* ---- launch plumbing over the C launch environment (include/pnol_b200.h: pnol_launch_env) ---- */
#define PNOL_DEV_ERR(env, ...)                                                   \
	do {                                                                         \
		if ((env)->err && (env)->err_len) snprintf((env)->err, (env)->err_len, __VA_ARGS__); \
	} while (0)
#define PNOL_DEV_CUDA(env, call)                                                                      \
	do {                                                                                              \
		cudaError_t _e = (call);                                                                      \
		if (_e != cudaSuccess) {                                                                      \
			PNOL_DEV_ERR(env, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));  \
			return PNOL_ERR_CUDA;                                                                     \
		}                                                                                             \
	} while (0)
#define PNOL_DEV_REQUIRE(env, cond, ...)        \
	do {                                        \
		if (!(cond)) {                          \
			PNOL_DEV_ERR(env, __VA_ARGS__);     \
			return PNOL_ERR_INVALID;            \
		}                                       \
	} while (0)
/* every launch is counted: pnol_ctx_launches() stays exact for user functors too */
#define PNOL_DEV_LAUNCH(env, kernel, grid, block, smem, ...)                                          \
	do {                                                                                              \
		kernel<<<(grid), (block), (smem), (cudaStream_t) (env)->stream>>>(__VA_ARGS__);               \
		if ((env)->launches) ++*(env)->launches;                                                      \
		cudaError_t _e = cudaGetLastError();                                                          \
		if (_e != cudaSuccess) {                                                                      \
			PNOL_DEV_ERR(env, "%s:%d: launch %s -> %s", __FILE__, __LINE__, #kernel, cudaGetErrorString(_e)); \
			return PNOL_ERR_CUDA;                                                                     \
		}                                                                                             \
	} while (0)

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


template <class F, bool kSep = is_separable<F>::value> struct SeparableLaunch {
	static int run(const pnol_launch_env *, const FunctorParams &, const double *, long long, int, long long, const unsigned char *, double *, bool * done)
	{
		*done = false;
		return PNOL_OK;
	}
};
template <class F> struct SeparableLaunch<F, true> {
	static int run(const pnol_launch_env * env, const FunctorParams & P, const double * pts, long long B, int n, long long ld,
	               const unsigned char * indicator, double * f_out, bool * done)
	{
		*done = false;
		const int pitch = n | 1;
		const size_t smem = (size_t) (kSepThreads / 32) * 32 * pitch * sizeof(double);
		const bool rowwise_ok = n % 4 == 0 && ld % 4 == 0 && (((size_t) pts) & 31) == 0;
		if (!rowwise_ok && (smem > env->smem_optin / 2 || B < 64 || n < 24)) return PNOL_OK;   // long / very short genomes, tiny batches: the generic kernels
		// default: the row-wise kernel (0.059 ms at 1M x 32, 68 % of the HBM roofline; the warp-tile kernel below it 0.078 ms).
		// PNOL_SWEEP_G = 4 / 8 / 16 forces the warp-tile kernel with that prefetch depth (tuning runs).
		static const int g = [] { const char * e = getenv("PNOL_SWEEP_G"); return e ? atoi(e) : 0; }();
		if (g == 0 && rowwise_ok) {
			static const int deep = [] { const char * e = getenv("PNOL_SWEEP_DEEP"); return e ? atoi(e) : 1; }();      // 0: one load ahead (A/B runs)
			if (deep && n % 16 == 0)
				PNOL_DEV_LAUNCH(env, eval_batch_rowwise16_kernel<F>, (unsigned) ((B + 127) / 128), 128, 0, P, pts, B, n, ld, indicator, f_out);
			else
				PNOL_DEV_LAUNCH(env, eval_batch_rowwise_kernel<F>, (unsigned) ((B + 127) / 128), 128, 0, P, pts, B, n, ld, indicator, f_out);
			*done = true;
			return PNOL_OK;
		}
		auto go = [&](auto kern) -> int {
			PNOL_DEV_CUDA(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			int per_sm = 1;
			PNOL_DEV_CUDA(env, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSepThreads, smem));
			if (per_sm < 1) per_sm = 1;
			const long long blocks = ((B + 31) / 32 + kSepThreads / 32 - 1) / (kSepThreads / 32);
			static const int waves = [] { const char * e = getenv("PNOL_SWEEP_WAVES"); return e ? atoi(e) : 1; }();   // tuning override
			const long long grid = blocks < (long long) env->sm_count * per_sm * waves ? blocks : (long long) env->sm_count * per_sm * waves;
			PNOL_DEV_LAUNCH(env, kern, (unsigned) grid, kSepThreads, smem, P, pts, B, n, ld, indicator, f_out, pitch);
			return PNOL_OK;
		};
		*done = true;
		if (g == 4) return go(eval_batch_separable_kernel<F, 4>);
		if (g == 16) return go(eval_batch_separable_kernel<F, 16>);
		return go(eval_batch_separable_kernel<F, 8>);
	}
};

/* a1 / a15: f_out[b] = F(pts[b * ld ..]) for rows with indicator[b] != 0 */
template <class F>
int eval_batch(const pnol_launch_env * env, const FunctorParams & P, const double * pts, long long B, int n, long long ld,
               const unsigned char * indicator, double * f_out)
{
	if (B <= 0) return PNOL_OK;
	bool done = false;
	int st = SeparableLaunch<F>::run(env, P, pts, B, n, ld, indicator, f_out, &done);
	if (st != PNOL_OK || done) return st;
	int pitch = n | 1;
	size_t smem = (size_t) kSweepThreads * pitch * sizeof(double);
	if (smem <= env->smem_optin) {
		auto kern = eval_batch_tile_kernel<F>;
		PNOL_DEV_CUDA(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		long long tiles = (B + kSweepThreads - 1) / kSweepThreads;
		size_t q = env->smem_optin / (smem > 0 ? smem : 1);
		int per_sm = (int) (q < 1 ? 1 : (q > 8 ? 8 : q));
		long long grid = tiles < (long long) env->sm_count * per_sm ? tiles : (long long) env->sm_count * per_sm;
		PNOL_DEV_LAUNCH(env, kern, (unsigned) grid, kSweepThreads, smem, P, pts, B, n, ld, indicator, f_out, pitch);
	} else {
		auto kern = eval_batch_direct_kernel<F>;
		PNOL_DEV_LAUNCH(env, kern, (unsigned) ((B + 127) / 128), 128, 0, P, pts, B, n, ld, indicator, f_out);
	}
	return PNOL_OK;
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


template <class F>
int fd_points(const pnol_launch_env * env, const FunctorParams & P, const double * xfull, int nfull, const int * pos,
              const double * dx, int i0, int i1, double * fdx_out, double * f0_out)
{
	int npts = (i1 - i0) + (f0_out ? 1 : 0);
	if (npts <= 0) return PNOL_OK;
	size_t smem = (size_t) nfull * sizeof(double);
	PNOL_DEV_REQUIRE(env, smem <= env->smem_optin, "fd stencil: n = %d does not fit in shared memory", nfull);
	auto kern = fd_points_kernel<F>;
	PNOL_DEV_CUDA(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	PNOL_DEV_LAUNCH(env, kern, (unsigned) ((npts + kFdThreads - 1) / kFdThreads), kFdThreads, smem, P, xfull, nfull, pos, dx, i0, i1,
	                fdx_out, f0_out);
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


template <class F>
int fd_hessian(const pnol_launch_env * env, const FunctorParams & P, const double * x, const double * dx, int n,
               const double * fdx, const double * f0, double * B)
{
	size_t smem = (size_t) n * sizeof(double);
	PNOL_DEV_REQUIRE(env, smem <= env->smem_optin, "fd hessian: n = %d does not fit in shared memory", n);
	auto kern = fd_hessian_kernel<F>;
	PNOL_DEV_CUDA(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	long long npairs = (long long) n * (n + 1) / 2;
	PNOL_DEV_LAUNCH(env, kern, (unsigned) ((npairs + kFdThreads - 1) / kFdThreads), kFdThreads, smem, P, x, dx, n, fdx, f0, B);
	return PNOL_OK;
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


/* vals[t], t < npool: f(x + alpha_t p); vals[npool + t]: f(x + (alpha_t + dalpha) p) when want_shifted (the library turns them into
 * phi / dphi and applies the 1e10 sentinel) */
template <class F>
int alpha_pool(const pnol_launch_env * env, const FunctorParams & P, const double * xfull, const double * pfull,
               const unsigned char * is_const, int nfull, const double * alpha, int npool, double dalpha,
               const unsigned char * eval_ind, int want_shifted, double * vals)
{
	size_t smem = (size_t) nfull * (2 * sizeof(double) + 1) + 16;
	PNOL_DEV_REQUIRE(env, smem <= env->smem_optin, "alpha pool: n = %d does not fit in shared memory", nfull);
	auto kern = alpha_pool_kernel<F>;
	PNOL_DEV_CUDA(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	int npts = want_shifted ? 2 * npool : npool;
	// one point per block of one warp: the pool is tiny and each point is a long dependent chain
	PNOL_DEV_LAUNCH(env, kern, (unsigned) npts, 32, smem, P, xfull, pfull, is_const, nfull, alpha, npool, dalpha, eval_ind, vals);
	return PNOL_OK;
}

// ---------------------------------------------------------------------------------------------------
// generic residual evaluation: F[i] = r_i(x)        (MultiObjective::objEval, Source/PNOL_Objective.hpp:57)
// ---------------------------------------------------------------------------------------------------
template <class R>
__global__ void __launch_bounds__(256)
residual_kernel(FunctorParams P, const double * __restrict__ x, int n, double * __restrict__ F)
{
	extern __shared__ double xs[];
	for (int j = threadIdx.x; j < n; j += blockDim.x) xs[j] = x[j];
	__syncthreads();
	PtrAcc acc{xs};
	for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < P.m; i += (long long) gridDim.x * blockDim.x)
		F[i] = R::residual(P, acc, n, i);
}


template <class R>
int residual(const pnol_launch_env * env, const FunctorParams & P, const double * x, int n, double * F)
{
	const long long m = P.m;
	size_t smem = (size_t) n * sizeof(double);
	PNOL_DEV_REQUIRE(env, smem <= env->smem_optin, "residual: n = %d does not fit in shared memory", n);
	auto kern = residual_kernel<R>;
	PNOL_DEV_CUDA(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	long long blocks = (m + 255) / 256;
	long long grid = blocks < (long long) env->sm_count * 8 ? blocks : (long long) env->sm_count * 8;
	if (grid < 1) grid = 1;
	PNOL_DEV_LAUNCH(env, kern, (unsigned) grid, 256, smem, P, x, n, F);
	return PNOL_OK;
}

// ---------------------------------------------------------------------------------------------------
// black-box forward-difference Jacobian  (MultiObjective::gradientApproximation, Source/PNOL_Objective.cpp:165-197)
//   J[i][j] = (F_i(x + dx_j e_j) - F_i(x)) / dx_j
// ---------------------------------------------------------------------------------------------------
constexpr int kBbRows = 128;      // rows per block (one per thread)
constexpr int kBbCols = 32;       // J columns staged per pass

template <class R>
__global__ void __launch_bounds__(kBbRows)
fd_jacobian_blackbox_kernel(FunctorParams P, const double * __restrict__ x, const double * __restrict__ dx, int n,
                            double * __restrict__ J, double * __restrict__ F)
{
	extern __shared__ double sm[];
	double * xs = sm;                       // n
	double * dxs = sm + n;                  // n
	double * rdx = sm + 2 * n;              // n   RN(1/dx) or 0 (exact_div.cuh)
	double * tile = sm + 3 * n;             // kBbRows x (kBbCols + 1)
	constexpr int pitch = kBbCols + 1;
	for (int j = threadIdx.x; j < n; j += blockDim.x) { xs[j] = x[j]; dxs[j] = dx[j]; rdx[j] = make_recip(dx[j]).r; }
	__syncthreads();
	for (long long row0 = (long long) blockIdx.x * kBbRows; row0 < P.m; row0 += (long long) gridDim.x * kBbRows) {
		const long long i = row0 + threadIdx.x;
		const bool live = i < P.m;
		const int rows = (int) min((long long) kBbRows, P.m - row0);
		double r0 = 0;
		if (live) {
			PtrAcc acc{xs};
			r0 = R::residual(P, acc, n, i);
			if (F) F[i] = r0;
		}
		for (int c0 = 0; c0 < n; c0 += kBbCols) {
			const int cols = min(kBbCols, n - c0);
			if (live) {
				for (int c = 0; c < cols; c++) {
					const int j = c0 + c;
					PerturbAcc acc{xs, j, xs[j] + dxs[j]};       // XdX[j] = XdX[j] + dX[j]   (PNOL_Objective.cpp:186)
					double rj = R::residual(P, acc, n, i);
					RecipDiv rd; rd.d = dxs[j]; rd.r = rdx[j];
					tile[threadIdx.x * pitch + c] = div_exact(rj - r0, rd);   // (FdX[i] - F[i])/dX[j]  (:192)
				}
			}
			__syncthreads();
			// coalesced write: consecutive threads write consecutive columns of one row
			for (int e = threadIdx.x; e < rows * cols; e += kBbRows) {
				int r = e / cols, c = e - r * cols;
				J[(row0 + r) * n + c0 + c] = tile[r * pitch + c];
			}
			__syncthreads();
		}
	}
}


template <class R>
int fd_jacobian(const pnol_launch_env * env, const FunctorParams & P, const double * x, const double * dx, int n, double * J, double * F)
{
	const long long m = P.m;
	size_t smem = ((size_t) 3 * n + (size_t) kBbRows * (kBbCols + 1)) * sizeof(double);
	PNOL_DEV_REQUIRE(env, smem <= env->smem_optin, "fd jacobian: n = %d does not fit in shared memory", n);
	auto kern = fd_jacobian_blackbox_kernel<R>;
	PNOL_DEV_CUDA(env, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	long long blocks = (m + kBbRows - 1) / kBbRows;
	long long grid = blocks < (long long) env->sm_count * 4 ? blocks : (long long) env->sm_count * 4;
	if (grid < 1) grid = 1;
	PNOL_DEV_LAUNCH(env, kern, (unsigned) grid, kBbRows, smem, P, x, dx, n, J, F);
	return PNOL_OK;
}

/* ---- launch tables (pnol_functor_vtable) of a functor type ---- */
template <class F> struct ScalarTable {
	static int eval_batch_(const pnol_launch_env * env, const pnol_functor_params * P, const double * pts, long long B, int n, long long ld,
	                       const unsigned char * ind, double * f_out) { return eval_batch<F>(env, *P, pts, B, n, ld, ind, f_out); }
	static int fd_points_(const pnol_launch_env * env, const pnol_functor_params * P, const double * xfull, int nfull, const int * pos,
	                      const double * dx, int i0, int i1, double * fdx, double * f0) { return fd_points<F>(env, *P, xfull, nfull, pos, dx, i0, i1, fdx, f0); }
	static int fd_hessian_(const pnol_launch_env * env, const pnol_functor_params * P, const double * x, const double * dx, int n,
	                       const double * fdx, const double * f0, double * B) { return fd_hessian<F>(env, *P, x, dx, n, fdx, f0, B); }
	static int alpha_pool_(const pnol_launch_env * env, const pnol_functor_params * P, const double * xfull, const double * pfull,
	                       const unsigned char * is_const, int nfull, const double * alpha, int npool, double dalpha,
	                       const unsigned char * eval_ind, int want_shifted, double * vals)
	{ return alpha_pool<F>(env, *P, xfull, pfull, is_const, nfull, alpha, npool, dalpha, eval_ind, want_shifted, vals); }
	static pnol_functor_vtable make(int n_columns)
	{
		pnol_functor_vtable vt = {};
		vt.abi_version = PNOL_FUNCTOR_ABI;
		vt.n_columns = n_columns;
		vt.eval_batch = eval_batch_;
		vt.fd_points = fd_points_;
		vt.fd_hessian = fd_hessian_;
		vt.alpha_pool = alpha_pool_;
		return vt;
	}
};
template <class R> struct ResidualTable {
	static int residual_(const pnol_launch_env * env, const pnol_functor_params * P, const double * x, int n, double * F) { return residual<R>(env, *P, x, n, F); }
	static int fd_jacobian_(const pnol_launch_env * env, const pnol_functor_params * P, const double * x, const double * dx, int n, double * J,
	                        double * F) { return fd_jacobian<R>(env, *P, x, dx, n, J, F); }
	static pnol_functor_vtable make(int n_columns)
	{
		pnol_functor_vtable vt = {};
		vt.abi_version = PNOL_FUNCTOR_ABI;
		vt.n_columns = n_columns;
		vt.residual = residual_;
		vt.fd_jacobian = fd_jacobian_;
		return vt;
	}
};

} // namespace dev
} // namespace pnol

/* Registration at load time of the translation unit (one per functor). KIND: an integer constant in the user ranges of pnol_b200.h. */
#define PNOL_REGISTER_SCALAR_FUNCTOR(KIND, F, NCOLS)                                                              \
	static const pnol_functor_vtable pnol_vt_##F = pnol::dev::ScalarTable<F>::make(NCOLS);                        \
	static const int pnol_reg_##F = pnol_register_functor((KIND), &pnol_vt_##F);
#define PNOL_REGISTER_RESIDUAL_FUNCTOR(KIND, R, NCOLS)                                                            \
	static const pnol_functor_vtable pnol_vt_##R = pnol::dev::ResidualTable<R>::make(NCOLS);                      \
	static const int pnol_reg_##R = pnol_register_functor((KIND), &pnol_vt_##R);

#endif /* PNOL_DEVICE_FUNCTOR_KERNELS_CUH_ */
