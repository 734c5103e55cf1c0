// peer.cuh -- device side of the NVLink peer-memory exchange (peer.cu): the table of mapped buffers and the one-double exchange
// that other kernels fuse (capi.cu: lm_tail_kernel = sum of squares + exchange + accept / reject in one launch).
#pragma once
#include <cuda_runtime.h>

struct pnol_ctx;

namespace pnol {

constexpr int kPeerMax = 16;
constexpr long long kPeerSpinLimit = 4000000000LL;      // clock cycles (about 2 s)

struct PeerTable { double * base[kPeerMax]; };

// what a kernel needs for ONE scalar exchange (peer_scalar_next hands them out and advances the epoch); R == 0: single rank, no exchange
struct PeerScalarArgs {
	PeerTable T;
	int R, me, slot;
	size_t cap;
	unsigned long long epoch;
	int * err;
};

__device__ __forceinline__ unsigned long long ld_vol_u64(const unsigned long long * p) { return *(volatile const unsigned long long *) p; }
__device__ __forceinline__ double ld_vol_f64(const double * p) { return *(volatile const double *) p; }
__device__ __forceinline__ void st_release_sys(unsigned long long * p, unsigned long long v)
{
	__threadfence_system();
	*(volatile unsigned long long *) p = v;
}

// Called by ONE FULL WARP (all 32 lanes, converged): every rank pushes `mine` into slot [me] of every rank's value array, raises its
// flag there, waits for the R flags in its own memory and returns the sum of the R values in RANK ORDER (the same bits on every
// rank) in all lanes. A rank whose peers never arrive writes 1 to *err and returns what it has.
__device__ __forceinline__ double peer_scalar_exchange_warp(const PeerScalarArgs & P, double mine)
{
	const int lane = threadIdx.x & 31;
	const size_t fl = 2 * P.cap + 2 * kPeerMax;                            // sc_flag, then sc_val
	bool bad = false;
	double v = 0.0;
	if (lane < P.R) {
		double * pv = P.T.base[lane] + fl + 2 * kPeerMax + P.slot * kPeerMax + P.me;
		*(volatile double *) pv = mine;
		st_release_sys(reinterpret_cast<unsigned long long *>(P.T.base[lane] + fl) + P.slot * kPeerMax + P.me, P.epoch);
		const unsigned long long * f = reinterpret_cast<const unsigned long long *>(P.T.base[P.me] + fl) + P.slot * kPeerMax + lane;
		const long long t0 = clock64();
		while (ld_vol_u64(f) < P.epoch) {
			if (clock64() - t0 > kPeerSpinLimit) { bad = true; break; }
			__nanosleep(64);
		}
		v = ld_vol_f64(P.T.base[P.me] + fl + 2 * kPeerMax + P.slot * kPeerMax + lane);
	}
	if (__any_sync(0xffffffffu, bad) && lane == 0) *P.err = 1;
	double total = __shfl_sync(0xffffffffu, v, 0);
	for (int r = 1; r < P.R; r++) total = total + __shfl_sync(0xffffffffu, v, r);
	return total;
}

// host side (peer.cu): the arguments of the NEXT scalar exchange; advances the epoch, so exactly one kernel performing it must follow
PeerScalarArgs peer_scalar_next(pnol_ctx * ctx);

} // namespace pnol
