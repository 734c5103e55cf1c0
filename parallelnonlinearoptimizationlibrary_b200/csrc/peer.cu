// peer.cu -- the sharded Levenberg-Marquardt step's exchange over NVLink peer memory, fused with its consumers.
//
// Row-sharded LM sums the packed J^T J | J^T F of the ranks once per iteration and the trial chi^2 once more
// (the reference: MPI_Allreduce inside gradientApproximationMPI / LevMarqMPI::findMin, Source/LevenbergMarquardtMPI.cpp:60-108).
// Through NCCL that is two collective launches plus the damping kernel behind the first; at 8 GPUs an iteration is 1.7 ms and
// the n = 256 payload is 526 KB, so the cost is launch and protocol latency, not bandwidth. Here every rank publishes its partial
// in a buffer the other ranks have mapped (CUDA IPC), and ONE kernel per exchange does the rest:
//   * peer_reduce_damp_kernel: signals "my partial of epoch e is complete" into every peer's flag array (the kernel starts after
//     this rank's SYRK / J^T F kernels in stream order), waits for the peers' flags in LOCAL memory, then reads the R partials over
//     NVLink, adds them in RANK ORDER (the same order on every rank: bit-identical results on all ranks, and deterministic from
//     run to run) and writes J^T J, the damped matrix A and rhs = -J^T F -- the all-reduce and lm_damp_kernel in one launch;
//   * peer_scalar_sum_kernel: the same for one double (the trial sum of squares), values pushed into the peers' slots.
// Two slots alternate by epoch parity: a rank re-publishes slot s at epoch e + 2, after its epoch e + 1 kernel has seen every
// peer's e + 1 flag, which a peer raises only after its own epoch-e kernel is complete. Every spin has a clock limit; a rank whose
// peers never arrive reports PNOL_ERR_COMM instead of hanging the GPU. Without peer mappings (or PNOL_LM_PEER=0) the NCCL path is used.
#include "common.cuh"
#include "ga_common.cuh"      // comm_allgather_bytes_dev
#include "peer.cuh"

#include <stdlib.h>
#include <vector>

namespace pnol {

struct PnolPeer {
	int R = 0, me = 0;
	size_t cap = 0;                  // doubles per data slot
	double * own = nullptr;          // cudaMalloc'ed: data[2][cap] | mat_flag[2][kPeerMax] | sc_flag[2][kPeerMax] | sc_val[2][kPeerMax]
	void * maps[kPeerMax] = {};
	PeerTable table = {};
	unsigned long long ep_mat = 0, ep_sc = 0;
	int * err_host = nullptr;        // mapped pinned: a kernel that gave up waiting writes 1 here
	int * err_dev = nullptr;
	bool ok = false, tried = false;
};

// J^T J | J^T F = sum over the ranks' partials (rank order), then A = J^T J with (1 + lambda) on the diagonal, rhs = -J^T F
// (lm_damp_kernel of dense.cu, Source/LevenbergMarquardtMPI.cpp:66-85)
__global__ void __launch_bounds__(256)
peer_reduce_damp_kernel(PeerTable T, int R, int me, size_t cap, int slot, unsigned long long epoch, int n, double lambda,
                        const double * __restrict__ lambda_dev, double * __restrict__ JTJ, double * __restrict__ A, double * __restrict__ rhs,
                        int * __restrict__ err)
{
	__shared__ int s_bad;
	const size_t fl = 2 * cap;                                             // flags start here (in 8-byte words)
	if (threadIdx.x == 0) s_bad = 0;
	if (blockIdx.x == 0 && (int) threadIdx.x < R && (int) threadIdx.x != me)
		st_release_sys(reinterpret_cast<unsigned long long *>(T.base[threadIdx.x] + fl) + slot * kPeerMax + me, epoch);
	__syncthreads();
	if ((int) threadIdx.x < R && (int) threadIdx.x != me) {
		const unsigned long long * f = reinterpret_cast<const unsigned long long *>(T.base[me] + fl) + slot * kPeerMax + threadIdx.x;
		const long long t0 = clock64();
		while (ld_vol_u64(f) < epoch) {
			if (clock64() - t0 > kPeerSpinLimit) { s_bad = 1; break; }
			__nanosleep(64);
		}
	}
	__syncthreads();
	if (s_bad && threadIdx.x == 0) *err = 1;
	if (lambda_dev) lambda = *lambda_dev;
	const long long total = (long long) n * n;
	const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= total + n) return;
	const size_t off = (size_t) slot * cap + (size_t) idx;
	double v = ld_vol_f64(T.base[0] + off);
	for (int r = 1; r < R; r++) v = v + ld_vol_f64(T.base[r] + off);
	if (idx < total) {
		if (JTJ) JTJ[idx] = v;
		if (A) {
			const int i = (int) (idx / n), j = (int) (idx - (long long) i * n);
			A[idx] = (i == j) ? (1 + lambda) * v : v;
		}
	} else {
		if (JTJ) JTJ[idx] = -v;          // the right-hand side is kept behind J^T J in the caller's buffer (capi.cu: lm_step_enqueue)
		if (rhs) rhs[idx - total] = -v;
	}
}

// ss[0] = sum over the ranks of their ss[0], rank order
__global__ void __launch_bounds__(32)
peer_scalar_sum_kernel(PeerScalarArgs P, double * __restrict__ ss)
{
	const double total = peer_scalar_exchange_warp(P, ss[0]);
	if (threadIdx.x == 0) ss[0] = total;
}

static void peer_teardown(PnolPeer * P)
{
	for (int r = 0; r < kPeerMax; r++) if (P->maps[r]) { cudaIpcCloseMemHandle(P->maps[r]); P->maps[r] = nullptr; }
	if (P->own) { cudaFree(P->own); P->own = nullptr; }
	P->ok = false; P->cap = 0; P->R = 0;
}

void peer_destroy(pnol_ctx * ctx)
{
	PnolPeer * P = ctx->peer;
	if (!P) return;
	cudaStreamSynchronize(ctx->stream);
	peer_teardown(P);
	if (P->err_host) cudaFreeHost(P->err_host);
	delete P;
	ctx->peer = nullptr;
}

// COLLECTIVE (every rank of the communicator, same arguments): makes sure an exchange buffer of `count` doubles per slot exists and
// is mapped on every rank. false: use NCCL (decided identically on all ranks).
bool peer_ensure(pnol_ctx * ctx, size_t count)
{
	static const bool disabled = [] { const char * e = getenv("PNOL_LM_PEER"); return e && atoi(e) == 0; }();
	if (disabled || ctx->nranks <= 1 || ctx->nranks > kPeerMax) return false;
	if (!ctx->peer) ctx->peer = new PnolPeer();
	PnolPeer * P = ctx->peer;
	if (P->ok && P->R == ctx->nranks && P->me == ctx->rank && P->cap >= count) return true;
	if (P->tried && !P->ok) return false;                                // one attempt per context: every rank remembers the same outcome
	cudaStreamSynchronize(ctx->stream);
	peer_teardown(P);
	P->tried = true;
	const int R = ctx->nranks;
	const size_t cap = (count + 1023) & ~(size_t) 1023;
	const size_t words = 2 * cap + 6 * kPeerMax;
	int ok = 1;
	if (!P->err_host) {
		if (cudaHostAlloc((void **) &P->err_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); ok = 0; P->err_host = nullptr; }
		else { *P->err_host = 0; if (cudaHostGetDevicePointer((void **) &P->err_dev, P->err_host, 0) != cudaSuccess) { cudaGetLastError(); ok = 0; } }
	}
	if (ok && cudaMalloc((void **) &P->own, words * sizeof(double)) != cudaSuccess) { cudaGetLastError(); ok = 0; P->own = nullptr; }
	if (ok) cudaMemsetAsync(P->own, 0, words * sizeof(double), ctx->stream);
	cudaIpcMemHandle_t mine;
	memset(&mine, 0, sizeof mine);
	if (ok && cudaIpcGetMemHandle(&mine, P->own) != cudaSuccess) { cudaGetLastError(); ok = 0; }
	// everybody takes the same path: exchange [ok | handle] through the communicator
	const size_t slot = 8 + sizeof(cudaIpcMemHandle_t);
	std::vector<unsigned char> send(slot, 0), recv(slot * R, 0);
	send[0] = (unsigned char) ok;
	memcpy(send.data() + 8, &mine, sizeof mine);
	unsigned char * dsend = nullptr, * drecv = nullptr;
	bool good = cudaMalloc((void **) &dsend, slot) == cudaSuccess && cudaMalloc((void **) &drecv, slot * R) == cudaSuccess;
	if (good) {
		cudaMemcpyAsync(dsend, send.data(), slot, cudaMemcpyHostToDevice, ctx->stream);
		good = comm_allgather_bytes_dev(ctx, dsend, drecv, slot) == PNOL_OK;
		cudaMemcpyAsync(recv.data(), drecv, slot * R, cudaMemcpyDeviceToHost, ctx->stream);
		cudaStreamSynchronize(ctx->stream);
	}
	if (dsend) cudaFree(dsend);
	if (drecv) cudaFree(drecv);
	bool opened = good;
	for (int r = 0; r < R && opened; r++) if (!recv[(size_t) r * slot]) opened = false;
	if (opened) {
		for (int r = 0; r < R; r++) {
			if (r == ctx->rank) { P->table.base[r] = P->own; continue; }
			cudaIpcMemHandle_t h;
			memcpy(&h, recv.data() + (size_t) r * slot + 8, sizeof h);
			void * p = nullptr;
			if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = false; p = nullptr; }
			P->maps[r] = p;
			P->table.base[r] = (double *) p;
		}
	}
	// a common decision again (one failed mapping anywhere sends every rank to NCCL); it is also the barrier behind which every
	// rank's buffer is zeroed before anybody raises a flag in it
	double flag = opened ? 0.0 : 1.0;
	double * dflag = nullptr;
	if (cudaMalloc((void **) &dflag, sizeof(double)) != cudaSuccess) { peer_teardown(P); return false; }
	cudaMemcpyAsync(dflag, &flag, sizeof flag, cudaMemcpyHostToDevice, ctx->stream);
	const bool reduced = comm_allreduce_dev(ctx, dflag, 1) == PNOL_OK;
	cudaMemcpyAsync(&flag, dflag, sizeof flag, cudaMemcpyDeviceToHost, ctx->stream);
	cudaStreamSynchronize(ctx->stream);
	cudaFree(dflag);
	if (!reduced || flag != 0.0) { peer_teardown(P); return false; }
	P->R = R; P->me = ctx->rank; P->cap = cap; P->ep_mat = 0; P->ep_sc = 0; P->ok = true;
	return true;
}

// where this rank's producers (SYRK finish, J^T F) write the partial of the NEXT exchange
double * peer_partial_slot(pnol_ctx * ctx)
{
	PnolPeer * P = ctx->peer;
	return P->own + (size_t) ((P->ep_mat + 1) & 1ULL) * P->cap;
}

int launch_peer_reduce_damp(pnol_ctx * ctx, int n, double lambda, const double * lambda_dev, double * JTJ, double * A, double * rhs)
{
	PnolPeer * P = ctx->peer;
	TimerScope ts(ctx, "allreduce");
	P->ep_mat++;
	const long long total = (long long) n * n + n;
	PNOL_LAUNCH(ctx, peer_reduce_damp_kernel, (unsigned) ((total + 255) / 256), 256, 0, P->table, P->R, P->me, P->cap, (int) (P->ep_mat & 1ULL),
	            P->ep_mat, n, lambda, lambda_dev, JTJ, A, rhs, P->err_dev);
	return PNOL_OK;
}

// the arguments of the NEXT scalar exchange (advances the epoch: the caller must launch exactly one kernel that performs it)
PeerScalarArgs peer_scalar_next(pnol_ctx * ctx)
{
	PnolPeer * P = ctx->peer;
	P->ep_sc++;
	PeerScalarArgs a;
	a.T = P->table; a.R = P->R; a.me = P->me; a.slot = (int) (P->ep_sc & 1ULL); a.cap = P->cap; a.epoch = P->ep_sc; a.err = P->err_dev;
	return a;
}

int launch_peer_scalar_sum(pnol_ctx * ctx, double * ss)
{
	PNOL_LAUNCH(ctx, peer_scalar_sum_kernel, 1, 32, 0, peer_scalar_next(ctx), ss);
	return PNOL_OK;
}

} // namespace pnol

// which way the sharded LM step's sums travel on this context: 0 one rank (no exchange), 1 NVLink peer memory (the fused kernels
// above), 2 NCCL all-reduce, -1 not decided yet (the decision is taken collectively by the first sharded pnol_lm_step / pnol_lm_iterate)
extern "C" int pnol_lm_exchange_mode(pnol_ctx * ctx)
{
	if (!ctx) return -1;
	if (ctx->nranks <= 1) return 0;
	const pnol::PnolPeer * P = ctx->peer;
	if (P && P->ok) return 1;
	if (P && P->tried) return 2;
	static const bool disabled = [] { const char * e = getenv("PNOL_LM_PEER"); return e && atoi(e) == 0; }();
	return disabled ? 2 : -1;
}

namespace pnol {

// after a synchronisation: did a kernel of this context give up waiting for its peers?
int peer_check(pnol_ctx * ctx)
{
	PnolPeer * P = ctx->peer;
	if (P && P->err_host && *P->err_host) {
		*P->err_host = 0;
		PNOL_SET_ERR(ctx, "peer exchange: a rank did not publish its partial in time (ranks out of step?)");
		return PNOL_ERR_COMM;
	}
	return PNOL_OK;
}

} // namespace pnol
