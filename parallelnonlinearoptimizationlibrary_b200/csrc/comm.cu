// comm.cu -- NCCL communicator of a context (one process per GPU). Replaces the reference's MPI_COMM_WORLD
// collectives. NCCL is loaded with dlopen at first use so that single-GPU users need no NCCL:
// inside a torch process this resolves to the already-loaded torch-bundled libnccl.so.2.
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <mutex>

namespace pnol {

struct NcclApi {
	void * handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	const char * (*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
	std::string why;
};

static void nccl_api_load(NcclApi & api)
{
	const char * names[] = {"libnccl.so.2", "libnccl.so"};
	for (const char * nm : names) {
		api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
		if (api.handle) break;
	}
	if (!api.handle) { api.why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return; }
#define PNOL_SYM(field, name)                                                        \
	api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));      \
	if (!api.field) { api.why = std::string("missing NCCL symbol ") + name; return; }
	PNOL_SYM(GetUniqueId, "ncclGetUniqueId");
	PNOL_SYM(CommInitRank, "ncclCommInitRank");
	PNOL_SYM(CommDestroy, "ncclCommDestroy");
	PNOL_SYM(AllReduce, "ncclAllReduce");
	PNOL_SYM(AllGather, "ncclAllGather");
	PNOL_SYM(Broadcast, "ncclBroadcast");
	PNOL_SYM(GetErrorString, "ncclGetErrorString");
#undef PNOL_SYM
	api.ok = true;
}

// loaded once per process, whichever thread / context asks first
static NcclApi & nccl_api()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] { nccl_api_load(api); });
	return api;
}

#define PNOL_NCCL(ctx, call)                                                                             \
	do {                                                                                                 \
		ncclResult_t _r = (call);                                                                        \
		if (_r != ncclSuccess) {                                                                         \
			PNOL_SET_ERR(ctx, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, nccl_api().GetErrorString(_r)); \
			return PNOL_ERR_COMM;                                                                        \
		}                                                                                                \
	} while (0)

int comm_allreduce_dev(pnol_ctx * ctx, double * dev_buf, size_t count)
{
	if (ctx->nranks <= 1 || count == 0) return PNOL_OK;
	TimerScope ts(ctx, "allreduce");
	PNOL_NCCL(ctx, nccl_api().AllReduce(dev_buf, dev_buf, count, ncclDouble, ncclSum, (ncclComm_t) ctx->comm, ctx->stream));
	return PNOL_OK;
}

int comm_allgather_dev(pnol_ctx * ctx, const double * send, double * recv, size_t count_per_rank)
{
	if (count_per_rank == 0) return PNOL_OK;
	if (ctx->nranks <= 1) {
		if (send != recv)
			PNOL_CUDA(ctx, cudaMemcpyAsync(recv, send, count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
		return PNOL_OK;
	}
	TimerScope ts(ctx, "allgather");
	PNOL_NCCL(ctx, nccl_api().AllGather(send, recv, count_per_rank, ncclDouble, (ncclComm_t) ctx->comm, ctx->stream));
	return PNOL_OK;
}

int comm_allgather_bytes_dev(pnol_ctx * ctx, const void * send, void * recv, size_t bytes_per_rank)
{
	if (bytes_per_rank == 0) return PNOL_OK;
	if (ctx->nranks <= 1) {
		if (send != recv) PNOL_CUDA(ctx, cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
		return PNOL_OK;
	}
	TimerScope ts(ctx, "allgather");
	PNOL_NCCL(ctx, nccl_api().AllGather(send, recv, bytes_per_rank, ncclChar, (ncclComm_t) ctx->comm, ctx->stream));
	return PNOL_OK;
}

int comm_broadcast_dev(pnol_ctx * ctx, double * buf, size_t count, int root)
{
	if (ctx->nranks <= 1 || count == 0) return PNOL_OK;
	PNOL_NCCL(ctx, nccl_api().Broadcast(buf, buf, count, ncclDouble, root, (ncclComm_t) ctx->comm, ctx->stream));
	return PNOL_OK;
}

void comm_destroy(pnol_ctx * ctx)
{
	peer_destroy(ctx);                               // the peer mappings belong to this communicator's ranks
	if (ctx->comm && nccl_api().ok) nccl_api().CommDestroy((ncclComm_t) ctx->comm);
	ctx->comm = nullptr; ctx->nranks = 1; ctx->rank = 0;
	ctx->local = false; ctx->comm_nranks = 1; ctx->comm_rank = 0;
}

} // namespace pnol

using namespace pnol;

extern "C" int pnol_comm_unique_id(char id[PNOL_COMM_ID_BYTES])
{
	static_assert(sizeof(ncclUniqueId) <= PNOL_COMM_ID_BYTES, "unique id size");
	NcclApi & api = nccl_api();
	if (!api.ok) return PNOL_ERR_COMM;
	ncclUniqueId uid;
	if (api.GetUniqueId(&uid) != ncclSuccess) return PNOL_ERR_COMM;
	memset(id, 0, PNOL_COMM_ID_BYTES);
	memcpy(id, &uid, sizeof uid);
	return PNOL_OK;
}

extern "C" int pnol_comm_init(pnol_ctx * ctx, const char id[PNOL_COMM_ID_BYTES], int nranks, int rank)
{
	if (!ctx) return PNOL_ERR_INVALID;
	PNOL_REQUIRE(ctx, nranks >= 1 && rank >= 0 && rank < nranks, "comm_init: bad rank %d of %d", rank, nranks);
	NcclApi & api = nccl_api();
	if (!api.ok) { PNOL_SET_ERR(ctx, "NCCL unavailable: %s", api.why.c_str()); return PNOL_ERR_COMM; }
	PNOL_CUDA(ctx, cudaSetDevice(ctx->device));
	if (ctx->comm) {                                     // re-initialisation: the old communicator must not leak
		PNOL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		comm_destroy(ctx);
	}
	ncclUniqueId uid;
	memcpy(&uid, id, sizeof uid);
	ncclComm_t comm;
	PNOL_NCCL(ctx, api.CommInitRank(&comm, nranks, uid, rank));
	ctx->comm = (ncclComm *) comm;
	ctx->nranks = nranks;
	ctx->rank = rank;
	ctx->local = false; ctx->comm_nranks = nranks; ctx->comm_rank = rank;
	return PNOL_OK;
}

// Local (non-collective) mode. The reference's serial classes (BFGS, BFGS_Bnd, LevMarq, GeneticAlgorithm, SimplexSearch and the
// non-MPI stencil members of Objective) never touch MPI, so a program may call them on one rank only. While local mode is on this
// context behaves like a single-GPU context: no entry point issues a collective, a residual functor's rows are ALL the rows,
// pnol_comm_rank / pnol_comm_size answer 0 / 1. Returns the previous setting (for nesting).
extern "C" int pnol_comm_set_local(pnol_ctx * ctx, int on)
{
	if (!ctx) return 0;
	const int prev = ctx->local ? 1 : 0;
	if (on && !ctx->local) { ctx->local = true; ctx->nranks = 1; ctx->rank = 0; }
	else if (!on && ctx->local) { ctx->local = false; ctx->nranks = ctx->comm_nranks; ctx->rank = ctx->comm_rank; }
	return prev;
}

extern "C" int pnol_comm_rank(pnol_ctx * ctx) { return ctx ? ctx->rank : 0; }
extern "C" int pnol_comm_size(pnol_ctx * ctx) { return ctx ? ctx->nranks : 1; }

extern "C" int pnol_comm_allreduce_sum(pnol_ctx * ctx, double * buf, size_t count)
{
	if (!ctx) return PNOL_ERR_INVALID;
	if (ctx->nranks <= 1 || count == 0) return PNOL_OK;
	DevOut<double> d;
	PNOL_CHECK(d.init(ctx, buf, count, true));
	PNOL_CHECK(comm_allreduce_dev(ctx, d.get(), count));
	PNOL_CHECK(d.commit());
	return finish(ctx);
}

extern "C" int pnol_comm_allgather(pnol_ctx * ctx, const double * send, double * recv, size_t count_per_rank)
{
	if (!ctx) return PNOL_ERR_INVALID;
	DevIn<double> s;
	DevOut<double> r;
	PNOL_CHECK(s.init(ctx, send, count_per_rank));
	PNOL_CHECK(r.init(ctx, recv, count_per_rank * ctx->nranks));
	PNOL_CHECK(comm_allgather_dev(ctx, s.get(), r.get(), count_per_rank));
	PNOL_CHECK(r.commit());
	return finish(ctx);
}

extern "C" int pnol_comm_broadcast(pnol_ctx * ctx, double * buf, size_t count, int root)
{
	if (!ctx) return PNOL_ERR_INVALID;
	if (ctx->nranks <= 1 || count == 0) return PNOL_OK;
	DevOut<double> d;
	PNOL_CHECK(d.init(ctx, buf, count, true));
	PNOL_CHECK(comm_broadcast_dev(ctx, d.get(), count, root));
	PNOL_CHECK(d.commit());
	return finish(ctx);
}
