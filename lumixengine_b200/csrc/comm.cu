// Multi-GPU exchange of the visible lists (include/lumix_b200.h "Multi-GPU").
//
// The reference has no distributed code at all; the path shards by whole cell pages (each rank owns a subset of the
// entities and runs the same cull), and the only exchange step is the all-gather of the compacted visible lists.
// NCCL is resolved with dlopen so that single-GPU users never need it; inside a torch process the already-loaded
// libnccl.so.2 (torch's bundled copy) is the one that gets picked up.
#include "lb200_internal.h"

#include <dlfcn.h>

// minimal NCCL ABI (nccl.h): only what is called here
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8_dt = 1, ncclUint32_dt = 3 };

typedef ncclResult_t (*PFN_ncclGetUniqueId)(ncclUniqueId*);
typedef ncclResult_t (*PFN_ncclCommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
typedef ncclResult_t (*PFN_ncclCommDestroy)(ncclComm_t);
typedef ncclResult_t (*PFN_ncclAllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
typedef const char* (*PFN_ncclGetErrorString)(ncclResult_t);

static PFN_ncclGetUniqueId p_ncclGetUniqueId;
static PFN_ncclCommInitRank p_ncclCommInitRank;
static PFN_ncclCommDestroy p_ncclCommDestroy;
static PFN_ncclAllGather p_ncclAllGather;
static PFN_ncclGetErrorString p_ncclGetErrorString;


static int loadNccl(lb200_ctx* ctx) {
	if (ctx->nccl_lib) return LB200_OK;
	void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!lib) {
		lb200_set_error(ctx, "cannot dlopen libnccl.so.2: %s", dlerror());
		return LB200_ERR_NCCL;
	}
	p_ncclGetUniqueId = (PFN_ncclGetUniqueId)dlsym(lib, "ncclGetUniqueId");
	p_ncclCommInitRank = (PFN_ncclCommInitRank)dlsym(lib, "ncclCommInitRank");
	p_ncclCommDestroy = (PFN_ncclCommDestroy)dlsym(lib, "ncclCommDestroy");
	p_ncclAllGather = (PFN_ncclAllGather)dlsym(lib, "ncclAllGather");
	p_ncclGetErrorString = (PFN_ncclGetErrorString)dlsym(lib, "ncclGetErrorString");
	if (!p_ncclGetUniqueId || !p_ncclCommInitRank || !p_ncclCommDestroy || !p_ncclAllGather) {
		lb200_set_error(ctx, "libnccl lacks a required symbol");
		dlclose(lib);
		return LB200_ERR_NCCL;
	}
	ctx->nccl_lib = lib;
	return LB200_OK;
}

#define LB200_NCCL(ctx, expr)                                                                                  \
	do {                                                                                                       \
		ncclResult_t r__ = (expr);                                                                             \
		if (r__ != 0) {                                                                                        \
			lb200_set_error((ctx), "%s failed: %s", #expr, p_ncclGetErrorString ? p_ncclGetErrorString(r__) : "?"); \
			return LB200_ERR_NCCL;                                                                             \
		}                                                                                                      \
	} while (0)

// used by culling.cu: all-gather `words` u32 per rank on the context stream (asynchronous)
int lb200_comm_allgather_u32(lb200_ctx* ctx, const uint32_t* send, uint32_t* recv, size_t words) {
	if (!ctx->nccl_comm) { lb200_set_error(ctx, "lb200_comm_init has not been called"); return LB200_ERR_STATE; }
	LB200_NCCL(ctx, p_ncclAllGather(send, recv, words, ncclUint32_dt, (ncclComm_t)ctx->nccl_comm, ctx->stream));
	ctx->launches.fetch_add(1, std::memory_order_relaxed);
	return LB200_OK;
}

extern "C" {

int lb200_comm_get_unique_id(lb200_ctx* ctx, uint8_t out_id[128]) {
	if (!ctx || !out_id) return LB200_ERR_INVALID;
	int rc = loadNccl(ctx);
	if (rc) return rc;
	ncclUniqueId id;
	LB200_NCCL(ctx, p_ncclGetUniqueId(&id));
	memcpy(out_id, &id, 128);
	return LB200_OK;
}

int lb200_comm_init(lb200_ctx* ctx, int n_ranks, int rank, const uint8_t unique_id[128]) {
	if (!ctx || !unique_id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return LB200_ERR_INVALID;
	int rc = loadNccl(ctx);
	if (rc) return rc;
	LB200_CUDA(ctx, cudaSetDevice(ctx->device));
	ncclUniqueId id;
	memcpy(&id, unique_id, 128);
	ncclComm_t comm = nullptr;
	LB200_NCCL(ctx, p_ncclCommInitRank(&comm, n_ranks, id, rank));
	ctx->nccl_comm = comm;
	ctx->n_ranks = n_ranks;
	ctx->rank = rank;
	return LB200_OK;
}

void lb200_comm_destroy(lb200_ctx* ctx) {
	if (!ctx || !ctx->nccl_comm) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	p_ncclCommDestroy((ncclComm_t)ctx->nccl_comm);
	ctx->nccl_comm = nullptr;
	ctx->n_ranks = 1;
	ctx->rank = 0;
}

} // extern "C"
