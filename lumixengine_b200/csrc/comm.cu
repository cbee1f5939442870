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

// from culling.cu
struct lb200_culling;
lb200_ctx* lb200_culling_ctx(lb200_culling* cs);
const uint32_t* lb200_culling_dev_ids(lb200_culling* cs);
const lb200_cull_result* lb200_culling_last_result(lb200_culling* cs);
uint32_t** lb200_culling_gather_ids_slot(lb200_culling* cs, size_t** cap);
uint32_t** lb200_culling_gather_counts_slot(lb200_culling* cs);
uint32_t** lb200_culling_slab_slot(lb200_culling* cs, size_t** cap);

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

int lb200_culling_allgather(lb200_culling* cs, uint32_t slab_ids, const uint32_t** out_dev_ids, uint32_t* out_counts) {
	if (!cs) return LB200_ERR_INVALID;
	lb200_ctx* ctx = lb200_culling_ctx(cs);
	if (!ctx) return LB200_ERR_NO_DEVICE;
	if (!ctx->nccl_comm) { lb200_set_error(ctx, "lb200_comm_init has not been called"); return LB200_ERR_STATE; }
	const lb200_cull_result* last = lb200_culling_last_result(cs);
	if (!last) { lb200_set_error(ctx, "allgather needs a preceding cull with counts"); return LB200_ERR_STATE; }
	const int R = ctx->n_ranks;
	LB200_CUDA(ctx, cudaSetDevice(ctx->device));

	size_t* ids_cap = nullptr;
	uint32_t** d_ids = lb200_culling_gather_ids_slot(cs, &ids_cap);
	uint32_t** d_counts = lb200_culling_gather_counts_slot(cs);
	size_t* slab_cap = nullptr;
	uint32_t** d_slab = lb200_culling_slab_slot(cs, &slab_cap);
	if (!*d_counts) LB200_CUDA(ctx, cudaMalloc(d_counts, sizeof(uint32_t) * 256 * (size_t)(R + 1)));
	if (*ids_cap < (size_t)slab_ids * R) {
		cudaFree(*d_ids);
		*d_ids = nullptr;
		LB200_CUDA(ctx, cudaMalloc(d_ids, sizeof(uint32_t) * (size_t)slab_ids * R));
		*ids_cap = (size_t)slab_ids * R;
	}
	if (*slab_cap < slab_ids) {
		cudaFree(*d_slab);
		*d_slab = nullptr;
		LB200_CUDA(ctx, cudaMalloc(d_slab, sizeof(uint32_t) * (size_t)slab_ids));
		*slab_cap = slab_ids;
	}
	// pack this rank's per-type segments contiguously into the send slab
	uint32_t off = 0;
	const uint32_t* src = lb200_culling_dev_ids(cs);
	uint32_t packed_counts[256];
	for (int t = 0; t < 256; ++t) {
		packed_counts[t] = last->type_count[t];
		if (!last->type_count[t]) continue;
		if (off + last->type_count[t] > slab_ids) { lb200_set_error(ctx, "slab_ids too small"); return LB200_ERR_CAPACITY; }
		LB200_CUDA(ctx, cudaMemcpyAsync(*d_slab + off, src + last->type_offset[t], sizeof(uint32_t) * last->type_count[t], cudaMemcpyDeviceToDevice, ctx->stream));
		off += last->type_count[t];
	}
	uint32_t* d_my_counts = *d_counts + 256 * (size_t)R;
	LB200_CUDA(ctx, cudaMemcpyAsync(d_my_counts, packed_counts, sizeof(packed_counts), cudaMemcpyHostToDevice, ctx->stream));
	LB200_NCCL(ctx, p_ncclAllGather(d_my_counts, *d_counts, 256, ncclUint32_dt, (ncclComm_t)ctx->nccl_comm, ctx->stream));
	LB200_NCCL(ctx, p_ncclAllGather(*d_slab, *d_ids, slab_ids, ncclUint32_dt, (ncclComm_t)ctx->nccl_comm, ctx->stream));
	ctx->launches.fetch_add(2, std::memory_order_relaxed);
	if (out_counts) {
		LB200_CUDA(ctx, cudaMemcpyAsync(out_counts, *d_counts, sizeof(uint32_t) * 256 * (size_t)R, cudaMemcpyDeviceToHost, ctx->stream));
	}
	LB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (out_dev_ids) *out_dev_ids = *d_ids;
	return LB200_OK;
}

} // extern "C"
