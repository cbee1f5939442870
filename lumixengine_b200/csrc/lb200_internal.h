// Internal declarations shared by the translation units of liblumix_b200.so.
#pragma once

#include "../../include/lumix_b200.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only; a no-op unless a tool (ncu, nsys) injects itself

#include <atomic>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#define LB200_MAX_RANKS 8
#define LB200_CULL_STAGE_DEFAULT 0 // pages in flight per warp through the bulk-copy engine (cull_kernel.cuh); LB200_CULL_STAGE overrides
#define LB200_MAX_LANES 8  // concurrent culls (streams / output lanes); exchange buffers = 3 x lanes

struct lb200_ctx {
	int device = -1;
	cudaStream_t stream = nullptr;
	cudaStream_t copy_stream = nullptr;
	int sm_count = 0;
	std::atomic<uint64_t> launches{0};
	char error[512] = {0};
	// NCCL (dlopen) state, see comm.cu
	void* nccl_lib = nullptr;
	void* nccl_comm = nullptr;
	int n_ranks = 1;
	int rank = 0;
	// NVLink peer exchange (comm.cu lb200_comm_enable_p2p): every rank's gather buffers mapped into every process
	struct Peer {
		bool ready = false;
		size_t slab_words = 0;            // capacity of one rank's slab (header + ids)
		// Exchange epoch e uses buffer e % n_buffers, n_buffers = 3 x lanes.  With lanes > 1 (lb200_culling_cull_exchange_n) epoch e is
		// issued on stream e % lanes as ONE kernel: the cull of e publishes the lane's previous epoch (e - lanes) from its prologue and
		// holds its record stores back until every rank has published e - 2 x lanes.  So a rank overwrites buffer b for epoch e only after
		// every rank published e - 2 x lanes, which a rank does from its cull of e - lanes — issued behind whatever consumed e - 3 x lanes,
		// the previous owner of b.  tests/test_exchange_protocol_model.py replays this (and the two-kernel forms culling.cu keeps as
		// options) under a random scheduler, and shows that fewer buffers or no flow control would not do.
		uint32_t lanes = 1, n_buffers = 3;
		void* local_block = nullptr;      // this rank's allocation: [flags n_buffers x 8 x u32 in 512 B][gather 0] .. [gather n_buffers-1]
		uint32_t* gather[3 * LB200_MAX_LANES][LB200_MAX_RANKS] = {}; // gather[b][r] = rank r's buffer b as seen from this process
		uint32_t* flags[LB200_MAX_RANKS] = {};     // flags[r] = rank r's flag block
		void* opened[LB200_MAX_RANKS] = {};        // cudaIpcOpenMemHandle results to close
		uint32_t* done_counter = nullptr; // local, one per lane, for the last-block election
		uint32_t epoch = 0;
		// a wait kernel that gave up on a peer (~4 s) raises this word; page-locked + mapped so the host sees it without a copy.
		// lb200_comm_check() turns it into LB200_ERR_NCCL and resets it (called by lb200_synchronize and every exchange entry point)
		uint32_t* h_timeout = nullptr;
		uint32_t* d_timeout = nullptr;
	} peer;
};

void lb200_set_error(lb200_ctx* ctx, const char* fmt, ...);

// NVTX range named like the reference's PROFILE_BLOCK / PROFILE_FUNCTION scopes (culling_system.cpp:330 "culling",
// animation_module.cpp:743 "update animables"), so that a timeline of the engine with this library reads like the reference's own.
struct lb200_range {
	explicit lb200_range(const char* name) { nvtxRangePushA(name); }
	~lb200_range() { nvtxRangePop(); }
	lb200_range(const lb200_range&) = delete;
	lb200_range& operator=(const lb200_range&) = delete;
};
// culling.cu: where the last cull left its result (device: ids, counters; host: per-type segment bases and entity counts, 256 each)
int lb200_culling_internal_last(lb200_culling* cs, const uint32_t** out_ids, const uint32_t** counters, const uint32_t** type_base, const uint32_t** type_counts);
// sortkeys.cu: stable LSD radix sort of (u64 key, u64 value) pairs, count read on the device; the result ends in buffer 0
size_t lb200_radix_sort_state_bytes();
int lb200_radix_sort_pairs(lb200_ctx* ctx, cudaStream_t stream, uint64_t* keys0, uint64_t* keys1, uint64_t* values0, uint64_t* values1, const uint32_t* count_dev, uint32_t cap,
	void* state, uint32_t* block_hist, uint32_t blocks);
int lb200_comm_check(lb200_ctx* ctx); // comm.cu: LB200_ERR_NCCL (and reset) if a peer wait timed out since the last check
uint32_t lb200_cull_lanes(); // LB200_CULL_LANES, default 3, 1..LB200_MAX_LANES (context.cu)

#define LB200_CUDA(ctx, expr)                                                                        \
	do {                                                                                             \
		cudaError_t e__ = (expr);                                                                    \
		if (e__ != cudaSuccess) {                                                                    \
			lb200_set_error((ctx), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
			return LB200_ERR_CUDA;                                                                   \
		}                                                                                            \
	} while (0)

#define LB200_CHECK_LAUNCH(ctx)                                                                      \
	do {                                                                                             \
		(ctx)->launches.fetch_add(1, std::memory_order_relaxed);                                     \
		cudaError_t e__ = cudaGetLastError();                                                        \
		if (e__ != cudaSuccess) {                                                                    \
			lb200_set_error((ctx), "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
			return LB200_ERR_CUDA;                                                                   \
		}                                                                                            \
	} while (0)

// Device page layout: page p owns slots [p*200, p*200+200) of the sphere / entity arrays.
struct alignas(32) lb200_page_desc {
	double origin[3]; // CellPage::header.origin, culling_system.cpp:55
	uint32_t count;   // header.count (0 = free page, skipped by the kernel)
	uint8_t type;     // header.indices.type
	uint8_t is_big;   // header.indices.is_big
	uint16_t pad;
};
static_assert(sizeof(lb200_page_desc) == 32, "page descriptor is one 32-byte sector");
static_assert(sizeof(lb200_shifted_frustum) == 256, "ShiftedFrustum image, geometry.h:99-149");
static_assert(sizeof(lb200_transform) == 56, "Transform image, math.h:306-327");
static_assert(sizeof(lb200_track) == 32, "track descriptor");
static_assert(sizeof(lb200_sk_model) == 64 && sizeof(lb200_sk_mesh) == 16 && sizeof(lb200_sk_view) == 1352, "sort-key tables");
