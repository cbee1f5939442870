// CullingSystemImpl::cullInternal + doCulling (src/renderer/culling_system.cpp:260-369) on the GPU, as two kernels per cull:
//
//   cull_classify_kernel  one THREAD per cell page, contiguous pages per block (coalesced 32-byte descriptors).
//        A1  cheap pass: free pages, the type filter, and "definitely outside" — the reference's intersectsAABB expression
//            (geometry.cpp:159-178, bit-identical dp) failing by a safe margin, which also rules out the shifted containsAABB box.
//            ~3/4 of the pages of a typical view end here.
//        A2  exact pass on the compacted survivors: the cell tests of culling_system.cpp:342-363 (is_big -> test;
//            containsAABB(origin + cs, cs) -> copy every id; intersectsAABB(origin - cs, 2cs) -> test; else nothing), the plane mask,
//            and for pages that need sphere tests the plane offsets re-based to the cell origin (ShiftedFrustum::getRelative,
//            geometry.cpp:121-149 — only d changes).
//        Output: two dense work lists in global memory — TEST items {page, count, type, planes needed, 6 re-based d} and COPY items
//            {page, count, type, destination offset}: the output range of a copied page is known here, so it is claimed here (one global
//            atomic per block and type).  L2 prefetches of the sphere / id rows of every listed page are issued on the spot, so the work
//            kernel finds them in L2.  The per-page statistics of the reference's classification are counted here too.
//   cull_work_kernel      persistent grid (SMs x resident blocks), one WARP per work item, items dealt round-robin over all warps of the
//        grid: perfectly balanced whatever the view looks like, and skipped pages never reach a warp.
//        T   test items: <=200 spheres streamed with 128-bit loads, the planes of the mask walked by a warp-uniform loop with the rows
//            unrolled inside (no branch per sphere), the reference's op order and sign-bit test (culling_system.cpp:284-295, simd.h:119);
//            ballots kept in shared memory; output space claimed with one global atomic per (block, type);
//        C   copy items: ids streamed to their pre-claimed range (runs while the claims of T are in flight);
//        W   visible ids of the tested pages gathered and written compacted behind the block's claim.
//        The 32-byte visibility row of every worked page goes to the mask (row = page id; rows of untouched pages were zeroed by the
//        classify kernel), or — exchange mode — as {page id, row} records straight into every rank's slab over NVLink: only rows
//        that can be non-zero cross the links (the bitmask is the exchanged product).
//
// Both kernels release their dependents at once (programmatic dependent launch): the classify kernel of the next cull of the stream
// computes while this cull's work kernel drains, and writes nothing before its cudaGridDependencySynchronize().
// HBM-bound: 32 B descriptor per page + 16 B per tested sphere + 4 B read + 4 B write per visible id + 32 B mask row per page.
#pragma once

#include "lb200_internal.h"
#include "lb200_math.cuh"

namespace lbcull {

using namespace lb;

constexpr int ROWS = 7;                 // ceil(200 / 32)
constexpr int N_STATS = 8;
enum { ST_PAGES_TESTED = 0, ST_PAGES_INSIDE, ST_PAGES_OUTSIDE, ST_PAGES_FILTERED, ST_ENT_TESTED, ST_ENT_INSIDE, ST_ENT_STREAMED };
// counters of one cull: [0,256) visible per type, [256,264) statistics, [264] TEST items, [265] COPY items
constexpr int CNT_N_TEST = 256 + N_STATS;
constexpr int CNT_N_COPY = CNT_N_TEST + 1;
constexpr int COUNTER_WORDS = 256 + N_STATS + 8;

constexpr int CLASSIFY_THREADS = 128;   // pages per classify block
constexpr int WORK_THREADS = 256;
constexpr int WORK_WARPS = WORK_THREADS / 32;
constexpr int MAX_T = 2;                // test items per warp per round (their rows are staged in shared memory)
constexpr int MAX_C = 2;                // copy items per warp per round

struct CullParams {
	// planes NEAR, FAR, LEFT, RIGHT, TOP, BOTTOM of the ShiftedFrustum (relative to `origin`)
	float nx[6], ny[6], nz[6], d[6];
	// the frustum point each plane is re-anchored on by getRelative (geometry.cpp:134-142): points[0,4,1,0,0,2]
	float px[6], py[6], pz[6];
	double ox, oy, oz;
	uint32_t n_pages;
	uint32_t type_filter;   // 0xff = all
	uint32_t plane_masking; // 1 unless some sphere has a negative / NaN radius
	uint32_t item_cap;      // capacity of each work list and of an exchange slab's record area (>= n_pages)
	uint32_t trace;         // profiling: stamp phase boundaries into g_trace
	// exchange mode (n_ranks > 0): {page, row} records go straight into every rank's slab (peer memory)
	uint32_t n_ranks;
	uint32_t* xdst[LB200_MAX_RANKS]; // rank r's exchange buffer of this epoch, already offset to MY slab inside it
	uint32_t type_base[256];
};
// exchange slab = [256 per-type counts][n_pages, n_test, n_copy, item_cap, 0, 0, 0, 0][page ids: item_cap][rows: item_cap x 8]
// record i < n_test: the i-th TEST item; record item_cap - 1 - j: the j-th COPY item
constexpr uint32_t XHEADER_WORDS = 264;

struct TestItem { // 32 B
	uint32_t page;
	uint32_t meta; // count | type << 8 | planes needed << 16
	float rd[6];   // plane offsets relative to the cell origin
};
struct CopyItem { // 8 B
	uint32_t page;
	uint32_t meta; // count | type << 8
};
static_assert(sizeof(TestItem) == 32 && sizeof(CopyItem) == 8, "");

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
	float4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
	return r;
}

__device__ __forceinline__ int ldg_stream_i32(const int* p) {
	int r;
	asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
	return r;
}

// Ask the L2 to fetch the 128-byte line at `p`: nothing to wait on.  (One line per lane; the bulk form cp.async.bulk.prefetch.L2 takes
// uniform operands, which costs a serialising loop over the lanes when every thread has its own page.)
__device__ __forceinline__ void prefetch_l2_line(const void* p) {
	asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

enum { CLS_SKIP = 0, CLS_COPY = 1, CLS_TEST = 2 };

// Profiling aid (LB200_CULL_TRACE=1, lb200_culling_read_trace): thread 0 of every block stamps %globaltimer at the phase boundaries.
constexpr int TRACE_BLOCKS = 2048, TRACE_POINTS = 8;
__device__ unsigned long long g_trace[2][TRACE_BLOCKS][TRACE_POINTS];
__device__ __forceinline__ void trace_point(uint32_t on, int kernel, int point) {
	if (on && threadIdx.x == 0 && blockIdx.x < TRACE_BLOCKS) {
		unsigned long long t;
		asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
		g_trace[kernel][blockIdx.x][point] = t;
	}
}

__global__ void __launch_bounds__(CLASSIFY_THREADS) cull_classify_kernel(const __grid_constant__ CullParams P,
	const lb200_page_desc* __restrict__ desc, const float4* __restrict__ spheres, const int* __restrict__ entities,
	uint32_t* __restrict__ counters, TestItem* __restrict__ test_items, CopyItem* __restrict__ copy_items, uint32_t* __restrict__ mask_out)
{
	__shared__ int4 s_desc[CLASSIFY_THREADS][2];
	__shared__ uint16_t s_cand[CLASSIFY_THREADS];
	__shared__ uint32_t s_stats[N_STATS];
	__shared__ uint32_t s_work[CLASSIFY_THREADS]; // listed pages of this block: page - page0 | count << 8 | is TEST << 16
	__shared__ uint32_t s_ncand, s_ntest, s_ncopy, s_test_base, s_copy_base;

	// the work kernel of this cull may be scheduled right away (it waits for this grid at its cudaGridDependencySynchronize)
	cudaTriggerProgrammaticLaunchCompletion();

	const int tid = threadIdx.x;
	const int lane = tid & 31;
	trace_point(P.trace, 0, 0);
	if (tid < N_STATS) s_stats[tid] = 0;
	if (tid == 0) { s_ncand = 0; s_ntest = 0; s_ncopy = 0; }
	__syncthreads();

	// ---------------- A1. cheap pass: "definitely outside" only ----------------
	const uint32_t page0 = blockIdx.x * CLASSIFY_THREADS;
	{
		const uint32_t page = page0 + tid;
		bool cand = false;
		if (page < P.n_pages) {
			const int4* dp = reinterpret_cast<const int4*>(desc + page);
			const int4 a = __ldg(dp);
			const int4 b = __ldg(dp + 1);
			s_desc[tid][0] = a;
			s_desc[tid][1] = b;
			const uint32_t count = (uint32_t)b.z;
			const uint32_t type = (uint32_t)b.w & 0xffu;
			const bool is_big = (((uint32_t)b.w >> 8) & 0xffu) != 0;
			if (count != 0) {
				if (P.type_filter != 0xffu && type != P.type_filter) atomicAdd(&s_stats[ST_PAGES_FILTERED], 1u);
				else {
					bool outside = false;
					if (!is_big) {
						const double org_x = __hiloint2double(a.y, a.x);
						const double org_y = __hiloint2double(a.w, a.z);
						const double org_z = __hiloint2double(b.y, b.x);
						const float cs = LB200_CELL_SIZE;
						const float cs2 = 2 * LB200_CELL_SIZE;
						const V3 rel_i = tofloat(sub(d3(LB_DSUB(org_x, (double)cs), LB_DSUB(org_y, (double)cs), LB_DSUB(org_z, (double)cs)), d3(P.ox, P.oy, P.oz)));
						const V3 max_i = add(rel_i, v3(cs2, cs2, cs2));
#pragma unroll
						for (int p = 0; p < 6; ++p) {
							const float nx = P.nx[p], ny = P.ny[p], nz = P.nz[p], nd = -P.d[p];
							const float tx = LB_FMUL(nx, nx > 0.0f ? max_i.x : rel_i.x);
							const float ty = LB_FMUL(ny, ny > 0.0f ? max_i.y : rel_i.y);
							const float tz = LB_FMUL(nz, nz > 0.0f ? max_i.z : rel_i.z);
							const float dp_i = LB_FADD(LB_FADD(tx, ty), tz); // the exact pass computes the same value
							const float margin = 1e-4f * (fabsf(nd) + fabsf(tx) + fabsf(ty) + fabsf(tz)) + 0.05f;
							if (dp_i + margin < nd) outside = true; // NaN anywhere: false, the page stays a candidate
						}
					}
					if (outside) atomicAdd(&s_stats[ST_PAGES_OUTSIDE], 1u);
					else cand = true;
				}
			}
		}
		// warp-aggregated append to the candidate list
		const uint32_t bal = __ballot_sync(0xffffffffu, cand);
		uint32_t base = 0;
		if (lane == 0 && bal) base = atomicAdd(&s_ncand, (uint32_t)__popc(bal));
		base = __shfl_sync(0xffffffffu, base, 0);
		if (cand) s_cand[base + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)tid;
	}
	__syncthreads();
	trace_point(P.trace, 0, 1);

	// ---------------- A2. exact classification of the candidates (dense threads) ----------------
	int cls = CLS_SKIP;
	uint32_t page = 0, meta = 0, in_list = 0;
	float rd[6];
	if ((uint32_t)tid < s_ncand) {
		const uint32_t t0 = s_cand[tid];
		page = page0 + t0;
		const int4 a = s_desc[t0][0];
		const int4 b = s_desc[t0][1];
		const double org_x = __hiloint2double(a.y, a.x);
		const double org_y = __hiloint2double(a.w, a.z);
		const double org_z = __hiloint2double(b.y, b.x);
		const uint32_t count = (uint32_t)b.z;
		const uint32_t type = (uint32_t)b.w & 0xffu;
		const bool is_big = (((uint32_t)b.w >> 8) & 0xffu) != 0;
		{
			// containsAABB(cell.origin + Vec3(cs), Vec3(cs)), geometry.cpp:99-118 (DVec3 + Vec3: math.cpp:512)
			const float cs = LB200_CELL_SIZE;
			const V3 rel_c = tofloat(sub(d3(LB_DADD(org_x, (double)cs), LB_DADD(org_y, (double)cs), LB_DADD(org_z, (double)cs)), d3(P.ox, P.oy, P.oz)));
			const V3 max_c = add(rel_c, v3(cs, cs, cs));
			// intersectsAABB(cell.origin - Vec3(cs), Vec3(2cs)), geometry.cpp:159-178 (DVec3 - Vec3: math.cpp:510)
			const float cs2 = 2 * LB200_CELL_SIZE;
			const V3 rel_i = tofloat(sub(d3(LB_DSUB(org_x, (double)cs), LB_DSUB(org_y, (double)cs), LB_DSUB(org_z, (double)cs)), d3(P.ox, P.oy, P.oz)));
			const V3 max_i = add(rel_i, v3(cs2, cs2, cs2));
			bool contains = true, intersects = true;
#pragma unroll
			for (int p = 0; p < 6; ++p) {
				const float nx = P.nx[p], ny = P.ny[p], nz = P.nz[p], nd = -P.d[p];
				const float cbx = nx < 0.0f ? max_c.x : rel_c.x;
				const float cby = ny < 0.0f ? max_c.y : rel_c.y;
				const float cbz = nz < 0.0f ? max_c.z : rel_c.z;
				const float dp_c = LB_FADD(LB_FADD(LB_FMUL(nx, cbx), LB_FMUL(ny, cby)), LB_FMUL(nz, cbz));
				if (dp_c < nd) contains = false;
				const float ibx = nx > 0.0f ? max_i.x : rel_i.x;
				const float iby = ny > 0.0f ? max_i.y : rel_i.y;
				const float ibz = nz > 0.0f ? max_i.z : rel_i.z;
				const float dp_i = LB_FADD(LB_FADD(LB_FMUL(nx, ibx), LB_FMUL(ny, iby)), LB_FMUL(nz, ibz));
				if (dp_i < nd) intersects = false;
			}
			// culling_system.cpp:342-363
			if (is_big) cls = CLS_TEST;
			else if (contains) cls = CLS_COPY;
			else if (intersects) cls = CLS_TEST;
			else atomicAdd(&s_stats[ST_PAGES_OUTSIDE], 1u);
		}
		uint32_t need = 0x3fu;
		// statistics follow the reference's classification (culling_system.cpp:342-363), not the masking shortcut below
		if (cls == CLS_TEST) { atomicAdd(&s_stats[ST_PAGES_TESTED], 1u); atomicAdd(&s_stats[ST_ENT_TESTED], count); }
		else if (cls == CLS_COPY) { atomicAdd(&s_stats[ST_PAGES_INSIDE], 1u); atomicAdd(&s_stats[ST_ENT_INSIDE], count); }
		if (cls == CLS_TEST) {
			// ShiftedFrustum::getRelative(cell.origin), geometry.cpp:121-149: offset = Vec3(this->origin - origin);
			// d = -dot(point + offset, normal) (setPlane, geometry.cpp:412-418)
			const V3 offset = tofloat(sub(d3(P.ox, P.oy, P.oz), d3(org_x, org_y, org_z)));
#pragma unroll
			for (int p = 0; p < 6; ++p) rd[p] = -dot(add(v3(P.px[p], P.py[p], P.pz[p]), offset), v3(P.nx[p], P.ny[p], P.nz[p]));
			if (P.plane_masking) {
				// Plane masking: a plane cannot cull any sphere of this cell when its signed distance is positive over the whole cell box
				// by more than every rounding error of the reference's expression — then sign(t - r) is 0 for every sphere (radius >= 0)
				// and evaluating the plane changes nothing.  Cell box relative to the cell origin: [0,300] for positive cell indices,
				// [-300,0] for negative ones, [-300,300] for index 0 (truncation toward zero, math.cpp:133-138), widened by `e` because
				// the cell index comes from pos * float(1/300) and may put a sphere marginally outside its nominal cell.
				const float cs = LB200_CELL_SIZE;
				const float e = 1.0f + 1e-6f * fmaxf(fmaxf(fabsf((float)org_x), fabsf((float)org_y)), fabsf((float)org_z));
				const float lox = (org_x > 0.0 ? 0.0f : -cs) - e, hix = (org_x < 0.0 ? 0.0f : cs) + e;
				const float loy = (org_y > 0.0 ? 0.0f : -cs) - e, hiy = (org_y < 0.0 ? 0.0f : cs) + e;
				const float loz = (org_z > 0.0 ? 0.0f : -cs) - e, hiz = (org_z < 0.0 ? 0.0f : cs) + e;
				need = 0;
#pragma unroll
				for (int p = 0; p < 6; ++p) {
					const float nx = P.nx[p], ny = P.ny[p], nz = P.nz[p];
					const float dp = rd[p];
					const float low = dp + fminf(nx * lox, nx * hix) + fminf(ny * loy, ny * hiy) + fminf(nz * loz, nz * hiz);
					const float margin = 1e-5f * (fabsf(dp) + 1000.0f * (fabsf(nx) + fabsf(ny) + fabsf(nz))) + 1e-3f;
					if (!(low > margin)) need |= 1u << p; // NaN keeps the plane
				}
				if (need == 0) cls = CLS_COPY; // every sphere of the page is visible: ids only, no sphere traffic
			}
		}
		if (cls == CLS_TEST) {
			atomicAdd(&s_stats[ST_ENT_STREAMED], count);
			in_list = atomicAdd(&s_ntest, 1u);
			s_work[in_list] = t0 | (count << 8) | (1u << 16); // TEST pages from the front
		}
		if (cls == CLS_COPY) {
			in_list = atomicAdd(&s_ncopy, 1u);
			s_work[CLASSIFY_THREADS - 1 - in_list] = t0 | (count << 8); // COPY pages from the back
		}
		meta = count | (type << 8) | (need << 16);
	}
	__syncthreads();
	trace_point(P.trace, 0, 2);
	// everything above only READ scene data; everything below writes buffers of this output lane, which the previous cull of the
	// lane (and whatever consumed it) may still be using
	cudaGridDependencySynchronize();
	trace_point(P.trace, 0, 3);
	// ---------------- claim list slots: ONE returning atomic per block (both list lengths in one 64-bit word) ----------------
	const uint32_t nt = s_ntest, nc = s_ncopy;
	if (tid == 0 && (nt | nc)) {
		const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(counters + CNT_N_TEST), (unsigned long long)nt | ((unsigned long long)nc << 32));
		s_test_base = (uint32_t)old;
		s_copy_base = (uint32_t)(old >> 32);
	}
	if (tid >= 64 && tid < 64 + N_STATS && s_stats[tid - 64]) atomicAdd(&counters[256 + tid - 64], s_stats[tid - 64]);
	// rows of pages without work stay zero; the work kernel (ordered behind this grid) writes the others
	if (mask_out && page0 + tid < P.n_pages) {
		uint4* row = reinterpret_cast<uint4*>(mask_out + (size_t)(page0 + tid) * 8);
		row[0] = make_uint4(0u, 0u, 0u, 0u);
		row[1] = make_uint4(0u, 0u, 0u, 0u);
	}
	// L2 prefetch of everything the work kernel will read, while the claim is in flight: one warp instruction per listed page
	// (lanes 0-24: the 25 lines of the sphere rows of a TEST page, lanes 25-31: the 7 lines of the id rows)
	for (uint32_t w = tid >> 5; w < nt + nc; w += CLASSIFY_THREADS / 32) {
		const uint32_t e = s_work[w < nt ? w : CLASSIFY_THREADS - 1 - (w - nt)];
		const size_t slot0 = (size_t)(page0 + (e & 0xffu)) * LB200_PAGE_SLOTS;
		const uint32_t count = (e >> 8) & 0xffu;
		if (lane < 25) { if ((e >> 16) && (uint32_t)lane * 8u < count) prefetch_l2_line(reinterpret_cast<const char*>(spheres + slot0) + lane * 128); }
		else if ((uint32_t)(lane - 25) * 32u < count) prefetch_l2_line(reinterpret_cast<const char*>(entities + slot0) + (lane - 25) * 128);
	}
	trace_point(P.trace, 0, 4);
	__syncthreads();
	trace_point(P.trace, 0, 5);
	if (cls == CLS_TEST) {
		uint4* it = reinterpret_cast<uint4*>(test_items + s_test_base + in_list);
		it[0] = make_uint4(page, meta, __float_as_uint(rd[0]), __float_as_uint(rd[1]));
		it[1] = make_uint4(__float_as_uint(rd[2]), __float_as_uint(rd[3]), __float_as_uint(rd[4]), __float_as_uint(rd[5]));
	}
	else if (cls == CLS_COPY) *reinterpret_cast<uint2*>(copy_items + s_copy_base + in_list) = make_uint2(page, meta);
	trace_point(P.trace, 0, 6);
}

// ---- shared-memory staging with the bulk-copy engine (TMA, 1-D form) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	uint32_t done;
	do {
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
	} while (!done);
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is counted on the mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// per-warp staging area: the sphere rows and id rows of the warp's TEST items and the id rows of its COPY items of one round
struct WarpStage {
	float4 sph[MAX_T][LB200_PAGE_SLOTS]; // 3200 B each
	int tid_[MAX_T][LB200_PAGE_SLOTS];   // 800 B each
	int cid[MAX_C][LB200_PAGE_SLOTS];
};
static_assert(sizeof(WarpStage) % 16 == 0, "");
constexpr size_t WORK_SMEM = sizeof(WarpStage) * WORK_WARPS;

__global__ void __launch_bounds__(WORK_THREADS) cull_work_kernel(const __grid_constant__ CullParams P,
	const float4* __restrict__ spheres, const int* __restrict__ entities, const TestItem* test_items, const CopyItem* copy_items,
	uint32_t* __restrict__ out_ids, uint32_t* counters, uint32_t* __restrict__ next_counters, uint32_t* __restrict__ mask_out)
{
	extern __shared__ __align__(128) unsigned char s_dyn[];
	__shared__ uint32_t s_cnt[256];  // ids this block will write per type (one round): visible ones of tested pages + all of copied pages
	__shared__ uint32_t s_base[256]; // the block's claim inside the type's segment
	__shared__ __align__(16) uint32_t s_bal[WORK_WARPS][MAX_T][8]; // 7 ballots + the page's offset inside the block's claim
	__shared__ __align__(8) uint64_t s_bar[WORK_WARPS][3];         // per warp: spheres landed / copy ids landed / test ids landed

	// the classify kernel of the next cull of this stream may start computing now
	cudaTriggerProgrammaticLaunchCompletion();

	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int warp = tid >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	WarpStage& st = reinterpret_cast<WarpStage*>(s_dyn)[warp];
	const uint32_t bar_s = smem_u32(&s_bar[warp][0]), bar_c = smem_u32(&s_bar[warp][1]), bar_t = smem_u32(&s_bar[warp][2]);
	s_cnt[tid] = 0; // WORK_THREADS == 256
	if (lane == 0) { mbar_init(bar_s, 1); mbar_init(bar_c, 1); mbar_init(bar_t, 1); }
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	trace_point(P.trace, 1, 0);

	cudaGridDependencySynchronize(); // the work lists and their lengths
	trace_point(P.trace, 1, 1);
	const uint32_t n_test = __ldcg(counters + CNT_N_TEST);
	const uint32_t n_copy = __ldcg(counters + CNT_N_COPY);
	const uint32_t total_warps = gridDim.x * WORK_WARPS;
	const uint32_t gw = (uint32_t)warp * gridDim.x + blockIdx.x; // neighbouring items go to different SMs
	uint32_t phase = 0; // parity of the three barriers (they are armed together, once per round)

	for (uint32_t round = 0; round * MAX_T * total_warps < n_test || round * MAX_C * total_warps < n_copy; ++round) {
		// ---------------- headers of this warp's items of the round; all their rows requested at once ----------------
		// lane j < MAX_T holds TEST item j, lane MAX_T + j holds COPY item j
		uint4 ia = make_uint4(0u, 0u, 0u, 0u), ib = make_uint4(0u, 0u, 0u, 0u);
		{
			const uint32_t it = (round * MAX_T + lane) * total_warps + gw;
			const uint32_t ic = (round * MAX_C + (lane - MAX_T)) * total_warps + gw;
			if (lane < MAX_T) {
				if (it < n_test) { ia = __ldcg(reinterpret_cast<const uint4*>(test_items + it)); ib = __ldcg(reinterpret_cast<const uint4*>(test_items + it) + 1); }
			}
			else if (lane < MAX_T + MAX_C && ic < n_copy) {
				const uint2 c = __ldcg(reinterpret_cast<const uint2*>(copy_items + ic));
				ia.x = c.x; ia.y = c.y;
			}
		}
		const uint32_t my_count = ia.y & 0xffu; // 0: no item in this lane
		const uint32_t has = __ballot_sync(0xffffffffu, my_count != 0);
		const uint32_t my_nt = __popc(has & ((1u << MAX_T) - 1u));           // items are dense from j = 0
		const uint32_t my_nc = __popc(has >> MAX_T);
		{
			const uint32_t sph_bytes = lane < MAX_T ? my_count * 16u : 0u;
			const uint32_t id_bytes = (my_count * 4u + 15u) & ~15u;
			uint32_t tot_s = sph_bytes, tot_t = lane < MAX_T ? id_bytes : 0u, tot_c = lane >= MAX_T ? id_bytes : 0u;
#pragma unroll
			for (int d = 1; d < MAX_T + MAX_C; d <<= 1) {
				tot_s += __shfl_xor_sync(0xffffffffu, tot_s, d);
				tot_t += __shfl_xor_sync(0xffffffffu, tot_t, d);
				tot_c += __shfl_xor_sync(0xffffffffu, tot_c, d);
			}
			if (lane == 0) {
				if (my_nt) { mbar_expect_tx(bar_s, tot_s); mbar_expect_tx(bar_t, tot_t); }
				if (my_nc) mbar_expect_tx(bar_c, tot_c);
			}
			__syncwarp();
			if (my_count) {
				const size_t slot0 = (size_t)ia.x * LB200_PAGE_SLOTS;
				if (lane < MAX_T) {
					bulk_load(smem_u32(&st.sph[lane][0]), spheres + slot0, sph_bytes, bar_s);
					bulk_load(smem_u32(&st.tid_[lane][0]), entities + slot0, id_bytes, bar_t);
				}
				else bulk_load(smem_u32(&st.cid[lane - MAX_T][0]), entities + slot0, id_bytes, bar_c);
			}
		}
		__syncthreads(); // s_cnt is zero
		// ---------------- T. sphere tests (rows come from shared memory) ----------------
		if (my_nt) mbar_wait(bar_s, phase);
		for (uint32_t j = 0; j < my_nt; ++j) {
			const uint32_t i = (round * MAX_T + j) * total_warps + gw;
			const uint32_t page = __shfl_sync(0xffffffffu, ia.x, (int)j);
			const uint32_t meta = __shfl_sync(0xffffffffu, ia.y, (int)j);
			const float rd0 = __uint_as_float(__shfl_sync(0xffffffffu, ia.z, (int)j)), rd1 = __uint_as_float(__shfl_sync(0xffffffffu, ia.w, (int)j));
			const float rd2 = __uint_as_float(__shfl_sync(0xffffffffu, ib.x, (int)j)), rd3 = __uint_as_float(__shfl_sync(0xffffffffu, ib.y, (int)j));
			const float rd4 = __uint_as_float(__shfl_sync(0xffffffffu, ib.z, (int)j)), rd5 = __uint_as_float(__shfl_sync(0xffffffffu, ib.w, (int)j));
			const uint32_t count = meta & 0xffu;
			const uint32_t type = (meta >> 8) & 0xffu;
			const uint32_t need = (meta >> 16) & 0x3fu;
			const bool upper = count > 128u; // rows 4-6 exist (warp-uniform): half of the tested pages of a typical scene stop before
			float4 s[ROWS];
#pragma unroll
			for (int k = 0; k < 4; ++k) s[k] = st.sph[j][k * 32 + lane]; // slots past `count` hold stale rows: masked at the ballot
			if (upper) {
#pragma unroll
				for (int k = 4; k < ROWS; ++k) { const int slot = k * 32 + lane; s[k] = st.sph[j][slot < LB200_PAGE_SLOTS ? slot : LB200_PAGE_SLOTS - 1]; }
			}
			else {
#pragma unroll
				for (int k = 4; k < ROWS; ++k) s[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			}
			// doCulling, culling_system.cpp:260-308, plane-outer: per sphere and plane exactly :284,291
			//   t = cx*px + cy*py + cz*pz + pd ;  t = t - r (r = -radius) ;  movemask = sign bits
			uint32_t acc[ROWS];
#pragma unroll
			for (int k = 0; k < ROWS; ++k) acc[k] = 0;
#define LB_ROWS(p, pd, k0, k1)                                                                                                   \
				_Pragma("unroll") for (int k = k0; k < k1; ++k) {                                                                 \
					float t = LB_FADD(LB_FADD(LB_FADD(LB_FMUL(s[k].x, nx), LB_FMUL(s[k].y, ny)), LB_FMUL(s[k].z, nz)), pd);       \
					t = LB_FSUB(t, -s[k].w); /* :282 f4Splat(-sphere->radius) */                                                  \
					acc[k] |= __float_as_uint(t);                                                                                 \
				}
#define LB_PLANE(p, pd)                                                                                                          \
			if (need & (1u << p)) {                                                                                               \
				const float nx = P.nx[p], ny = P.ny[p], nz = P.nz[p];                                                             \
				LB_ROWS(p, pd, 0, 4)                                                                                              \
				if (upper) { LB_ROWS(p, pd, 4, ROWS) }                                                                            \
			}
			LB_PLANE(0, rd0)
			LB_PLANE(1, rd1)
			LB_PLANE(2, rd2)
			LB_PLANE(3, rd3)
			LB_PLANE(4, rd4)
			LB_PLANE(5, rd5)
#undef LB_PLANE
#undef LB_ROWS
			uint32_t bal[ROWS];
			uint32_t page_visible = 0;
#pragma unroll
			for (int k = 0; k < ROWS; ++k) {
				// A NaN radius: on the reference's SSE path t - (-radius) hands the NaN through with the sign of -radius, and that sign is
				// what movemask reads (+NaN radius: culled by every plane; -NaN radius: passes every plane).  The GPU's subtraction returns
				// the canonical positive NaN instead, so the sign is taken from the radius directly.  (tests/golden/cull_kat.npz: special_*)
				// (The complement goes through inline PTX: the compiler otherwise rewrites "sign of ~bits" as neg.f32 + a sign test, and
				// neg.f32 of a NaN does not keep the sign.)
				// Only scenes that hold a negative / NaN radius get here (the host tracks them and switches plane masking off with them).
				if (!P.plane_masking) {
					const uint32_t rbits = __float_as_uint(s[k].w);
					uint32_t flipped;
					asm("not.b32 %0, %1;" : "=r"(flipped) : "r"(rbits));
					if ((rbits & 0x7fffffffu) > 0x7f800000u && need) acc[k] = flipped & 0x80000000u;
				}
				const bool visible = (acc[k] >> 31) == 0 && (uint32_t)(k * 32 + lane) < count;
				bal[k] = __ballot_sync(0xffffffffu, visible);
				page_visible += __popc(bal[k]);
			}
			if (lane == 0) {
				const uint32_t off = atomicAdd(&s_cnt[type], page_visible);
				*reinterpret_cast<uint4*>(&s_bal[warp][j][0]) = make_uint4(bal[0], bal[1], bal[2], bal[3]);
				*reinterpret_cast<uint4*>(&s_bal[warp][j][4]) = make_uint4(bal[4], bal[5], bal[6], off);
				if (mask_out) {
					uint4* row = reinterpret_cast<uint4*>(mask_out + (size_t)page * 8);
					row[0] = make_uint4(bal[0], bal[1], bal[2], bal[3]);
					row[1] = make_uint4(bal[4], bal[5], bal[6], 0u);
				}
			}
			if ((uint32_t)lane < P.n_ranks) { // exchange mode: record i of my slab in rank `lane`'s memory
				uint32_t* slab = P.xdst[lane];
				slab[XHEADER_WORDS + i] = page;
				uint4* row = reinterpret_cast<uint4*>(slab + XHEADER_WORDS + P.item_cap + (size_t)i * 8);
				row[0] = make_uint4(bal[0], bal[1], bal[2], bal[3]);
				row[1] = make_uint4(bal[4], bal[5], bal[6], 0u);
			}
		}
		// copied pages join the block's claim with their full count (culling_system.cpp:345-360: every entity of the page is visible)
		uint32_t copy_off = 0;
		if (lane >= MAX_T && my_count) copy_off = atomicAdd(&s_cnt[(ia.y >> 8) & 0xffu], my_count);
		trace_point(P.trace, 1, 2);
		__syncthreads();
		// ---------------- claim: one global atomic per (block, type) ----------------
		if (s_cnt[tid]) s_base[tid] = atomicAdd(&counters[tid], s_cnt[tid]);
		trace_point(P.trace, 1, 3);
		__syncthreads(); // s_base is there
		trace_point(P.trace, 1, 4);
		// ---------------- C. copy items: ids from shared memory to their range ----------------
		if (my_nc) mbar_wait(bar_c, phase);
		for (uint32_t j = 0; j < my_nc; ++j) {
			const uint32_t i = (round * MAX_C + j) * total_warps + gw;
			const uint32_t page = __shfl_sync(0xffffffffu, ia.x, (int)(MAX_T + j));
			const uint32_t meta = __shfl_sync(0xffffffffu, ia.y, (int)(MAX_T + j));
			const uint32_t off = __shfl_sync(0xffffffffu, copy_off, (int)(MAX_T + j));
			const uint32_t count = meta & 0xffu;
			const uint32_t type = (meta >> 8) & 0xffu;
			uint32_t* dst = out_ids + P.type_base[type] + s_base[type] + off;
#pragma unroll
			for (int k = 0; k < ROWS; ++k) if ((uint32_t)(k * 32 + lane) < count) dst[k * 32 + lane] = (uint32_t)st.cid[j][k * 32 + lane];
			if (mask_out || P.n_ranks) { // lane k < 8 holds word k of the row: all ones up to `count`
				const int rem = (int)count - (lane & 7) * 32;
				const uint32_t w = rem >= 32 ? 0xffffffffu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
				if (mask_out && lane < 8) mask_out[(size_t)page * 8 + lane] = w;
				if (P.n_ranks) {
					const uint32_t rec = P.item_cap - 1u - i;
					for (uint32_t r = (uint32_t)lane >> 3; r < P.n_ranks; r += 4) { // 8 lanes per destination rank
						uint32_t* slab = P.xdst[r];
						slab[XHEADER_WORDS + P.item_cap + (size_t)rec * 8 + (lane & 7)] = w;
						if ((lane & 7) == 0) slab[XHEADER_WORDS + rec] = page;
					}
				}
			}
		}
		// ---------------- W. visible ids of the tested pages ----------------
		if (my_nt) mbar_wait(bar_t, phase);
		for (uint32_t j = 0; j < my_nt; ++j) {
			const uint4 b0 = *reinterpret_cast<const uint4*>(&s_bal[warp][j][0]);
			const uint4 b1 = *reinterpret_cast<const uint4*>(&s_bal[warp][j][4]);
			const uint32_t bal[ROWS] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z};
			const uint32_t type = (__shfl_sync(0xffffffffu, ia.y, (int)j) >> 8) & 0xffu;
			uint32_t* dst = out_ids + P.type_base[type] + s_base[type] + b1.w;
			uint32_t prefix = 0;
#pragma unroll
			for (int k = 0; k < ROWS; ++k) {
				if ((bal[k] >> lane) & 1u) dst[prefix + __popc(bal[k] & lt_mask)] = (uint32_t)st.tid_[j][k * 32 + lane];
				prefix += __popc(bal[k]);
			}
		}
		phase ^= 1u; // a warp arms its barriers in every round from the first up to its last one with items: parity = round & 1
		if ((round + 1) * MAX_T * total_warps < n_test || (round + 1) * MAX_C * total_warps < n_copy) {
			__syncthreads(); // every warp is done with s_cnt / s_base
			s_cnt[tid] = 0;
		}
	}
	trace_point(P.trace, 1, 5);
	// the other counter buffer is the lane's next cull's: zero it now so no memset sits between two culls
	if (blockIdx.x == 0) {
		for (int i = tid; i < COUNTER_WORDS; i += WORK_THREADS) next_counters[i] = 0;
	}
	// exchange mode: nothing more to do here.  The records were stored without a fence; publish_wait_kernel (culling.cu), which runs
	// after this grid has completed, sends the header, fences once at system scope and raises the epoch flags.
}

} // namespace lbcull
