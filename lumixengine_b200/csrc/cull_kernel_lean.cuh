// cull_pages_lean_kernel — the same cull as cull_kernel.cuh with fewer instructions per page (with the latency
// chain hidden behind neighbouring culls the kernel is instruction-issue-bound; profiles/r1_lanes_instruction_breakdown.txt).
//
//   STATUS: compiled into the library, selected with LB200_CULL_LEAN=1, NOT YET RUN ON A GPU (written after the round's GPU budget was
//   spent).  The default path is the validated kernel of cull_kernel.cuh.  Results must be bit-identical to it: every floating-point
//   expression that decides visibility or classification is the same; only control flow, work distribution and loads change.
//
//   A  two passes: A1 drops pages that fail the reference's intersectsAABB expression by a safe margin (which also rules out the shifted
//      containsAABB box); A2 runs the exact classification + plane mask on the compacted candidates only.
//   B  plane-outer / row-inner: warp-uniform loop over the planes of the mask, seven rows unrolled inside, no branch per sphere; lanes past
//      `count` load a clamped slot and are masked at the ballot.
//   D  pages whose ids are all visible (the reference's "fully inside" pages and pages with an empty plane mask) are copied with one base
//      address per lane and immediate row offsets; only tested pages go through ballots and ranks.
//
//   Static SASS (cuobjdump, sm_100a): 64 registers, 28 B of spills, 1888 instructions in all (default kernel: 2200).  The plane loop of B is
//   67 instructions per plane for the seven rows of a page (56 FMUL / FADD / LOP3 + 11 of loop overhead, plane load and shuffle), i.e.
//   about 45 + 67 x planes + 40 per tested page against the 353 the default kernel executes per tested page on the C2 view; a copied
//   page costs about 40 instructions in D against 131.  Estimate for the C2 view: 4.2 M warp instructions per cull against 7.2 M.
#pragma once

#include "cull_kernel.cuh"

namespace lbcull {

template <int CULL_THREADS>
__global__ void __launch_bounds__(CULL_THREADS, 1024 / CULL_THREADS) cull_pages_lean_kernel(const __grid_constant__ CullParams P,
	const lb200_page_desc* __restrict__ desc, const float4* __restrict__ spheres, const int* __restrict__ entities,
	uint32_t* __restrict__ out_ids, uint32_t* __restrict__ counters, uint32_t* __restrict__ next_counters, uint32_t* __restrict__ mask_out)
{
	constexpr int CULL_WARPS = CULL_THREADS / 32;
	constexpr int MAX_CHUNK = CULL_THREADS;
	__shared__ WorkItem s_item[MAX_CHUNK];
	__shared__ __align__(16) uint32_t s_bal[MAX_CHUNK][ROWS + 1]; // [ROWS] = visible count, then offset of the page inside its type's output segment
	__shared__ uint16_t s_slot[MAX_CHUNK]; // classify thread -> work slot of its page (SLOT_NONE: skipped / no page)
	__shared__ uint32_t s_stats[N_STATS];
	__shared__ uint32_t s_nwork;
	__shared__ uint16_t s_cand[MAX_CHUNK]; // classify threads whose page survived the cheap pass
	__shared__ uint32_t s_ncand;

	// let the next cull of the stream start its read-only prologue as soon as SM resources free up
	cudaTriggerProgrammaticLaunchCompletion();

	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int warp = tid >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;

	if (tid < N_STATS) s_stats[tid] = 0;
	if (tid == 0) { s_nwork = 0; s_ncand = 0; }
	__syncthreads();

	// Pages are dealt to blocks round-robin (page = j * gridDim + block): pages that always need sphere tests (is_big cells) and
	// frustum-boundary cells cluster in page-id space, and contiguous chunks left a few blocks with twice the work of the rest.
	for (uint32_t round = 0; round * P.chunk * gridDim.x < P.n_pages; ++round) {
		// ---------------- A1. cheap pass: one thread per page, "definitely outside" only ----------------
		// 73 % of the pages of a typical view are outside the frustum, but the exact classification (both cell tests + the plane mask) costs
		// ~900 instructions per warp of page threads whatever the outcome.  This pass evaluates only the reference's intersectsAABB
		// expression (bit-identical dp) and drops a page when it fails by a margin that also rules out the shifted containsAABB box
		// (the two boxes share the corner origin + cs up to rounding: margin >> that rounding); everything else —
		// is_big pages, pages near a plane, NaNs — goes to the exact pass, which then runs on densely packed threads.
		if ((uint32_t)tid < P.chunk) {
			const uint32_t page = (round * P.chunk + tid) * gridDim.x + blockIdx.x;
			s_slot[tid] = (uint16_t)SLOT_NONE;
			if (page < P.n_pages) {
				const int4* dp = reinterpret_cast<const int4*>(desc + page);
				const int4 a = __ldg(dp);
				const int4 b = __ldg(dp + 1);
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
						else s_cand[atomicAdd(&s_ncand, 1u)] = (uint16_t)tid;
					}
				}
			}
		}
		__syncthreads();
		// ---------------- A2. exact classification of the candidates (dense) ----------------
		for (uint32_t c = tid; c < s_ncand; c += CULL_THREADS) {
			const uint32_t t0 = s_cand[c];
			const uint32_t page = (round * P.chunk + t0) * gridDim.x + blockIdx.x;
			uint32_t my_slot = SLOT_NONE;
			{
				const int4* dp = reinterpret_cast<const int4*>(desc + page);
				const int4 a = __ldg(dp);
				const int4 b = __ldg(dp + 1);
				const double org_x = __hiloint2double(a.y, a.x);
				const double org_y = __hiloint2double(a.w, a.z);
				const double org_z = __hiloint2double(b.y, b.x);
				const uint32_t count = (uint32_t)b.z;
				const uint32_t type = (uint32_t)b.w & 0xffu;
				const bool is_big = (((uint32_t)b.w >> 8) & 0xffu) != 0;
				int cls = CLS_SKIP;
				if (count != 0) {
					if (P.type_filter == 0xffu || type == P.type_filter) {
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
					else atomicAdd(&s_stats[ST_PAGES_FILTERED], 1u);
				}
				uint32_t need = 0x3fu, as_test = 0;
				if (cls == CLS_TEST) {
					as_test = 1;
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
						const V3 offset = tofloat(sub(d3(P.ox, P.oy, P.oz), d3(org_x, org_y, org_z))); // getRelative, geometry.cpp:124
						need = 0;
#pragma unroll 1
						for (int p = 0; p < 6; ++p) {
							const float nx = P.nx[p], ny = P.ny[p], nz = P.nz[p];
							const float dp = -dot(add(v3(P.px[p], P.py[p], P.pz[p]), offset), v3(nx, ny, nz));
							const float low = dp + fminf(nx * lox, nx * hix) + fminf(ny * loy, ny * hiy) + fminf(nz * loz, nz * hiz);
							const float margin = 1e-5f * (fabsf(dp) + 1000.0f * (fabsf(nx) + fabsf(ny) + fabsf(nz))) + 1e-3f;
							if (!(low > margin)) need |= 1u << p; // NaN keeps the plane
						}
						if (need == 0) cls = CLS_COPY; // every sphere of the page is visible: ids only, no sphere traffic
					}
				}
				if (cls == CLS_TEST) prefetch_l2(spheres + (size_t)page * LB200_PAGE_SLOTS, count * 16u);
				if (cls != CLS_SKIP && (cls == CLS_COPY || P.prefetch_test_ids)) prefetch_l2(entities + (size_t)page * LB200_PAGE_SLOTS, (count * 4u + 15u) & ~15u);
				if (cls != CLS_SKIP) {
					const uint32_t slot = atomicAdd(&s_nwork, 1u);
					WorkItem it;
					it.ox = org_x; it.oy = org_y; it.oz = org_z;
					it.page = page;
					it.meta = count | (type << 8) | ((uint32_t)cls << 16) | (as_test << 18) | (need << 24);
					s_item[slot] = it;
					my_slot = slot;
				}
			}
			s_slot[t0] = (uint16_t)my_slot;
		}
		__syncthreads();
		const uint32_t n_work = s_nwork;

		// ---------------- B. test / copy: one warp per listed page ----------------
		for (uint32_t w = warp; w < n_work; w += CULL_WARPS) {
			const WorkItem it = s_item[w];
			const uint32_t count = it.meta & 0xffu;
			const int cls = (int)((it.meta >> 16) & 3u);
			const uint32_t as_test = (it.meta >> 18) & 1u;
			const uint32_t need = it.meta >> 24;
			uint32_t bal[ROWS];
			uint32_t page_visible = 0;
			if (cls == CLS_TEST) {
				float4 s[ROWS];
				const float4* sp = spheres + (size_t)it.page * LB200_PAGE_SLOTS;
				const uint32_t last = count - 1u; // count >= 1 for listed pages
#pragma unroll
				for (int k = 0; k < ROWS; ++k) {
					const uint32_t slot = k * 32 + lane;
					s[k] = ldg_stream(sp + (slot < last ? slot : last)); // lanes past the page re-read its last sphere: no predicate, masked at the ballot
				}
				// ShiftedFrustum::getRelative(cell.origin), geometry.cpp:121-149: offset = Vec3(this->origin - origin);
				// d = -dot(point + offset, normal) (setPlane, geometry.cpp:412-418); lane p < 6 computes plane p
				const int pl = lane < 6 ? lane : 0;
				const V3 offset = tofloat(sub(d3(P.ox, P.oy, P.oz), d3(it.ox, it.oy, it.oz)));
				const V3 pnt = add(v3(P.px[pl], P.py[pl], P.pz[pl]), offset);
				const float my_d = -dot(pnt, v3(P.nx[pl], P.ny[pl], P.nz[pl]));
				// doCulling, culling_system.cpp:260-308, plane-outer: the planes of the mask are walked by a warp-uniform loop, the rows are
				// unrolled inside with no branch per sphere; per sphere and plane exactly :284,291
				//   t = cx*px + cy*py + cz*pz + pd ;  t = t - r (r = -radius) ;  movemask = sign bits
				uint32_t acc[ROWS];
#pragma unroll
				for (int k = 0; k < ROWS; ++k) acc[k] = 0;
				for (uint32_t nb = need; nb; nb &= nb - 1u) {
					const int p = __ffs((int)nb) - 1;
					const float nx = P.nx[p], ny = P.ny[p], nz = P.nz[p];
					const float pd = __shfl_sync(0xffffffffu, my_d, p);
#pragma unroll
					for (int k = 0; k < ROWS; ++k) {
						float t = LB_FADD(LB_FADD(LB_FADD(LB_FMUL(s[k].x, nx), LB_FMUL(s[k].y, ny)), LB_FMUL(s[k].z, nz)), pd);
						t = LB_FSUB(t, -s[k].w); // :282 f4Splat(-sphere->radius)
						acc[k] |= __float_as_uint(t);
					}
				}
#pragma unroll
				for (int k = 0; k < ROWS; ++k) {
					// A NaN radius: on the reference's SSE path t - (-radius) hands the NaN through with the sign of -radius, and that sign is
					// what movemask reads (+NaN radius: culled by every plane; -NaN radius: passes every plane).  The GPU's subtraction returns
					// the canonical positive NaN instead, so the sign is taken from the radius directly.  (tests/golden/cull_kat.npz: special_*)
					const uint32_t rbits = __float_as_uint(s[k].w);
					if ((rbits & 0x7fffffffu) > 0x7f800000u && need) acc[k] = ~rbits & 0x80000000u;
					const bool visible = (acc[k] >> 31) == 0 && (uint32_t)(k * 32 + lane) < count;
					bal[k] = __ballot_sync(0xffffffffu, visible);
					page_visible += __popc(bal[k]);
				}
				if (lane == 0) {
					atomicAdd(&s_stats[ST_PAGES_TESTED], 1u);
					atomicAdd(&s_stats[ST_ENT_TESTED], count);
					atomicAdd(&s_stats[ST_ENT_STREAMED], count);
				}
			}
			else { // CLS_COPY, culling_system.cpp:345-360: every entity of the page is visible
#pragma unroll
				for (int k = 0; k < ROWS; ++k) {
					const int rem = (int)count - k * 32;
					bal[k] = rem >= 32 ? 0xffffffffu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
				}
				page_visible = count;
				if (lane == 0) { // statistics follow the reference's classification (culling_system.cpp:342-363), not the masking shortcut
					atomicAdd(&s_stats[as_test ? ST_PAGES_TESTED : ST_PAGES_INSIDE], 1u);
					atomicAdd(&s_stats[as_test ? ST_ENT_TESTED : ST_ENT_INSIDE], count);
				}
			}
			if (lane == 0) {
#pragma unroll
				for (int k = 0; k < ROWS; ++k) s_bal[w][k] = bal[k];
				s_bal[w][ROWS] = page_visible;
			}
		}
		// nothing above wrote global memory (A and B read scene data, results sit in shared memory); everything below does
		// (counters, ids, mask rows) and has to wait for the previous kernel of the stream
		if (round == 0) cudaGridDependencySynchronize();
		// ---------------- C. claim: one global atomic per (warp, type) — no block barrier between B, C and D ----------------
		// lane i stands for the warp's i-th page (w = warp + i * CULL_WARPS; at most 32 per warp since chunk <= CULL_THREADS)
		__syncwarp();
		{
			const uint32_t wi = warp + (uint32_t)lane * CULL_WARPS;
			const bool has = wi < n_work;
			const uint32_t my_type = has ? ((s_item[wi].meta >> 8) & 0xffu) : 0xffffffffu;
			const uint32_t my_count = has ? s_bal[wi][ROWS] : 0u;
			const uint32_t n_mine = (n_work + CULL_WARPS - 1 - warp) / CULL_WARPS; // pages of this warp (warp-uniform)
			uint32_t prefix = 0, total = 0;
			for (uint32_t l = 0; l < n_mine; ++l) {
				const uint32_t c = __shfl_sync(0xffffffffu, my_count, (int)l);
				const uint32_t t = __shfl_sync(0xffffffffu, my_type, (int)l);
				if (t == my_type) { total += c; if (l < (uint32_t)lane) prefix += c; }
			}
			const unsigned grp = __match_any_sync(0xffffffffu, my_type);
			const int leader = __ffs((int)grp) - 1;
			uint32_t base = 0;
			if (has && lane == leader && total) base = atomicAdd(&counters[my_type], total);
			base = __shfl_sync(0xffffffffu, base, leader);
			if (has) s_bal[wi][ROWS] = base + prefix; // offset of the page inside its type's output segment
		}
		__syncwarp();

		// ---------------- D. write: gather the visible ids of each listed page ----------------
		// Memory-level parallelism bounds this phase (Little's law at ~1 us loaded latency: 32 warps x 7 x 128 B per SM).  Batching two
		// or four pages per warp iteration was tried twice and lost to register spills under the 64-register cap of 4 blocks/SM.
		for (uint32_t w = warp; w < n_work; w += CULL_WARPS) {
			const uint32_t page = s_item[w].page;
			const uint32_t meta = s_item[w].meta;
			const uint32_t type = (meta >> 8) & 0xffu;
			uint32_t* dst = out_ids + P.type_base[type] + s_bal[w][ROWS];
			const int* ep = entities + (size_t)page * LB200_PAGE_SLOTS;
			if (((meta >> 16) & 3u) == CLS_COPY) {
				// every id of the page is visible (culling_system.cpp:345-360, or an empty plane mask): a straight copy, one base address per
				// lane and immediate offsets per row — no ballots, no ranks
				const uint32_t count = meta & 0xffu;
				const int* src = ep + lane;
				uint32_t* d = dst + lane;
				int id[ROWS];
#pragma unroll
				for (int k = 0; k < ROWS; ++k) if ((uint32_t)(k * 32 + lane) < count) id[k] = ldg_stream_i32(src + k * 32);
#pragma unroll
				for (int k = 0; k < ROWS; ++k) if ((uint32_t)(k * 32 + lane) < count) d[k * 32] = (uint32_t)id[k];
				continue;
			}
			uint32_t bal[ROWS];
#pragma unroll
			for (int k = 0; k < ROWS; ++k) bal[k] = s_bal[w][k];
			int id[ROWS];
#pragma unroll
			for (int k = 0; k < ROWS; ++k) if ((bal[k] >> lane) & 1u) id[k] = ldg_stream_i32(ep + k * 32 + lane);
			uint32_t prefix = 0;
#pragma unroll
			for (int k = 0; k < ROWS; ++k) {
				if ((bal[k] >> lane) & 1u) dst[prefix + __popc(bal[k] & lt_mask)] = (uint32_t)id[k];
				prefix += __popc(bal[k]);
			}
		}
		__syncthreads(); // every warp is done with s_item / s_bal
		// ---------------- E. mask rows of this round: one contiguous run per block, 128-bit stores ----------------
		if (mask_out || P.n_ranks) {
			const size_t run = ((size_t)blockIdx.x * P.rows_per_block + (size_t)round * P.chunk) * 2; // in uint4
			for (uint32_t i = tid; i < P.chunk * 2; i += CULL_THREADS) {
				const uint32_t slot = s_slot[i >> 1];
				uint4 v = make_uint4(0u, 0u, 0u, 0u);
				if (slot != SLOT_NONE) {
					v = *reinterpret_cast<const uint4*>(&s_bal[slot][(i & 1u) * 4]);
					if (i & 1u) v.w = 0u; // that word holds the output offset
				}
				if (P.n_ranks == 0) reinterpret_cast<uint4*>(mask_out)[run + i] = v;
				else {
					for (uint32_t r = 0; r < P.n_ranks; ++r) reinterpret_cast<uint4*>(P.xdst[r] + XHEADER_WORDS)[run + i] = v;
				}
			}
		}
		if (tid == 0) { s_nwork = 0; s_ncand = 0; }
		__syncthreads();
	}

	if (tid < N_STATS && s_stats[tid]) atomicAdd(&counters[256 + tid], s_stats[tid]);
	// the other counter buffer is the next cull's: zero it now so no memset sits between two culls
	if (blockIdx.x == 0) {
		for (int i = tid; i < COUNTER_WORDS; i += CULL_THREADS) next_counters[i] = 0;
	}
	// exchange mode: nothing more to do here.  The rows were stored without a fence; publish_wait_kernel (culling.cu), which runs
	// after this grid has completed, sends the per-type counts, fences once at system scope and raises the epoch flags.
}

} // namespace lbcull
