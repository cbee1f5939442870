// Consumer of the visible list on the device: PipelineImpl::createSortKeys (src/renderer/pipeline.cpp:3789-4018) —
// LOD selection with its smoothing state, sort keys / sort values (:53-143), auto-instancing groups + their instance data (:452-523,
// :3958-4016) — and PipelineImpl::radixSort (:4020-4144).  The ids the cull kernel compacted never leave HBM: this stage reads them
// where they lie (lb200_culling's per-type segments + counters) and leaves sorted keys / values and per-group instance data in HBM; the
// host reads back four counters.
//
//   emit_kernel        one thread per visible renderable (grid-stride over the MESH, DECAL and CURVE_DECAL segments):
//                      MESH  — squared distance to the LOD reference point in fp64 -> float, Model::getLODMeshIndices (model.h:173-179),
//                              the lod smoothing of :3926-3941 (ModelInstance::lod is updated in place), then per mesh of the LOD range(s)
//                              one of: skinned -> key/value + the instance joins the pose list once per frame (the compare-exchange on
//                              Pose::frame, :3890-3897); moved -> key/value; bucket < 0xff -> a record for the auto-instancer (its rank
//                              inside the group comes from an atomic on the group's counter); depth-sorted -> depth key/value;
//                      DECAL / CURVE_DECAL — key/value from the material's sort key and layer.
//   groups_kernel      exclusive scan of the group counters (one block), one key/value per non-empty group (:3958-3969).
//   fill_kernel        one thread per record: 48 bytes of instance data (rot, camera-relative position, lod - mesh.lod, scale, material
//                      index) at group_offset + rank (:3990-4008).
//   radix sort         LSD, 8 passes of 8 bits over the 64-bit keys, stable, hand-written: one global histogram kernel decides which
//                      passes move anything (the reference skips passes whose keys share one bin too, :4120), then per pass
//                      block histograms -> scan -> stable scatter (warp match + per-warp counts).  No library sort.
// The reference runs createSortKeys on every job worker with one AutoInstancer per worker; this is the one-instancer form (instancer
// index 0 in the group values), every mesh's instances in one group.  Order inside a group and among equal keys is unspecified in the
// reference too (it depends on the workers' race for result pages).
#include "lb200_internal.h"
#include "lb200_math.cuh"

#include <algorithm>
#include <new>

namespace {

using namespace lb;

constexpr int SK_THREADS = 256;
constexpr uint64_t SORT_KEY_BUCKET_SHIFT = 56;                 // pipeline.cpp:70-77
constexpr uint64_t SORT_KEY_INSTANCED_FLAG = (uint64_t)1 << 55;
constexpr uint64_t SORT_VALUE_INSTANCER_SHIFT = 16;
constexpr uint64_t SORT_VALUE_MESH_IDX_SHIFT = 40;
constexpr uint64_t SORT_VALUE_TYPE_SHIFT = 32;
enum { DRAW_MESH = 0, DRAW_AUTOINSTANCED = 1, DRAW_SKINNED = 2, DRAW_DECAL = 3, DRAW_CURVE_DECAL = 4 }; // :41-51
enum { RT_MESH = 0, RT_DECAL = 1, RT_LOCAL_LIGHT = 2, RT_CURVE_DECAL = 3 };                                // render_module.h:293-301
enum { CNT_KEYS = 0, CNT_RECS, CNT_INST, CNT_POSE, CNT_DIRTY, CNT_WORDS = 8 };

struct EmitParams {
	lb200_sk_view view;
	uint32_t type_base[4]; // offsets of the MESH / DECAL / LOCAL_LIGHT / CURVE_DECAL segments inside out_ids
	uint32_t cap_keys, cap_recs, cap_pose, cap_dirty;
};

// :57-60
__device__ __forceinline__ uint32_t float_flip(uint32_t bits) { return bits ^ ((uint32_t)(-(int32_t)(bits >> 31)) | 0x80000000u); }
__device__ __forceinline__ uint64_t sext(int32_t e) { return (uint64_t)(int64_t)e; } // EntityPtr::index is an i32: `entity.index | u64` sign-extends

// The same counter for lanes of a warp that add to the same auto-instancer group: one atomic per (warp, group).
__device__ __forceinline__ uint32_t warp_claim_keyed(uint32_t* counters, uint32_t key) {
	const uint32_t active = __activemask();
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t peers = __match_any_sync(active, key);
	const int leader = __ffs((int)peers) - 1;
	uint32_t base = 0;
	if ((int)lane == leader) base = atomicAdd(&counters[key], (uint32_t)__popc(peers));
	base = __shfl_sync(peers, base, leader);
	return base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}

// Where one renderable's outputs go.  The per-renderable logic runs twice: first with a counting sink (how many keys / instancer records /
// pose-list entries it emits), then — after ONE block-wide scan and ONE global atomic per counter and block — with a writing sink whose
// slots are already known.  A single hot counter bumped once per key serialises at the L2 (~1 M claims per view otherwise).
struct Sink {
	bool write;
	uint32_t k, r, p; // next key / record / pose slot (write) or running counts (count)
	uint32_t* s_grp;  // per block and group, in shared memory: instances counted (count), then the block's cursor inside the group (write);
	                  // null when the view has more groups than fit: ranks then come from the global counters directly
};

struct EmitArgs {
	const lb200_transform* __restrict__ transforms; const uint32_t* __restrict__ model_of; float* __restrict__ lod; const uint8_t* __restrict__ flags;
	uint32_t* __restrict__ pose_frame; const uint32_t* __restrict__ decal_sort_key; const uint8_t* __restrict__ decal_layer;
	const lb200_sk_model* __restrict__ models; const lb200_sk_mesh* __restrict__ meshes;
	uint64_t* __restrict__ keys; uint64_t* __restrict__ values; uint32_t* __restrict__ counts; uint32_t* __restrict__ group_count;
	uint8_t* __restrict__ group_layer; uint64_t* __restrict__ rec_value; uint2* __restrict__ rec_group_rank;
	uint32_t* __restrict__ pose_list; uint32_t* __restrict__ dirty_list;
	const uint32_t* s_bucket_map; // the view's bucket map in shared memory: every lane looks up its own layer
};

// the 64-byte model record / 16-byte mesh record through the read-only path, into registers
__device__ __forceinline__ lb200_sk_model load_model(const lb200_sk_model* p) {
	union { lb200_sk_model m; int4 q[4]; } u;
	const int4* s = reinterpret_cast<const int4*>(p);
	u.q[0] = __ldg(s); u.q[1] = __ldg(s + 1); u.q[2] = __ldg(s + 2); u.q[3] = __ldg(s + 3);
	return u.m;
}
__device__ __forceinline__ lb200_sk_mesh load_mesh(const lb200_sk_mesh* p) {
	union { lb200_sk_mesh m; int4 q; } u;
	u.q = __ldg(reinterpret_cast<const int4*>(p));
	return u.m;
}

__device__ __forceinline__ void push_key(const EmitParams& P, const EmitArgs& A, Sink& s, uint64_t key, uint64_t value) {
	if (s.write && s.k < P.cap_keys) { A.keys[s.k] = key; A.values[s.k] = value; }
	++s.k;
}

// DECAL / CURVE_DECAL renderable (:3840-3867)
__device__ __forceinline__ void decal_entity(const EmitParams& P, const EmitArgs& A, Sink& s, int32_t e, int type) {
	const uint8_t bucket = (uint8_t)A.s_bucket_map[A.decal_layer[e]];
	if (bucket < 0xff) {
		push_key(P, A, s, A.decal_sort_key[e] | ((uint64_t)bucket << SORT_KEY_BUCKET_SHIFT),
			sext(e) | ((uint64_t)(type == RT_DECAL ? DRAW_DECAL : DRAW_CURVE_DECAL) << SORT_VALUE_TYPE_SHIFT));
	}
}

// MESH renderable (:3868-3956)
__device__ __forceinline__ void mesh_entity(const EmitParams& P, const EmitArgs& A, Sink& s, int32_t e) {
	const float global_lod_multiplier_rcp = LB_FDIV(1.0f, P.view.lod_multiplier); // :3798-3799
	const float time_delta = P.view.time_delta;
	const bool is_shadow = P.view.is_shadow != 0;
	const lb200_sk_model model = load_model(A.models + A.model_of[e]);
	const double px = A.transforms[e].pos[0], py = A.transforms[e].pos[1], pz = A.transforms[e].pos[2];
	const double dx = LB_DSUB(px, P.view.lod_ref_point[0]), dy = LB_DSUB(py, P.view.lod_ref_point[1]), dz = LB_DSUB(pz, P.view.lod_ref_point[2]);
	const float squared_length = (float)LB_DADD(LB_DADD(LB_DMUL(dx, dx), LB_DMUL(dy, dy)), LB_DMUL(dz, dz)); // squaredLength(DVec3), math.cpp:397
	const float sd = LB_FMUL(squared_length, global_lod_multiplier_rcp);
	const uint32_t lod_idx = sd < model.lod_distances[0] ? 0u : sd < model.lod_distances[1] ? 1u : sd < model.lod_distances[2] ? 2u : sd < model.lod_distances[3] ? 3u : 4u;
	const uint8_t fl = A.flags[e];
	if (fl & LB200_SK_DIRTY) { // mi.dirty, :3878-3881: queueMaterialOverrideRefresh (rare: its own atomic)
		if (s.write) {
			const uint32_t slot = atomicAdd(&A.counts[CNT_DIRTY], 1u);
			if (slot < P.cap_dirty) A.dirty_list[slot] = (uint32_t)e;
		}
		return;
	}
	int lods[2], n_lods = 0;
	float cur = A.lod[e];
	if (cur != (float)lod_idx) { // :3926-3941
		const float d = LB_FSUB((float)lod_idx, cur);
		const float ad = fabsf(d);
		if (ad <= time_delta) {
			cur = (float)lod_idx;
			lods[n_lods++] = (int)lod_idx;
		}
		else {
			if (!is_shadow) cur = LB_FADD(cur, LB_FMUL(LB_FDIV(d, ad), time_delta));
			const uint32_t cur_lod_idx = (uint32_t)cur;
			lods[n_lods++] = (int)cur_lod_idx;
			if (cur_lod_idx < 3) lods[n_lods++] = (int)cur_lod_idx + 1;
		}
		if (s.write) A.lod[e] = cur;
	}
	else lods[n_lods++] = (int)lod_idx;
	bool pose_done = A.pose_frame[e] == P.view.frame_number;
	for (int li = 0; li < n_lods; ++li) { // create_key, :3883-3924
		const int l = lods[li]; // no dynamic indexing of the register copy
		const int from = l == 0 ? model.lod_from[0] : l == 1 ? model.lod_from[1] : l == 2 ? model.lod_from[2] : l == 3 ? model.lod_from[3] : model.lod_from[4];
		const int to = l == 0 ? model.lod_to[0] : l == 1 ? model.lod_to[1] : l == 2 ? model.lod_to[2] : l == 3 ? model.lod_to[3] : model.lod_to[4];
		for (int mesh_idx = from; mesh_idx <= to; ++mesh_idx) {
			const lb200_sk_mesh mm = load_mesh(A.meshes + model.mesh_base + (uint32_t)mesh_idx);
			const uint32_t bucket = A.s_bucket_map[mm.layer];
			if (mm.skinned) {
				// once per instance and frame: the instance's palette has to be built (PoseProcessor::push; the compare-exchange on
				// Pose::frame of :3890-3897 — one thread owns the instance within a view)
				if (!pose_done) {
					pose_done = true;
					if (s.write) {
						A.pose_frame[e] = P.view.frame_number;
						if (s.p < P.cap_pose) A.pose_list[s.p] = (uint32_t)e;
					}
					++s.p;
				}
				push_key(P, A, s, mm.sort_key | ((uint64_t)(uint8_t)bucket << SORT_KEY_BUCKET_SHIFT),
					sext(e) | ((uint64_t)DRAW_SKINNED << SORT_VALUE_TYPE_SHIFT) | ((uint64_t)mesh_idx << SORT_VALUE_MESH_IDX_SHIFT));
			}
			else if ((fl & LB200_SK_MOVED) && !is_shadow) {
				push_key(P, A, s, mm.sort_key | ((uint64_t)(uint8_t)bucket << SORT_KEY_BUCKET_SHIFT),
					sext(e) | ((uint64_t)DRAW_MESH << SORT_VALUE_TYPE_SHIFT) | ((uint64_t)mesh_idx << SORT_VALUE_MESH_IDX_SHIFT));
			}
			else if (bucket < 0xff) { // AutoInstancer::add(mesh_sort_key, e.index | mesh_idx << 40), :3913-3914
				if (!s.write) { if (s.s_grp) atomicAdd(&s.s_grp[mm.sort_key], 1u); }
				else {
					const uint32_t rank = s.s_grp ? atomicAdd(&s.s_grp[mm.sort_key], 1u) : warp_claim_keyed(A.group_count, mm.sort_key);
					if (s.r < P.cap_recs) {
						A.rec_value[s.r] = sext(e) | ((uint64_t)mesh_idx << SORT_VALUE_MESH_IDX_SHIFT);
						A.rec_group_rank[s.r] = make_uint2(mm.sort_key, rank);
					}
					if (rank == 0) A.group_layer[mm.sort_key] = mm.layer; // the same for every instance of the mesh: written once (a store per instance to a handful of bytes serialises at the L2)
				}
				++s.r;
			}
			else if (bucket < 0xffff) { // depth sorted, :3915-3922
				const double rx = LB_DSUB(px, P.view.camera_pos[0]), ry = LB_DSUB(py, P.view.camera_pos[1]), rz = LB_DSUB(pz, P.view.camera_pos[2]);
				const float sq = (float)LB_DADD(LB_DADD(LB_DMUL(rx, rx), LB_DMUL(ry, ry)), LB_DMUL(rz, rz));
				push_key(P, A, s, float_flip(__float_as_uint(sq)) | ((uint64_t)(uint8_t)bucket << SORT_KEY_BUCKET_SHIFT),
					sext(e) | ((uint64_t)DRAW_MESH << SORT_VALUE_TYPE_SHIFT) | ((uint64_t)mesh_idx << SORT_VALUE_MESH_IDX_SHIFT));
			}
		}
	}
}

// block-wide exclusive scan of one value per thread (SK_THREADS threads); returns the thread's prefix, *total = the block's sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp /* SK_THREADS / 32 + 1 */, uint32_t* total) {
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t x = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
		if (lane >= (uint32_t)d) x += y;
	}
	if (lane == 31) s_warp[warp] = x;
	__syncthreads();
	uint32_t before = 0, sum = 0;
#pragma unroll
	for (int w = 0; w < SK_THREADS / 32; ++w) { if ((uint32_t)w < warp) before += s_warp[w]; sum += s_warp[w]; }
	__syncthreads();
	*total = sum;
	return before + x - v;
}

constexpr uint32_t SK_SMEM_GROUPS = 8192; // group counters a block keeps in shared memory (32 KB)

// Every block walks its share of the visible renderables twice.  Pass 1 counts: keys / records / pose entries per thread, instances per
// auto-instancer group per block (shared-memory atomics).  Then ONE block-wide scan and one global atomic per counter claim the block's
// output ranges, and one global atomic per group the block touched claims its slice of the group.  Pass 2 repeats the logic and writes.
__global__ void __launch_bounds__(SK_THREADS) emit_kernel(const __grid_constant__ EmitParams P, const uint32_t* __restrict__ visible,
	const uint32_t* __restrict__ cull_counters, EmitArgs A, uint32_t n_groups)
{
	extern __shared__ uint32_t s_grp_mem[];
	__shared__ uint32_t s_warp[SK_THREADS / 32];
	__shared__ uint32_t s_base[3];
	__shared__ uint32_t s_bucket_map[256];
	s_bucket_map[threadIdx.x] = P.view.bucket_map[threadIdx.x]; // SK_THREADS == 256
	A.s_bucket_map = s_bucket_map;
	uint32_t* s_grp = n_groups <= SK_SMEM_GROUPS ? s_grp_mem : nullptr;
	if (s_grp) for (uint32_t g = threadIdx.x; g < n_groups; g += SK_THREADS) s_grp[g] = 0;
	__syncthreads();
	// the three segments as one index space: [MESH | DECAL | CURVE_DECAL]
	const uint32_t n_mesh = __ldg(cull_counters + RT_MESH), n_decal = __ldg(cull_counters + RT_DECAL), n_curve = __ldg(cull_counters + RT_CURVE_DECAL);
	const uint32_t n_all = n_mesh + n_decal + n_curve;
	Sink sink = {false, 0u, 0u, 0u, s_grp};
#pragma unroll 1
	for (int pass = 0; pass < 2; ++pass) {
		for (uint32_t i = blockIdx.x * SK_THREADS + threadIdx.x; i < n_all; i += gridDim.x * SK_THREADS) {
			if (i < n_mesh) mesh_entity(P, A, sink, (int32_t)visible[P.type_base[RT_MESH] + i]);
			else if (i < n_mesh + n_decal) decal_entity(P, A, sink, (int32_t)visible[P.type_base[RT_DECAL] + (i - n_mesh)], RT_DECAL);
			else decal_entity(P, A, sink, (int32_t)visible[P.type_base[RT_CURVE_DECAL] + (i - n_mesh - n_decal)], RT_CURVE_DECAL);
		}
		if (pass == 1) break;
		uint32_t tk, tr, tp;
		const uint32_t pk = block_exclusive_scan(sink.k, s_warp, &tk);
		const uint32_t pr = block_exclusive_scan(sink.r, s_warp, &tr);
		const uint32_t pp = block_exclusive_scan(sink.p, s_warp, &tp);
		if (threadIdx.x == 0) {
			s_base[0] = tk ? atomicAdd(&A.counts[CNT_KEYS], tk) : 0u;
			s_base[1] = tr ? atomicAdd(&A.counts[CNT_RECS], tr) : 0u;
			s_base[2] = tp ? atomicAdd(&A.counts[CNT_POSE], tp) : 0u;
		}
		// the block's slice of every group it has instances of: count -> cursor
		if (s_grp) for (uint32_t g = threadIdx.x; g < n_groups; g += SK_THREADS) if (s_grp[g]) s_grp[g] = atomicAdd(&A.group_count[g], s_grp[g]);
		__syncthreads();
		sink.write = true;
		sink.k = s_base[0] + pk; sink.r = s_base[1] + pr; sink.p = s_base[2] + pp;
	}
}

// one block: exclusive scan of the group counters, then one key/value per non-empty group (:3958-3969)
struct GroupParams { uint8_t layer_to_bucket[256]; uint32_t n_groups, cap_keys; };

__global__ void __launch_bounds__(1024) groups_kernel(const __grid_constant__ GroupParams P, const uint32_t* __restrict__ group_count,
	uint32_t* __restrict__ group_offset, const uint8_t* __restrict__ group_layer, uint64_t* __restrict__ keys, uint64_t* __restrict__ values,
	uint32_t* __restrict__ counts)
{
	__shared__ uint32_t s_warp[32];
	__shared__ uint32_t s_carry;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	if (tid == 0) s_carry = 0;
	__syncthreads();
	for (uint32_t base = 0; base < P.n_groups; base += 1024) {
		const uint32_t g = base + tid;
		const uint32_t c = g < P.n_groups ? group_count[g] : 0u;
		uint32_t x = c;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
			if (lane >= (uint32_t)d) x += y;
		}
		if (lane == 31) s_warp[warp] = x;
		__syncthreads();
		uint32_t before = s_carry;
		for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
		if (g < P.n_groups) {
			group_offset[g] = before + x - c;
			if (c) {
				const uint32_t slot = atomicAdd(&counts[CNT_KEYS], 1u);
				if (slot < P.cap_keys) {
					keys[slot] = (uint64_t)g | SORT_KEY_INSTANCED_FLAG | ((uint64_t)P.layer_to_bucket[group_layer[g]] << SORT_KEY_BUCKET_SHIFT); // :100-102
					values[slot] = (uint64_t)g | ((uint64_t)0 << SORT_VALUE_INSTANCER_SHIFT) | ((uint64_t)DRAW_AUTOINSTANCED << SORT_VALUE_TYPE_SHIFT); // :141-143
				}
			}
		}
		__syncthreads();
		if (tid == 1023) s_carry = before + x;
		__syncthreads();
	}
	if (tid == 0) counts[CNT_INST] = s_carry;
}

// instance data of the auto-instanced meshes, :3990-4008
__global__ void __launch_bounds__(SK_THREADS) fill_kernel(const double cx, const double cy, const double cz, uint32_t cap_recs, const uint32_t* __restrict__ counts,
	const uint64_t* __restrict__ rec_value, const uint2* __restrict__ rec_group_rank, const uint32_t* __restrict__ group_offset,
	const lb200_transform* __restrict__ transforms, const uint32_t* __restrict__ model_of, const float* __restrict__ lod,
	const lb200_sk_model* __restrict__ models, const lb200_sk_mesh* __restrict__ meshes, uint64_t* __restrict__ group_renderables, float4* __restrict__ instance_data)
{
	const uint32_t n = min(counts[CNT_RECS], cap_recs);
	for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
		const uint64_t v = rec_value[r];
		const uint2 gr = rec_group_rank[r];
		const uint32_t at = group_offset[gr.x] + gr.y;
		const int32_t e = (int32_t)(uint32_t)v;
		const uint32_t mesh_idx = (uint32_t)(v >> SORT_VALUE_MESH_IDX_SHIFT);
		const lb200_sk_mesh mm = meshes[__ldg(&models[model_of[e]].mesh_base) + mesh_idx];
		const lb200_transform& tr = transforms[e];
		const float lx = (float)LB_DSUB(tr.pos[0], cx), ly = (float)LB_DSUB(tr.pos[1], cy), lz = (float)LB_DSUB(tr.pos[2], cz); // Vec3(tr.pos - camera_pos)
		const float lod_d = LB_FSUB(lod[e], mm.lod);
		group_renderables[at] = v;
		float4* dst = instance_data + (size_t)at * 3;
		dst[0] = make_float4(tr.rot[0], tr.rot[1], tr.rot[2], tr.rot[3]);
		dst[1] = make_float4(lx, ly, lz, lod_d);
		dst[2] = make_float4(tr.scale[0], tr.scale[1], tr.scale[2], __uint_as_float(mm.material_index));
	}
}

// ---------------------------------------------------------------- radix sort ----------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_PASSES = 8;
struct SortState { uint32_t cur; uint32_t done_blocks; uint32_t pad[2]; uint32_t global_hist[RS_PASSES][256]; };

__global__ void __launch_bounds__(RS_THREADS) rs_global_hist_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint32_t cap, SortState* st) {
	__shared__ uint32_t s_h[RS_PASSES][256];
	for (int i = threadIdx.x; i < RS_PASSES * 256; i += RS_THREADS) (&s_h[0][0])[i] = 0;
	__syncthreads();
	const uint32_t n = min(counts[0], cap);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint64_t k = keys[i];
#pragma unroll
		for (int p = 0; p < RS_PASSES; ++p) atomicAdd(&s_h[p][(k >> (8 * p)) & 0xffu], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < RS_PASSES * 256; i += RS_THREADS) if ((&s_h[0][0])[i]) atomicAdd(&(&st->global_hist[0][0])[i], (&s_h[0][0])[i]);
}

// a pass moves nothing when every key has the same digit there (the reference's skip at pipeline.cpp:4120 is the bin-0 case of this)
__device__ __forceinline__ bool pass_is_trivial(const SortState* st, int pass, uint32_t n) {
	return st->global_hist[pass][0] == n;
}

// keys of block b: [b * per, min(n, (b + 1) * per)), per = ceil(n / blocks) rounded up to RS_THREADS
__device__ __forceinline__ void block_range(uint32_t n, uint32_t& begin, uint32_t& end) {
	uint32_t per = (n + gridDim.x - 1) / gridDim.x;
	per = (per + RS_THREADS - 1) / RS_THREADS * RS_THREADS;
	begin = min(n, blockIdx.x * per);
	end = min(n, begin + per);
}

__global__ void __launch_bounds__(RS_THREADS) rs_block_hist_kernel(int pass, const uint64_t* __restrict__ buf0, const uint64_t* __restrict__ buf1,
	const uint32_t* __restrict__ counts, uint32_t cap, const SortState* __restrict__ st, uint32_t* __restrict__ block_hist /* [256][gridDim] */)
{
	__shared__ uint32_t s_h[256];
	const uint32_t n = min(counts[0], cap);
	if (pass_is_trivial(st, pass, n)) return;
	const uint64_t* keys = st->cur ? buf1 : buf0;
	s_h[threadIdx.x] = 0;
	__syncthreads();
	uint32_t begin, end;
	block_range(n, begin, end);
	for (uint32_t i = begin + threadIdx.x; i < end; i += RS_THREADS) atomicAdd(&s_h[(keys[i] >> (8 * pass)) & 0xffu], 1u);
	__syncthreads();
	block_hist[threadIdx.x * gridDim.x + blockIdx.x] = s_h[threadIdx.x];
}

// exclusive scan over block_hist in (digit, block) order.  Block d of the grid owns digit d: its base is the number of keys with a
// smaller digit (the pass's global histogram), then one warp scans the digit's per-block counts 32 at a time.
__global__ void __launch_bounds__(32) rs_scan_kernel(int pass, const uint32_t* __restrict__ counts, uint32_t cap, const SortState* __restrict__ st, uint32_t* __restrict__ block_hist, uint32_t n_blocks) {
	const uint32_t n = min(counts[0], cap);
	if (pass_is_trivial(st, pass, n)) return;
	const uint32_t d = blockIdx.x, lane = threadIdx.x;
	uint32_t base = 0;
	for (uint32_t k = lane; k < d; k += 32) base += st->global_hist[pass][k];
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
	uint32_t* row = block_hist + (size_t)d * n_blocks;
	for (uint32_t b0 = 0; b0 < n_blocks; b0 += 32) {
		const uint32_t b = b0 + lane;
		const uint32_t c = b < n_blocks ? row[b] : 0u;
		uint32_t x = c;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
			if (lane >= (uint32_t)o) x += y;
		}
		if (b < n_blocks) row[b] = base + x - c;
		base += __shfl_sync(0xffffffffu, x, 31);
	}
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(int pass, uint64_t* __restrict__ kbuf0, uint64_t* __restrict__ kbuf1, uint64_t* __restrict__ vbuf0,
	uint64_t* __restrict__ vbuf1, const uint32_t* __restrict__ counts, uint32_t cap, SortState* st, const uint32_t* __restrict__ block_hist)
{
	__shared__ uint32_t s_digit_base[256];        // where this block's next key of digit d goes
	__shared__ uint16_t s_warp_cnt[RS_WARPS][256]; // keys of digit d in warp w of the current tile
	const uint32_t n = min(counts[0], cap);
	if (pass_is_trivial(st, pass, n)) return;
	const uint32_t cur = st->cur;
	const uint64_t* ksrc = cur ? kbuf1 : kbuf0;
	const uint64_t* vsrc = cur ? vbuf1 : vbuf0;
	uint64_t* kdst = cur ? kbuf0 : kbuf1;
	uint64_t* vdst = cur ? vbuf0 : vbuf1;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	s_digit_base[tid] = block_hist[tid * gridDim.x + blockIdx.x];
	for (int w = 0; w < RS_WARPS; ++w) s_warp_cnt[w][tid] = 0;
	__syncthreads();
	uint32_t begin, end;
	block_range(n, begin, end);
	for (uint32_t tile = begin; tile < end; tile += RS_THREADS) {
		const uint32_t i = tile + tid;
		const bool has = i < end;
		const uint64_t k = has ? ksrc[i] : 0;
		const uint64_t v = has ? vsrc[i] : 0;
		const uint32_t d = has ? (uint32_t)((k >> (8 * pass)) & 0xffu) : 0x100u; // inactive threads match only each other
		const uint32_t peers = __match_any_sync(0xffffffffu, d);
		const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
		if (has && rank_in_warp == 0) s_warp_cnt[warp][d] = (uint16_t)__popc(peers);
		__syncthreads();
		if (has) {
			uint32_t before = s_digit_base[d];
			for (uint32_t w = 0; w < warp; ++w) before += s_warp_cnt[w][d];
			const uint32_t dest = before + rank_in_warp;
			kdst[dest] = k;
			vdst[dest] = v;
		}
		__syncthreads();
		{ // thread d: advance the block's cursor of digit d past this tile, clear the per-warp counts
			uint32_t tot = 0;
#pragma unroll
			for (int w = 0; w < RS_WARPS; ++w) { tot += s_warp_cnt[w][tid]; s_warp_cnt[w][tid] = 0; }
			s_digit_base[tid] += tot;
		}
		__syncthreads();
	}
	// the last block to finish flips the buffers
	__shared__ bool s_last;
	__threadfence();
	if (tid == 0) s_last = atomicAdd(&st->done_blocks, 1u) == gridDim.x - 1;
	__syncthreads();
	if (s_last && tid == 0) { st->cur = cur ^ 1u; st->done_blocks = 0; }
}

// sorted data to buffer 0 if it ended up in buffer 1
__global__ void __launch_bounds__(RS_THREADS) rs_finish_kernel(uint64_t* __restrict__ kbuf0, const uint64_t* __restrict__ kbuf1, uint64_t* __restrict__ vbuf0,
	const uint64_t* __restrict__ vbuf1, const uint32_t* __restrict__ counts, uint32_t cap, const SortState* __restrict__ st)
{
	if (!st->cur) return;
	const uint32_t n = min(counts[0], cap);
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { kbuf0[i] = kbuf1[i]; vbuf0[i] = vbuf1[i]; }
}

} // namespace

// Stable LSD radix sort of n = min(*count_dev, cap) (key, value) pairs of 64 bits, all launches on `stream`, n read on the device.  The
// sorted pairs end in buffer 0.  state: lb200_radix_sort_state_bytes() bytes, block_hist: 256 x blocks words.  (Also used by the device
// re-binning of the culling structure, culling.cu.)
size_t lb200_radix_sort_state_bytes() { return sizeof(SortState); }

int lb200_radix_sort_pairs(lb200_ctx* ctx, cudaStream_t s, uint64_t* keys0, uint64_t* keys1, uint64_t* values0, uint64_t* values1, const uint32_t* count_dev, uint32_t cap,
	void* state, uint32_t* block_hist, uint32_t blocks)
{
	SortState* st = (SortState*)state;
	LB200_CUDA(ctx, cudaMemsetAsync(st, 0, sizeof(SortState), s));
	rs_global_hist_kernel<<<blocks, RS_THREADS, 0, s>>>(keys0, count_dev, cap, st);
	LB200_CHECK_LAUNCH(ctx);
	for (int pass = 0; pass < RS_PASSES; ++pass) {
		rs_block_hist_kernel<<<blocks, RS_THREADS, 0, s>>>(pass, keys0, keys1, count_dev, cap, st, block_hist);
		LB200_CHECK_LAUNCH(ctx);
		rs_scan_kernel<<<256, 32, 0, s>>>(pass, count_dev, cap, st, block_hist, blocks);
		LB200_CHECK_LAUNCH(ctx);
		rs_scatter_kernel<<<blocks, RS_THREADS, 0, s>>>(pass, keys0, keys1, values0, values1, count_dev, cap, st, block_hist);
		LB200_CHECK_LAUNCH(ctx);
	}
	rs_finish_kernel<<<blocks, RS_THREADS, 0, s>>>(keys0, keys1, values0, values1, count_dev, cap, st);
	LB200_CHECK_LAUNCH(ctx);
	return LB200_OK;
}

struct lb200_sortkeys {
	lb200_ctx* ctx = nullptr;
	uint32_t max_entities = 0, max_groups = 0;
	uint32_t cap_keys = 0, cap_recs = 0;
	// inputs
	lb200_transform* d_transforms = nullptr; // owned copy (set_transforms) ...
	const lb200_transform* transforms = nullptr; // ... or the caller's device array
	uint32_t* d_model_of = nullptr; float* d_lod = nullptr; uint8_t* d_flags = nullptr; uint32_t* d_pose_frame = nullptr;
	uint32_t* d_decal_sort_key = nullptr; uint8_t* d_decal_layer = nullptr;
	lb200_sk_model* d_models = nullptr; lb200_sk_mesh* d_meshes = nullptr; uint32_t n_models = 0, n_meshes = 0;
	// outputs
	uint64_t *d_keys[2] = {}, *d_values[2] = {};
	uint32_t* d_counts = nullptr; uint32_t* h_counts = nullptr; // pinned
	uint32_t *d_group_count = nullptr, *d_group_offset = nullptr; uint8_t* d_group_layer = nullptr;
	uint64_t* d_rec_value = nullptr; uint2* d_rec_group_rank = nullptr;
	uint64_t* d_group_renderables = nullptr; float4* d_instance_data = nullptr;
	uint32_t *d_pose_list = nullptr, *d_dirty_list = nullptr;
	SortState* d_sort_state = nullptr; uint32_t* d_block_hist = nullptr;
	uint32_t sort_blocks = 0;
	uint32_t last_groups = 0;
};

extern "C" {

int lb200_sortkeys_create(lb200_ctx* ctx, uint32_t max_entities, uint32_t max_groups, uint32_t max_keys, uint32_t max_instances, lb200_sortkeys** out) {
	if (!ctx || !out || !max_entities || !max_groups) return LB200_ERR_INVALID;
	LB200_CUDA(ctx, cudaSetDevice(ctx->device));
	lb200_sortkeys* sk = new (std::nothrow) lb200_sortkeys;
	if (!sk) return LB200_ERR_CUDA;
	sk->ctx = ctx; sk->max_entities = max_entities; sk->max_groups = max_groups;
	sk->cap_keys = max_keys ? max_keys : max_entities; sk->cap_recs = max_instances ? max_instances : max_entities;
	const size_t E = max_entities;
	LB200_CUDA(ctx, cudaMalloc(&sk->d_model_of, sizeof(uint32_t) * E));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_lod, sizeof(float) * E));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_flags, E));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_pose_frame, sizeof(uint32_t) * E));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_decal_sort_key, sizeof(uint32_t) * E));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_decal_layer, E));
	LB200_CUDA(ctx, cudaMemsetAsync(sk->d_model_of, 0, sizeof(uint32_t) * E, ctx->stream));
	LB200_CUDA(ctx, cudaMemsetAsync(sk->d_flags, 0, E, ctx->stream));
	LB200_CUDA(ctx, cudaMemsetAsync(sk->d_pose_frame, 0xff, sizeof(uint32_t) * E, ctx->stream)); // 0xffffffff marks "never" (pipeline.cpp:3814)
	LB200_CUDA(ctx, cudaMemsetAsync(sk->d_decal_sort_key, 0, sizeof(uint32_t) * E, ctx->stream));
	LB200_CUDA(ctx, cudaMemsetAsync(sk->d_decal_layer, 0, E, ctx->stream));
	for (int b = 0; b < 2; ++b) {
		LB200_CUDA(ctx, cudaMalloc(&sk->d_keys[b], sizeof(uint64_t) * sk->cap_keys));
		LB200_CUDA(ctx, cudaMalloc(&sk->d_values[b], sizeof(uint64_t) * sk->cap_keys));
	}
	LB200_CUDA(ctx, cudaMalloc(&sk->d_counts, sizeof(uint32_t) * CNT_WORDS));
	LB200_CUDA(ctx, cudaHostAlloc(&sk->h_counts, sizeof(uint32_t) * CNT_WORDS, cudaHostAllocDefault));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_group_count, sizeof(uint32_t) * max_groups));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_group_offset, sizeof(uint32_t) * max_groups));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_group_layer, max_groups));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_rec_value, sizeof(uint64_t) * sk->cap_recs));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_rec_group_rank, sizeof(uint2) * sk->cap_recs));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_group_renderables, sizeof(uint64_t) * sk->cap_recs));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_instance_data, 48 * (size_t)sk->cap_recs));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_pose_list, sizeof(uint32_t) * E));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_dirty_list, sizeof(uint32_t) * E));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_sort_state, sizeof(SortState)));
	sk->sort_blocks = (uint32_t)ctx->sm_count * 4;
	LB200_CUDA(ctx, cudaMalloc(&sk->d_block_hist, sizeof(uint32_t) * 256 * sk->sort_blocks));
	*out = sk;
	return LB200_OK;
}

void lb200_sortkeys_destroy(lb200_sortkeys* sk) {
	if (!sk) return;
	cudaSetDevice(sk->ctx->device);
	cudaStreamSynchronize(sk->ctx->stream);
	cudaFree(sk->d_transforms); cudaFree(sk->d_model_of); cudaFree(sk->d_lod); cudaFree(sk->d_flags); cudaFree(sk->d_pose_frame);
	cudaFree(sk->d_decal_sort_key); cudaFree(sk->d_decal_layer); cudaFree(sk->d_models); cudaFree(sk->d_meshes);
	for (int b = 0; b < 2; ++b) { cudaFree(sk->d_keys[b]); cudaFree(sk->d_values[b]); }
	cudaFree(sk->d_counts); if (sk->h_counts) cudaFreeHost(sk->h_counts);
	cudaFree(sk->d_group_count); cudaFree(sk->d_group_offset); cudaFree(sk->d_group_layer); cudaFree(sk->d_rec_value); cudaFree(sk->d_rec_group_rank);
	cudaFree(sk->d_group_renderables); cudaFree(sk->d_instance_data); cudaFree(sk->d_pose_list); cudaFree(sk->d_dirty_list);
	cudaFree(sk->d_sort_state); cudaFree(sk->d_block_hist);
	delete sk;
}

int lb200_sortkeys_set_models(lb200_sortkeys* sk, const lb200_sk_model* models, uint32_t n_models, const lb200_sk_mesh* meshes, uint32_t n_meshes) {
	if (!sk || !models || !meshes || !n_models || !n_meshes) return LB200_ERR_INVALID;
	lb200_ctx* ctx = sk->ctx;
	for (uint32_t i = 0; i < n_meshes; ++i) if (meshes[i].sort_key >= sk->max_groups) { lb200_set_error(ctx, "mesh %u: sort key %u >= max_groups %u", i, meshes[i].sort_key, sk->max_groups); return LB200_ERR_INVALID; }
	LB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFree(sk->d_models); cudaFree(sk->d_meshes);
	sk->d_models = nullptr; sk->d_meshes = nullptr;
	LB200_CUDA(ctx, cudaMalloc(&sk->d_models, sizeof(lb200_sk_model) * n_models));
	LB200_CUDA(ctx, cudaMalloc(&sk->d_meshes, sizeof(lb200_sk_mesh) * n_meshes));
	LB200_CUDA(ctx, cudaMemcpyAsync(sk->d_models, models, sizeof(lb200_sk_model) * n_models, cudaMemcpyHostToDevice, ctx->stream));
	LB200_CUDA(ctx, cudaMemcpyAsync(sk->d_meshes, meshes, sizeof(lb200_sk_mesh) * n_meshes, cudaMemcpyHostToDevice, ctx->stream));
	LB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	sk->n_models = n_models; sk->n_meshes = n_meshes;
	return LB200_OK;
}

// per-entity state, arrays indexed by entity id (n <= max_entities); null pointers leave that array as it is
int lb200_sortkeys_set_instances(lb200_sortkeys* sk, uint32_t n, const uint32_t* model_of, const float* lod, const uint8_t* flags, const uint32_t* pose_frame,
	const uint32_t* decal_sort_key, const uint8_t* decal_layer)
{
	if (!sk || n > sk->max_entities) return LB200_ERR_INVALID;
	lb200_ctx* ctx = sk->ctx;
	if (model_of) LB200_CUDA(ctx, cudaMemcpyAsync(sk->d_model_of, model_of, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
	if (lod) LB200_CUDA(ctx, cudaMemcpyAsync(sk->d_lod, lod, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
	if (flags) LB200_CUDA(ctx, cudaMemcpyAsync(sk->d_flags, flags, n, cudaMemcpyHostToDevice, ctx->stream));
	if (pose_frame) LB200_CUDA(ctx, cudaMemcpyAsync(sk->d_pose_frame, pose_frame, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
	if (decal_sort_key) LB200_CUDA(ctx, cudaMemcpyAsync(sk->d_decal_sort_key, decal_sort_key, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
	if (decal_layer) LB200_CUDA(ctx, cudaMemcpyAsync(sk->d_decal_layer, decal_layer, n, cudaMemcpyHostToDevice, ctx->stream));
	LB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return LB200_OK;
}

int lb200_sortkeys_set_transforms(lb200_sortkeys* sk, const lb200_transform* transforms, uint32_t n) {
	if (!sk || !transforms || n > sk->max_entities) return LB200_ERR_INVALID;
	lb200_ctx* ctx = sk->ctx;
	if (!sk->d_transforms) LB200_CUDA(ctx, cudaMalloc(&sk->d_transforms, sizeof(lb200_transform) * (size_t)sk->max_entities));
	LB200_CUDA(ctx, cudaMemcpyAsync(sk->d_transforms, transforms, sizeof(lb200_transform) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
	LB200_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	sk->transforms = sk->d_transforms;
	return LB200_OK;
}

int lb200_sortkeys_set_transforms_device(lb200_sortkeys* sk, const lb200_transform* dev_transforms) {
	if (!sk || !dev_transforms) return LB200_ERR_INVALID;
	sk->transforms = dev_transforms;
	return LB200_OK;
}

int lb200_sortkeys_create_keys(lb200_sortkeys* sk, lb200_culling* cs, const lb200_sk_view* view, int sort, int want_counts, lb200_sk_result* result) {
	if (!sk || !cs || !view) return LB200_ERR_INVALID;
	lb200_ctx* ctx = sk->ctx;
	lb200_range range("create keys"); // pipeline.cpp:3818
	if (!sk->transforms || !sk->d_models) { lb200_set_error(ctx, "create_keys needs set_models and set_transforms first"); return LB200_ERR_STATE; }
	if (view->max_sort_key >= sk->max_groups) { lb200_set_error(ctx, "view.max_sort_key %u >= max_groups %u", view->max_sort_key, sk->max_groups); return LB200_ERR_INVALID; }
	const uint32_t *visible = nullptr, *cull_counters = nullptr, *type_base = nullptr, *type_counts = nullptr;
	int rc = lb200_culling_internal_last(cs, &visible, &cull_counters, &type_base, &type_counts);
	if (rc) return rc;
	cudaStream_t s = ctx->stream;
	const uint32_t n_groups = view->max_sort_key + 1;
	sk->last_groups = n_groups;
	LB200_CUDA(ctx, cudaMemsetAsync(sk->d_counts, 0, sizeof(uint32_t) * CNT_WORDS, s));
	LB200_CUDA(ctx, cudaMemsetAsync(sk->d_group_count, 0, sizeof(uint32_t) * n_groups, s));
	EmitParams EP;
	EP.view = *view;
	for (int t = 0; t < 4; ++t) EP.type_base[t] = type_base[t];
	EP.cap_keys = sk->cap_keys; EP.cap_recs = sk->cap_recs; EP.cap_pose = sk->max_entities; EP.cap_dirty = sk->max_entities;
	const uint32_t work = type_counts[RT_MESH] + type_counts[RT_DECAL] + type_counts[RT_CURVE_DECAL]; // upper bound of visible renderables
	const uint32_t grid = std::max(1u, std::min((uint32_t)ctx->sm_count * 4u, (work + SK_THREADS - 1) / SK_THREADS));
	EmitArgs EA = {sk->transforms, sk->d_model_of, sk->d_lod, sk->d_flags, sk->d_pose_frame, sk->d_decal_sort_key, sk->d_decal_layer, sk->d_models, sk->d_meshes,
		sk->d_keys[0], sk->d_values[0], sk->d_counts, sk->d_group_count, sk->d_group_layer, sk->d_rec_value, sk->d_rec_group_rank, sk->d_pose_list, sk->d_dirty_list, nullptr};
	emit_kernel<<<grid, SK_THREADS, n_groups <= SK_SMEM_GROUPS ? sizeof(uint32_t) * n_groups : 0, s>>>(EP, visible, cull_counters, EA, n_groups);
	LB200_CHECK_LAUNCH(ctx);
	GroupParams GP;
	memcpy(GP.layer_to_bucket, view->layer_to_bucket, 256);
	GP.n_groups = n_groups; GP.cap_keys = sk->cap_keys;
	groups_kernel<<<1, 1024, 0, s>>>(GP, sk->d_group_count, sk->d_group_offset, sk->d_group_layer, sk->d_keys[0], sk->d_values[0], sk->d_counts);
	LB200_CHECK_LAUNCH(ctx);
	fill_kernel<<<grid, SK_THREADS, 0, s>>>(view->camera_pos[0], view->camera_pos[1], view->camera_pos[2], sk->cap_recs, sk->d_counts, sk->d_rec_value,
		sk->d_rec_group_rank, sk->d_group_offset, sk->transforms, sk->d_model_of, sk->d_lod, sk->d_models, sk->d_meshes, sk->d_group_renderables, sk->d_instance_data);
	LB200_CHECK_LAUNCH(ctx);
	if (sort) {
		lb200_range r2("radixSort"); // pipeline.cpp:4101
		rc = lb200_radix_sort_pairs(ctx, s, sk->d_keys[0], sk->d_keys[1], sk->d_values[0], sk->d_values[1], sk->d_counts + CNT_KEYS, sk->cap_keys, sk->d_sort_state, sk->d_block_hist, sk->sort_blocks);
		if (rc) return rc;
	}
	if (want_counts) {
		if (!result) return LB200_ERR_INVALID;
		LB200_CUDA(ctx, cudaMemcpyAsync(sk->h_counts, sk->d_counts, sizeof(uint32_t) * CNT_WORDS, cudaMemcpyDeviceToHost, s));
		LB200_CUDA(ctx, cudaStreamSynchronize(s));
		result->n_keys = sk->h_counts[CNT_KEYS]; result->n_instances = sk->h_counts[CNT_INST]; result->n_pose = sk->h_counts[CNT_POSE];
		result->n_dirty = sk->h_counts[CNT_DIRTY]; result->n_groups = n_groups;
		if (result->n_keys > sk->cap_keys || sk->h_counts[CNT_RECS] > sk->cap_recs) { lb200_set_error(ctx, "create_keys: %u keys / %u instances exceed the capacities %u / %u", result->n_keys, sk->h_counts[CNT_RECS], sk->cap_keys, sk->cap_recs); return LB200_ERR_CAPACITY; }
	}
	return LB200_OK;
}

// device pointers of the last create_keys (valid until the next one): sorted keys / values, group tables, instance data, lists
int lb200_sortkeys_device_outputs(lb200_sortkeys* sk, lb200_sk_outputs* out) {
	if (!sk || !out) return LB200_ERR_INVALID;
	out->keys = sk->d_keys[0]; out->values = sk->d_values[0]; out->group_count = sk->d_group_count; out->group_offset = sk->d_group_offset;
	out->group_renderables = sk->d_group_renderables; out->instance_data = sk->d_instance_data; out->pose_list = sk->d_pose_list; out->dirty_list = sk->d_dirty_list;
	out->lod = sk->d_lod; out->pose_frame = sk->d_pose_frame;
	return LB200_OK;
}

} // extern "C"
