// TEST INFRASTRUCTURE.  The arithmetic header of the CUDA kernels (lumixengine_b200/csrc/lb200_math.cuh) also compiles for the host;
// this file exports its functions with a C ABI so that tests/test_host_math.py can check, without a GPU, that every expression follows
// the reference's operation order: the outputs must equal the reference-run vectors of tests/golden/math_kat.npz bit for bit.
#include "lb200_math.cuh"

using namespace lb;

extern "C" {

void hm_quat_mul(const float* a, const float* b, float* out, unsigned n) {
	for (unsigned i = 0; i < n; ++i) {
		const Q4 r = qmul(q4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]), q4(b[4 * i], b[4 * i + 1], b[4 * i + 2], b[4 * i + 3]));
		out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
	}
}

void hm_quat_rotate(const float* q, const float* v, float* out, unsigned n) {
	for (unsigned i = 0; i < n; ++i) {
		const V3 r = rotate(q4(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]), v3(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
		out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
	}
}

void hm_simd_nlerp(const float* a, const float* b, const float* t, float* out, unsigned n) {
	for (unsigned i = 0; i < n; ++i) {
		const Q4 r = simd_nlerp(q4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]), q4(b[4 * i], b[4 * i + 1], b[4 * i + 2], b[4 * i + 3]), t[i]);
		out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
	}
}

void hm_lerp(const float* a, const float* b, const float* t, float* out, unsigned n) {
	for (unsigned i = 0; i < n; ++i) {
		const V3 r = lerp(v3(a[3 * i], a[3 * i + 1], a[3 * i + 2]), v3(b[3 * i], b[3 * i + 1], b[3 * i + 2]), t[i]);
		out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
	}
}

static Rigid rigid7(const float* p) { Rigid r; r.pos = v3(p[0], p[1], p[2]); r.rot = q4(p[3], p[4], p[5], p[6]); return r; }

void hm_lrt_mul(const float* a, const float* b, float* out7, unsigned n) {
	for (unsigned i = 0; i < n; ++i) {
		const Rigid r = rmul(rigid7(a + 7 * i), rigid7(b + 7 * i));
		float* o = out7 + 7 * i;
		o[0] = r.pos.x; o[1] = r.pos.y; o[2] = r.pos.z; o[3] = r.rot.x; o[4] = r.rot.y; o[5] = r.rot.z; o[6] = r.rot.w;
	}
}

void hm_lrt_to_dual_quat(const float* a, float* out8, unsigned n) {
	for (unsigned i = 0; i < n; ++i) {
		const DualQ d = to_dual_quat(rigid7(a + 7 * i));
		float* o = out8 + 8 * i;
		o[0] = d.r.x; o[1] = d.r.y; o[2] = d.r.z; o[3] = d.r.w; o[4] = d.d.x; o[5] = d.d.y; o[6] = d.d.z; o[7] = d.d.w;
	}
}

void hm_lrt_to_matrix(const float* a, float* out16, unsigned n) {
	for (unsigned i = 0; i < n; ++i) to_matrix(rigid7(a + 7 * i), out16 + 16 * i);
}

// Transform::compose (math.cpp:801-807) exactly as compose_node / propagate_level_kernel spell it out (csrc/hierarchy.cu)
struct Tr { double pos[3]; float rot[4]; float scale[3]; float pad; };
void hm_transform_compose(const Tr* parent, const Tr* local, Tr* out, unsigned n) {
	for (unsigned i = 0; i < n; ++i) {
		const Tr& p = parent[i];
		const Tr& l = local[i];
		const Q4 prot = q4(p.rot[0], p.rot[1], p.rot[2], p.rot[3]);
		const D3 scaled = d3(LB_DMUL(l.pos[0], (double)p.scale[0]), LB_DMUL(l.pos[1], (double)p.scale[1]), LB_DMUL(l.pos[2], (double)p.scale[2]));
		const D3 gpos = add(rotate(prot, scaled), d3(p.pos[0], p.pos[1], p.pos[2]));
		const Q4 grot = qmul(prot, q4(l.rot[0], l.rot[1], l.rot[2], l.rot[3]));
		const V3 gs = mul(v3(p.scale[0], p.scale[1], p.scale[2]), v3(l.scale[0], l.scale[1], l.scale[2]));
		Tr& o = out[i];
		o.pos[0] = gpos.x; o.pos[1] = gpos.y; o.pos[2] = gpos.z;
		o.rot[0] = grot.x; o.rot[1] = grot.y; o.rot[2] = grot.z; o.rot[3] = grot.w;
		o.scale[0] = gs.x; o.scale[1] = gs.y; o.scale[2] = gs.z; o.pad = 0;
	}
}

// the body of bone_attachments_kernel (csrc/animation.cu), expression for expression: updateBoneAttachment, render_module.cpp:399-403
void hm_bone_attachments(const Tr* parent, const float* bone7, const float* relative7, const float* scale3, Tr* out, unsigned n) {
	for (unsigned i = 0; i < n; ++i) {
		const Rigid local = rmul(rigid7(bone7 + 7 * i), rigid7(relative7 + 7 * i));
		const Tr& p = parent[i];
		const Q4 prot = q4(p.rot[0], p.rot[1], p.rot[2], p.rot[3]);
		const V3 rotated = rotate(prot, mul(local.pos, v3(p.scale[0], p.scale[1], p.scale[2])));
		const Q4 rot = qmul(prot, local.rot);
		Tr& o = out[i];
		o.pos[0] = LB_DADD(p.pos[0], (double)rotated.x); o.pos[1] = LB_DADD(p.pos[1], (double)rotated.y); o.pos[2] = LB_DADD(p.pos[2], (double)rotated.z);
		o.rot[0] = rot.x; o.rot[1] = rot.y; o.rot[2] = rot.z; o.rot[3] = rot.w;
		o.scale[0] = scale3[3 * i]; o.scale[1] = scale3[3 * i + 1]; o.scale[2] = scale3[3 * i + 2]; o.pad = 0;
	}
}

} // extern "C"
