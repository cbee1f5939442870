/* TEST INFRASTRUCTURE — CPU restatement of the reference's arithmetic for the hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 * The shipped product (lumixengine_b200/csrc) never includes, links or calls anything in oracle/.
 *
 * Every function cites the reference line it restates (paths relative to /root/reference).
 * Pinned against the reference's own compiled math.cpp/geometry.cpp (oracle/_ref) by
 * tests/test_oracle_ref.py — the reference has no golden vectors of its own for this path
 * (SURVEY.md F9), so "parity pinned by reference-run outputs", see tests/golden/.
 *
 * Build: gcc -O2 -msse2 -ffp-contract=off (no FMA, no fast-math: scripts/genie.lua:339-342).
 */
#ifndef ORACLE_MATH_H
#define ORACLE_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float x, y, z; } OVec3;
typedef struct { double x, y, z; } ODVec3;
typedef struct { int x, y, z; } OIVec3;
typedef struct { float x, y, z, w; } OQuat;
typedef struct { OVec3 pos; OQuat rot; } OLocalRigidTransform; /* math.h:266-274, 28 B */
typedef struct { OQuat r, d; } ODualQuat;                      /* math.h:261-264, 32 B */
typedef struct { float m[16]; } OMatrix;                       /* math.h:329-392, column-major Vec4 columns[4] */
typedef struct { ODVec3 pos; OQuat rot; OVec3 scale; } OTransform; /* math.h:306-327, 56 B */

/* geometry.h:17-26 */
typedef struct { OVec3 position; float radius; } OSphere;
/* geometry.h:29-96 ; planes order NEAR,FAR,LEFT,RIGHT,TOP,BOTTOM,EXTRA0,EXTRA1 */
typedef struct { float xs[8], ys[8], zs[8], ds[8]; OVec3 points[8]; } OFrustum; /* 224 B */
/* geometry.h:99-149 */
typedef struct { float xs[8], ys[8], zs[8], ds[8]; OVec3 points[8]; ODVec3 origin; } OShiftedFrustum; /* 256 B */

enum { O_NEAR = 0, O_FAR, O_LEFT, O_RIGHT, O_TOP, O_BOTTOM, O_EXTRA0, O_EXTRA1 };

static inline OVec3 ov3(float x, float y, float z) { OVec3 r = {x, y, z}; return r; }
static inline ODVec3 odv3(double x, double y, double z) { ODVec3 r = {x, y, z}; return r; }
static inline OQuat oquat(float x, float y, float z, float w) { OQuat r = {x, y, z, w}; return r; }

/* math.cpp:526-530 Vec3::Vec3(const DVec3&) */
static inline OVec3 ov3_from_d(ODVec3 v) { return ov3((float)v.x, (float)v.y, (float)v.z); }
/* math.cpp:133-138 IVec3::IVec3(const DVec3&) — truncation toward zero */
static inline OIVec3 oiv3_from_d(ODVec3 v) { OIVec3 r = {(int)v.x, (int)v.y, (int)v.z}; return r; }

static inline OVec3 ov3_add(OVec3 a, OVec3 b) { return ov3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline OVec3 ov3_sub(OVec3 a, OVec3 b) { return ov3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline OVec3 ov3_neg(OVec3 a) { return ov3(-a.x, -a.y, -a.z); }
static inline OVec3 ov3_muls(OVec3 a, float s) { return ov3(a.x * s, a.y * s, a.z * s); }
/* math.cpp:459 Vec3*Vec3 */
static inline OVec3 ov3_mul(OVec3 a, OVec3 b) { return ov3(a.x * b.x, a.y * b.y, a.z * b.z); }
/* math.cpp:506 */
static inline ODVec3 odv3_sub(ODVec3 a, ODVec3 b) { return odv3(a.x - b.x, a.y - b.y, a.z - b.z); }
/* math.cpp:508 */
static inline ODVec3 odv3_add(ODVec3 a, ODVec3 b) { return odv3(a.x + b.x, a.y + b.y, a.z + b.z); }
/* math.cpp:512 DVec3 + Vec3 */
static inline ODVec3 odv3_addf(ODVec3 a, OVec3 b) { return odv3(a.x + b.x, a.y + b.y, a.z + b.z); }
/* math.cpp:510 DVec3 - Vec3 */
static inline ODVec3 odv3_subf(ODVec3 a, OVec3 b) { return odv3(a.x - b.x, a.y - b.y, a.z - b.z); }
/* math.cpp:496 DVec3 * float */
static inline ODVec3 odv3_muls(ODVec3 a, float s) { return odv3(a.x * s, a.y * s, a.z * s); }
/* math.cpp:498 DVec3 * Vec3 */
static inline ODVec3 odv3_mulv(ODVec3 a, OVec3 b) { return odv3(a.x * b.x, a.y * b.y, a.z * b.z); }

/* math.cpp:1266-1268 */
static inline float ov3_dot(OVec3 a, OVec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* math.cpp:1274-1276 */
static inline OVec3 ov3_cross(OVec3 a, OVec3 b) { return ov3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
/* math.cpp:1278-1280 */
static inline ODVec3 odv3_cross(ODVec3 a, ODVec3 b) { return odv3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
/* math.cpp:367-376 */
static inline OVec3 ov3_normalize(OVec3 v) {
	const float inv_len = 1 / sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
	return ov3(v.x * inv_len, v.y * inv_len, v.z * inv_len);
}

/* math.cpp:164-175 Quat::rotate(Vec3) */
static inline OVec3 oquat_rotate(OQuat q, OVec3 v) {
	const OVec3 qvec = ov3(q.x, q.y, q.z);
	OVec3 uv = ov3_cross(qvec, v);
	OVec3 uuv = ov3_cross(qvec, uv);
	uv = ov3_muls(uv, 2.0f * q.w);
	uuv = ov3_muls(uuv, 2.0f);
	return ov3_add(ov3_add(v, uv), uuv);
}

/* math.cpp:177-188 Quat::rotate(const DVec3&) — fp64; qvec promoted from float, uv *= (2.0 * w) */
static inline ODVec3 oquat_rotate_d(OQuat q, ODVec3 v) {
	const ODVec3 qvec = odv3(q.x, q.y, q.z);
	ODVec3 uv = odv3_cross(qvec, v);
	ODVec3 uuv = odv3_cross(qvec, uv);
	const double s = 2.0 * q.w;
	uv.x *= s; uv.y *= s; uv.z *= s;
	uuv.x *= 2.0; uuv.y *= 2.0; uuv.z *= 2.0;
	return odv3_add(odv3_add(v, uv), uuv);
}

/* math.cpp:694-701 Quat::operator*(Quat) */
static inline OQuat oquat_mul(OQuat a, OQuat r) {
	return oquat(a.w * r.x + r.w * a.x + a.y * r.z - r.y * a.z,
		a.w * r.y + r.w * a.y + a.z * r.x - r.z * a.x,
		a.w * r.z + r.w * a.z + a.x * r.y - r.x * a.y,
		a.w * r.w - a.x * r.x - a.y * r.y - a.z * r.z);
}

/* math.cpp:664-667 */
static inline OQuat oquat_conjugated(OQuat q) { return oquat(q.x, q.y, q.z, -q.w); }

/* math.cpp:194-201 lerp(Vec3, Vec3, float) */
static inline OVec3 ov3_lerp(OVec3 a, OVec3 b, float t) {
	const float invt = 1.0f - t;
	return ov3(a.x * invt + b.x * t, a.y * invt + b.y * t, a.z * invt + b.z * t);
}

/* math.cpp:677-692 scalar nlerp: dot summed ((x+y)+z)+w */
static inline OQuat oquat_nlerp(OQuat q1, OQuat q2, float t) {
	OQuat res;
	const float inv = 1.0f - t;
	if (q1.x * q2.x + q1.y * q2.y + q1.z * q2.z + q1.w * q2.w < 0) t = -t;
	res.x = q1.x * inv + q2.x * t;
	res.y = q1.y * inv + q2.y * t;
	res.z = q1.z * inv + q2.z * t;
	res.w = q1.w * inv + q2.w * t;
	const float l = 1 / sqrtf(res.x * res.x + res.y * res.y + res.z * res.z + res.w * res.w);
	res.x *= l; res.y *= l; res.z *= l; res.w *= l;
	return res;
}

/* simd_math.h:107-123 simd_nlerp: both horizontal sums are hadd(hadd()) = (x+y)+(z+w) */
static inline OQuat oquat_simd_nlerp(OQuat q1, OQuat q2, float t) {
	const float inv = 1.0f - t;
	const float d = (q1.x * q2.x + q1.y * q2.y) + (q1.z * q2.z + q1.w * q2.w);
	if (d < 0) t = -t;
	OQuat q = oquat(q1.x * inv + q2.x * t, q1.y * inv + q2.y * t, q1.z * inv + q2.z * t, q1.w * inv + q2.w * t);
	const float len2 = (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);
	const float l = 1 / sqrtf(len2);
	q.x *= l; q.y *= l; q.z *= l; q.w *= l;
	return q;
}

/* math.cpp:801-807 Transform::compose */
static inline OTransform otransform_compose(const OTransform* a, const OTransform* rhs) {
	OTransform r;
	r.pos = odv3_add(oquat_rotate_d(a->rot, odv3_mulv(rhs->pos, a->scale)), a->pos);
	r.rot = oquat_mul(a->rot, rhs->rot);
	r.scale = ov3_mul(a->scale, rhs->scale);
	return r;
}

/* math.cpp:809-816 Transform::computeLocal(parent, child): conjugated() = (x, y, z, -w) (math.cpp:664-667), DVec3 unary minus
 * (math.cpp:494), Quat::rotate(DVec3) in fp64, DVec3 / Vec3 = double / float per component (math.cpp:502), Vec3 / Vec3 (math.cpp:468) */
static inline OTransform otransform_compute_local(const OTransform* parent, const OTransform* child) {
	const OQuat c = oquat_conjugated(parent->rot);
	const ODVec3 rp = oquat_rotate_d(c, odv3(-parent->pos.x, -parent->pos.y, -parent->pos.z));
	const ODVec3 inv_parent_pos = odv3(rp.x / parent->scale.x, rp.y / parent->scale.y, rp.z / parent->scale.z);
	const ODVec3 rc = oquat_rotate_d(c, child->pos);
	OTransform r;
	r.pos = odv3_add(odv3(rc.x / parent->scale.x, rc.y / parent->scale.y, rc.z / parent->scale.z), inv_parent_pos);
	r.rot = oquat_mul(c, child->rot);
	r.scale = ov3(child->scale.x / parent->scale.x, child->scale.y / parent->scale.y, child->scale.z / parent->scale.z);
	return r;
}

/* math.cpp:763 Transform::compose(const LocalRigidTransform&): { pos + rot.rotate(rhs.pos * scale), rot * rhs.rot, scale } —
 * Vec3 * Vec3 (math.cpp:459) and Quat::rotate(Vec3) in fp32, then DVec3 + Vec3 (math.cpp:512) */
static inline OTransform otransform_compose_rigid(const OTransform* a, OVec3 rhs_pos, OQuat rhs_rot) {
	OTransform r;
	r.pos = odv3_addf(a->pos, oquat_rotate(a->rot, ov3_mul(rhs_pos, a->scale)));
	r.rot = oquat_mul(a->rot, rhs_rot);
	r.scale = a->scale;
	return r;
}

/* math.cpp:859-861 LocalRigidTransform::operator* */
static inline OLocalRigidTransform olrt_mul(OLocalRigidTransform a, OLocalRigidTransform b) {
	OLocalRigidTransform r;
	r.pos = ov3_add(oquat_rotate(a.rot, b.pos), a.pos);
	r.rot = oquat_mul(a.rot, b.rot);
	return r;
}

/* math.cpp:836-841 / model.cpp:24-30 */
static inline OLocalRigidTransform olrt_inverted(OLocalRigidTransform a) {
	OLocalRigidTransform r;
	r.rot = oquat_conjugated(a.rot);
	r.pos = oquat_rotate(r.rot, ov3_neg(a.pos));
	return r;
}

/* math.cpp:843-853 LocalRigidTransform::toDualQuat */
static inline ODualQuat olrt_to_dual_quat(OLocalRigidTransform t) {
	ODualQuat res;
	const OVec3 pos = t.pos;
	const OQuat rot = t.rot;
	res.r = rot;
	res.d = oquat(0.5f * (pos.x * rot.w + pos.y * rot.z - pos.z * rot.y),
		0.5f * (-pos.x * rot.z + pos.y * rot.w + pos.z * rot.x),
		0.5f * (pos.x * rot.y - pos.y * rot.x + pos.z * rot.w),
		-0.5f * (pos.x * rot.x + pos.y * rot.y + pos.z * rot.z));
	return res;
}

/* math.cpp:727-756 Quat::toMatrix + :887-890 Matrix(pos, rot) (setTranslation) = LocalRigidTransform::toMatrix :855-857 */
static inline OMatrix olrt_to_matrix(OLocalRigidTransform t) {
	const OQuat q = t.rot;
	const float fx = q.x + q.x, fy = q.y + q.y, fz = q.z + q.z;
	const float fwx = fx * q.w, fwy = fy * q.w, fwz = fz * q.w;
	const float fxx = fx * q.x, fxy = fy * q.x, fxz = fz * q.x;
	const float fyy = fy * q.y, fyz = fz * q.y, fzz = fz * q.z;
	OMatrix mtx;
	float* m = mtx.m; /* m[col*4 + row] */
	m[0] = 1.0f - (fyy + fzz); m[4] = fxy - fwz;          m[8] = fxz + fwy;
	m[1] = fxy + fwz;          m[5] = 1.0f - (fxx + fzz); m[9] = fyz - fwx;
	m[2] = fxz - fwy;          m[6] = fyz + fwx;          m[10] = 1.0f - (fxx + fyy);
	m[3] = m[7] = m[11] = 0;
	m[12] = t.pos.x; m[13] = t.pos.y; m[14] = t.pos.z; m[15] = 1;
	return mtx;
}

#endif
