/* TEST INFRASTRUCTURE — CPU restatement of the reference's pose evaluation + skinning palette build
 * (see oracle_math.h for the rules of use).
 *
 *   src/renderer/model.cpp:226-237        Model::getRelativePose (bind pose copy)
 *   src/animation/animation.cpp:30-95      AnimationSampler::getRotation (bit unpack, smallest-three, simd_nlerp)
 *   src/animation/animation.cpp:117-204    AnimationSampler::getRelativePose<mask=false, weight>
 *   src/animation/animation.cpp:313-334    unpackChannel (through double) / Animation::getTranslation
 *   src/animation/animation.h:17-43        Time (u32 ticks, 1 s = 32768)
 *   src/renderer/pose.cpp:66-133           Pose::computeAbsolute (SIMD and scalar paths are the same arithmetic)
 *   src/renderer/pipeline.cpp:2680-2745    computeSkeletonDualQuats
 *   src/renderer/model.cpp:103-109,132-137 evaluateSkin / computeSkinMatrices
 *   src/animation/animation_module.cpp:439-472 updateAnimable (driver order + time advance)
 * Root-motion tracks (animation.cpp:33-37, 316) are outside this path's scope (animables with no root motion bone).
 */
#include "oracle.h"

static inline float o_max(float a, float b) { return a > b ? a : b; } /* math.h:472-475 */
static inline float o_min(float a, float b) { return a < b ? a : b; }

static inline uint64_t load_u64(const uint8_t* p) {
	uint64_t v;
	memcpy(&v, p, sizeof(v));
	return v;
}

/* animation.cpp:313-316 */
static float unpack_channel(uint64_t val, float min, float to_float_range, uint32_t bitsize) {
	const uint64_t mask = ((uint64_t)1 << bitsize) - 1;
	return (float)(min + to_float_range * (double)(val & mask));
}

/* animation.cpp:318-334 */
static OVec3 get_translation(const OracleClip* c, uint32_t frame, const OracleTrack* track) {
	const uint32_t offset = c->translations_frame_size_bits * frame + track->offset_bits;
	uint64_t tmp = load_u64(&c->translation_stream[offset / 8]);
	tmp >>= offset & 7;
	OVec3 res;
	res.x = unpack_channel(tmp, track->min[0], track->to_range[0], track->bitsizes[0]);
	tmp >>= track->bitsizes[0];
	res.y = unpack_channel(tmp, track->min[1], track->to_range[1], track->bitsizes[1]);
	tmp >>= track->bitsizes[1];
	res.z = unpack_channel(tmp, track->min[2], track->to_range[2], track->bitsizes[2]);
	return res;
}

/* animation.cpp:30-95 */
static OQuat get_rotation(const OracleClip* c, uint32_t frame, const OracleTrack* track, float t) {
	const uint32_t offset1 = c->rotations_frame_size_bits * frame + track->offset_bits;
	const uint32_t offset2 = offset1 + c->rotations_frame_size_bits;
	uint64_t packed1 = load_u64(&c->rotation_stream[offset1 / 8]);
	packed1 >>= offset1 & 7;
	uint64_t packed2 = load_u64(&c->rotation_stream[offset2 / 8]);
	packed2 >>= offset2 & 7;
	const int is_negative1 = (int)(packed1 & 1);
	packed1 >>= 1;
	const int is_negative2 = (int)(packed2 & 1);
	packed2 >>= 1;
	const uint64_t mask_x = ((uint64_t)1 << track->bitsizes[0]) - 1;
	const uint64_t mask_y = ((uint64_t)1 << track->bitsizes[1]) - 1;
	const uint64_t mask_z = ((uint64_t)1 << track->bitsizes[2]) - 1;
	const uint64_t packed1_y = packed1 >> track->bitsizes[0];
	const uint64_t packed1_z = packed1_y >> track->bitsizes[1];
	const uint64_t packed2_y = packed2 >> track->bitsizes[0];
	const uint64_t packed2_z = packed2_y >> track->bitsizes[1];
	OVec3 v1, v2;
	v1.x = track->min[0] + track->to_range[0] * (float)(packed1 & mask_x);
	v1.y = track->min[1] + track->to_range[1] * (float)(packed1_y & mask_y);
	v1.z = track->min[2] + track->to_range[2] * (float)(packed1_z & mask_z);
	v2.x = track->min[0] + track->to_range[0] * (float)(packed2 & mask_x);
	v2.y = track->min[1] + track->to_range[1] * (float)(packed2_y & mask_y);
	v2.z = track->min[2] + track->to_range[2] * (float)(packed2_z & mask_z);
	const float skipped1 = sqrtf(o_max(0.f, 1 - ov3_dot(v1, v1))) * (is_negative1 ? -1 : 1);
	const float skipped2 = sqrtf(o_max(0.f, 1 - ov3_dot(v2, v2))) * (is_negative2 ? -1 : 1);
	OQuat q1, q2;
	switch (track->skipped_channel) {
		case 0: q1 = oquat(skipped1, v1.x, v1.y, v1.z); q2 = oquat(skipped2, v2.x, v2.y, v2.z); break;
		case 1: q1 = oquat(v1.x, skipped1, v1.y, v1.z); q2 = oquat(v2.x, skipped2, v2.y, v2.z); break;
		case 2: q1 = oquat(v1.x, v1.y, skipped1, v1.z); q2 = oquat(v2.x, v2.y, skipped2, v2.z); break;
		default: q1 = oquat(v1.x, v1.y, v1.z, skipped1); q2 = oquat(v2.x, v2.y, v2.z, skipped2); break;
	}
	return oquat_simd_nlerp(q1, q2, t);
}

/* animation.cpp:117-204, use_mask=false */
static void get_relative_pose(const OracleClip* anim, uint32_t bone_count, uint32_t time_ticks, int use_weight, float weight, OVec3* pos, OQuat* rot) {
	(void)bone_count;
	/* animation.h:27 Time::toFrame: float(value / double(ONE_SECOND) * fps) */
	const float frame = (float)(time_ticks / (double)(1 << 15) * anim->fps);
	/* :131 clamp(frame, 0.f, m_frame_count - 0.00001f) = min(max(v, lo), hi) */
	const float sample = o_min(o_max(frame, 0.f), (float)anim->frame_count - 0.00001f);
	const uint32_t sample_idx = (uint32_t)sample;
	const float t = sample - (float)sample_idx;

	for (uint32_t i = 0; i < anim->n_const_translations; ++i) {
		const OracleConstTranslation* track = &anim->const_translations[i];
		const OVec3 v = ov3(track->value[0], track->value[1], track->value[2]);
		pos[track->bone_index] = use_weight ? ov3_lerp(pos[track->bone_index], v, weight) : v;
	}
	for (uint32_t i = 0; i < anim->n_translations; ++i) {
		const OracleTrack* track = &anim->translations[i];
		const OVec3 anim_pos = ov3_lerp(get_translation(anim, sample_idx, track), get_translation(anim, sample_idx + 1, track), t);
		pos[track->bone_index] = use_weight ? ov3_lerp(pos[track->bone_index], anim_pos, weight) : anim_pos;
	}
	for (uint32_t i = 0; i < anim->n_const_rotations; ++i) {
		const OracleConstRotation* track = &anim->const_rotations[i];
		const OQuat v = oquat(track->value[0], track->value[1], track->value[2], track->value[3]);
		rot[track->bone_index] = use_weight ? oquat_simd_nlerp(rot[track->bone_index], v, weight) : v;
	}
	for (uint32_t i = 0; i < anim->n_rotations; ++i) {
		const OracleTrack* track = &anim->rotations[i];
		const OQuat anim_rot = get_rotation(anim, sample_idx, track, t);
		rot[track->bone_index] = use_weight ? oquat_simd_nlerp(rot[track->bone_index], anim_rot, weight) : anim_rot;
	}
}

void oracle_pose_sample_weighted(const OracleClip* clip, uint32_t bone_count, uint32_t time_ticks, float weight, OVec3* pos, OQuat* rot) {
	/* animation.cpp:294-311 dispatcher: weight < 0.9999f selects the blending instantiation */
	get_relative_pose(clip, bone_count, time_ticks, weight < 0.9999f, weight, pos, rot);
}

/* pose.cpp:66-133 */
void oracle_pose_compute_absolute(const OracleSkeleton* sk, OVec3* pos, OQuat* rot) {
	for (uint32_t i = (uint32_t)sk->first_nonroot_bone_index; i < sk->bone_count; ++i) {
		const int32_t parent = sk->parents[i];
		/* :129-130 (the 4-wide path :71-126 evaluates the same expressions on 4 independent bones) */
		pos[i] = ov3_add(oquat_rotate(rot[parent], pos[i]), pos[parent]);
		rot[i] = oquat_mul(rot[parent], rot[i]);
	}
}

/* pose.cpp:136-146: i runs from the last bone down, so the parent (index < i) is still absolute when bone i is converted */
void oracle_pose_compute_relative(const OracleSkeleton* sk, OVec3* pos, OQuat* rot) {
	for (int32_t i = (int32_t)sk->bone_count - 1; i >= sk->first_nonroot_bone_index; --i) {
		const int32_t parent = sk->parents[i];
		const OQuat c = oquat_conjugated(rot[parent]);
		pos[i] = oquat_rotate(c, ov3_sub(pos[i], pos[parent]));
		rot[i] = oquat_mul(c, rot[i]);
	}
}

/* pose.cpp:30-41 */
void oracle_pose_blend(uint32_t bone_count, OVec3* pos_a, OQuat* rot_a, const OVec3* pos_b, const OQuat* rot_b, float weight) {
	if (weight <= 0.001f) return;
	weight = weight < 0.0f ? 0.0f : (weight > 1.0f ? 1.0f : weight); /* clamp, math.h */
	const float inv = 1.0f - weight;
	for (uint32_t i = 0; i < bone_count; ++i) {
		pos_a[i] = ov3_add(ov3_muls(pos_a[i], inv), ov3_muls(pos_b[i], weight)); /* Vec3 * float + Vec3 * float */
		rot_a[i] = oquat_nlerp(rot_a[i], rot_b[i], weight);                      /* scalar nlerp, math.cpp:677-692 */
	}
}

/* animation_module.cpp:439-456 */
void oracle_pose_evaluate(const OracleSkeleton* sk, const OracleClip* clip, uint32_t time_ticks, OVec3* pos, OQuat* rot) {
	for (uint32_t i = 0; i < sk->bone_count; ++i) { /* model.cpp:226-237 */
		pos[i] = sk->bind_relative[i].pos;
		rot[i] = sk->bind_relative[i].rot;
	}
	get_relative_pose(clip, sk->bone_count, time_ticks, 0, 1.0f, pos, rot);
	oracle_pose_compute_absolute(sk, pos, rot);
}

/* pipeline.cpp:2680-2745: out[j] = toDualQuat({pos[j], rot[j]} * inv_bind[j]) */
void oracle_palette_dual_quats(const OracleSkeleton* sk, const OVec3* pos, const OQuat* rot, ODualQuat* out) {
	for (uint32_t j = 0; j < sk->bone_count; ++j) {
		const OLocalRigidTransform tmp = {pos[j], rot[j]};
		out[j] = olrt_to_dual_quat(olrt_mul(tmp, sk->inverse_bind[j]));
	}
}

/* model.cpp:132-137 */
void oracle_palette_matrices(const OracleSkeleton* sk, const OVec3* pos, const OQuat* rot, OMatrix* out) {
	for (uint32_t j = 0; j < sk->bone_count; ++j) {
		const OLocalRigidTransform tmp = {pos[j], rot[j]};
		out[j] = olrt_to_matrix(olrt_mul(tmp, sk->inverse_bind[j]));
	}
}

/* model.cpp:103-109: M = m0*w.x + m1*w.y + m2*w.z + m3*w.w (math.cpp:1022-1071), then transformPoint (:1231-1235) */
void oracle_skin_vertices(const OMatrix* matrices, const OVec3* vertices, const float* weights4, const int16_t* indices4, OVec3* out, uint32_t n) {
	for (uint32_t v = 0; v < n; ++v) {
		const float* w = weights4 + 4 * v;
		const int16_t* idx = indices4 + 4 * v;
		float m[16];
		for (int e = 0; e < 16; ++e) {
			m[e] = ((matrices[idx[0]].m[e] * w[0] + matrices[idx[1]].m[e] * w[1]) + matrices[idx[2]].m[e] * w[2]) + matrices[idx[3]].m[e] * w[3];
		}
		const OVec3 p = vertices[v];
		out[v].x = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12];
		out[v].y = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13];
		out[v].z = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14];
	}
}

/* animation_module.cpp:458-461; animation.h:21-24,128 */
/* animation_module.cpp:458-469 with Time's operators (animation.h:17-43): fromSeconds = u32(time * 32768), +, -, % on u32 ticks;
 * getLength() = Time::fromSeconds(frame_count / fps) (animation.h:128) */
uint32_t oracle_time_advance(uint32_t time_ticks, float time_delta, float fps, uint32_t frame_count) {
	const uint32_t l = (uint32_t)(((float)frame_count / fps) * (1 << 15));
	if (time_delta > 0) {
		const uint32_t dt = (uint32_t)(time_delta * (1 << 15));
		return (time_ticks + dt) % l;
	}
	const uint32_t dt = (uint32_t)(-time_delta * (1 << 15)) % l;
	return (time_ticks + l - dt) % l;
}

void oracle_animate_instances(const OracleSkeleton* sk, const OracleClip* clips, const uint32_t* clip_index, const uint32_t* time_ticks,
	uint32_t n, OVec3* out_pos, OQuat* out_rot, ODualQuat* out_dq, OMatrix* out_mtx)
{
	const uint32_t B = sk->bone_count;
	OVec3 pos[256];
	OQuat rot[256];
	for (uint32_t i = 0; i < n; ++i) {
		oracle_pose_evaluate(sk, &clips[clip_index[i]], time_ticks[i], pos, rot);
		if (out_pos) memcpy(out_pos + (size_t)i * B, pos, sizeof(OVec3) * B);
		if (out_rot) memcpy(out_rot + (size_t)i * B, rot, sizeof(OQuat) * B);
		if (out_dq) oracle_palette_dual_quats(sk, pos, rot, out_dq + (size_t)i * B);
		if (out_mtx) oracle_palette_matrices(sk, pos, rot, out_mtx + (size_t)i * B);
	}
}

/* ---- element-wise shims over oracle_math.h so that tests can pin each primitive against the reference build ---- */
void oracle_transform_compose(const OTransform* a, const OTransform* b, OTransform* out, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) out[i] = otransform_compose(&a[i], &b[i]);
}
void oracle_quat_mul(const OQuat* a, const OQuat* b, OQuat* out, uint32_t n) { for (uint32_t i = 0; i < n; ++i) out[i] = oquat_mul(a[i], b[i]); }
void oracle_quat_rotate(const OQuat* q, const OVec3* v, OVec3* out, uint32_t n) { for (uint32_t i = 0; i < n; ++i) out[i] = oquat_rotate(q[i], v[i]); }
void oracle_nlerp(const OQuat* a, const OQuat* b, const float* t, OQuat* out, uint32_t n, int simd) {
	for (uint32_t i = 0; i < n; ++i) out[i] = simd ? oquat_simd_nlerp(a[i], b[i], t[i]) : oquat_nlerp(a[i], b[i], t[i]);
}
void oracle_lerp_vec3(const OVec3* a, const OVec3* b, const float* t, OVec3* out, uint32_t n) { for (uint32_t i = 0; i < n; ++i) out[i] = ov3_lerp(a[i], b[i], t[i]); }
void oracle_lrt_mul(const OLocalRigidTransform* a, const OLocalRigidTransform* b, OLocalRigidTransform* out, uint32_t n) { for (uint32_t i = 0; i < n; ++i) out[i] = olrt_mul(a[i], b[i]); }
void oracle_lrt_inverted(const OLocalRigidTransform* a, OLocalRigidTransform* out, uint32_t n) { for (uint32_t i = 0; i < n; ++i) out[i] = olrt_inverted(a[i]); }
void oracle_lrt_to_dual_quat(const OLocalRigidTransform* a, ODualQuat* out, uint32_t n) { for (uint32_t i = 0; i < n; ++i) out[i] = olrt_to_dual_quat(a[i]); }
void oracle_lrt_to_matrix(const OLocalRigidTransform* a, OMatrix* out, uint32_t n) { for (uint32_t i = 0; i < n; ++i) out[i] = olrt_to_matrix(a[i]); }
void oracle_cell_indices(const double* pos, float cell_size, int* out3) {
	const OIVec3 i = oiv3_from_d(odv3_muls(odv3(pos[0], pos[1], pos[2]), 1 / cell_size));
	out3[0] = i.x; out3[1] = i.y; out3[2] = i.z;
}
