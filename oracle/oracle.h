/* TEST INFRASTRUCTURE — public C API of the CPU oracle (see oracle_math.h for the rules of use). */
#ifndef ORACLE_H
#define ORACLE_H

#include "oracle_math.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- culling (oracle_cull.c) ---- */
typedef struct OracleCulling OracleCulling;

typedef struct {
	uint32_t pages_total, pages_filtered, pages_tested, pages_inside, pages_outside;
	uint32_t entities_total, entities_tested, entities_inside;
	uint32_t visible, visible_tested;
} OracleCullStats;

void oracle_frustum_perspective(OShiftedFrustum* f, const double* position, const float* direction, const float* up,
	float fov, float ratio, float near_distance, float far_distance);
void oracle_frustum_ortho(OShiftedFrustum* f, const double* position, const float* direction, const float* up,
	float width, float height, float near_distance, float far_distance);
void oracle_frustum_from_viewport(OShiftedFrustum* f, int is_ortho, float fov, float ortho_size, int w, int h, const double* pos,
	const float* rot4, float near_distance, float far_distance);
int oracle_frustum_contains_aabb(const OShiftedFrustum* f, ODVec3 pos, OVec3 size);
int oracle_frustum_intersects_aabb(const OShiftedFrustum* f, ODVec3 pos, OVec3 size);
void oracle_frustum_get_relative(const OShiftedFrustum* f, ODVec3 origin, OFrustum* res);

OracleCulling* oracle_culling_create(void);
void oracle_culling_destroy(OracleCulling* cs);
void oracle_culling_add(OracleCulling* cs, int32_t entity, uint8_t type, const double* pos3, float radius);
void oracle_culling_remove(OracleCulling* cs, int32_t entity);
void oracle_culling_set_position(OracleCulling* cs, int32_t entity, const double* pos3);
void oracle_culling_set_radius(OracleCulling* cs, int32_t entity, float radius);
void oracle_culling_set(OracleCulling* cs, int32_t entity, const double* pos3, float radius);
float oracle_culling_get_radius(const OracleCulling* cs, int32_t entity);
int oracle_culling_is_added(const OracleCulling* cs, int32_t entity);
uint32_t oracle_culling_cull(const OracleCulling* cs, const OShiftedFrustum* frustum, int type, uint32_t* out_ids, uint8_t* out_types,
	uint32_t cap, OracleCullStats* stats);
uint32_t oracle_culling_page_count(const OracleCulling* cs);
void oracle_culling_get_page(const OracleCulling* cs, uint32_t idx, double* origin3, int* indices3, uint8_t* type, uint8_t* is_big, int* count,
	float* spheres4, int32_t* entities);
void oracle_culling_add_many(OracleCulling* cs, const int32_t* entities, const uint8_t* types, const double* pos3, const float* radius, uint32_t n);
void oracle_culling_set_many(OracleCulling* cs, const int32_t* entities, const double* pos3, const float* radius, uint32_t n);
void oracle_culling_set_position_many(OracleCulling* cs, const int32_t* entities, const double* pos3, uint32_t n);
void oracle_culling_set_radius_many(OracleCulling* cs, const int32_t* entities, const float* radius, uint32_t n);
void oracle_culling_remove_many(OracleCulling* cs, const int32_t* entities, uint32_t n);
void oracle_rng_floats(uint32_t* u, uint32_t* v, uint32_t n, float* out);

/* ---- hierarchy propagation (oracle_propagate.c) ---- */
/* parents[i] = parent node index or -1; locals/globals are 56-byte Transforms (math.h:306-327).
 * Roots take globals[i] as input (their world transform); every other node is overwritten. */
void oracle_propagate(const int32_t* parents, const OTransform* locals, OTransform* globals, uint32_t n);
/* render_module.cpp:1544-1554: radius = bounding_radius * max(scale.xyz) */
void oracle_sphere_radius(const OTransform* globals, const float* bounding_radius, float* out_radius, uint32_t n);
/* The update_local branch of World::transformEntity (world.cpp:267-270) for every non-root node:
 * locals[i] = Transform::computeLocal(globals[parents[i]], globals[i]) (math.cpp:809-816); roots keep locals[i] untouched. */
void oracle_compute_locals(const int32_t* parents, const OTransform* globals, OTransform* locals, uint32_t n);
void oracle_transform_compute_local(const OTransform* parent, const OTransform* child, OTransform* out, uint32_t n);
/* RenderModuleImpl::updateBoneAttachment (render_module.cpp:377-405): world transform of an entity attached to a bone of a posed model
 * instance = parent_entity_transform.compose(bone_transform * relative_transform), scale replaced by the entity's own.
 * bone7 / relative7: pos xyz + rot xyzw per attachment. */
void oracle_bone_attachments(const OTransform* parent_transforms, const float* bone7, const float* relative7, const float* original_scale3,
	OTransform* out, uint32_t n);
/* world.cpp:370-377 World::getRelativeMatrix: rot.toMatrix(), translation = Vec3(pos - base_pos), multiply3x3(scale) */
void oracle_relative_matrices(const OTransform* globals, const double* base_pos3, OMatrix* out, uint32_t n);

/* ---- animation / pose / palette / skin (oracle_anim.c) ---- */
/* animation.h:92-118 track descriptors in a flat, pointer-free form shared with the product ABI */
typedef struct {
	uint16_t bone_index;
	uint16_t offset_bits;
	uint8_t bitsizes[3];
	uint8_t skipped_channel; /* rotations only */
	float min[3];
	float to_range[3];
} OracleTrack; /* 32 B */

typedef struct { uint16_t bone_index; uint16_t pad; float value[3]; } OracleConstTranslation; /* 16 B */
typedef struct { uint16_t bone_index; uint16_t pad; float value[4]; } OracleConstRotation;    /* 20 B */

typedef struct {
	float fps;
	uint32_t frame_count;
	uint32_t translations_frame_size_bits, rotations_frame_size_bits;
	uint32_t n_translations, n_const_translations, n_rotations, n_const_rotations;
	const OracleTrack* translations;
	const OracleConstTranslation* const_translations;
	const OracleTrack* rotations;
	const OracleConstRotation* const_rotations;
	const uint8_t* translation_stream; /* (frame_count+1) frames, +8 B tail padding (animation.cpp:439) */
	const uint8_t* rotation_stream;
} OracleClip;

typedef struct {
	uint32_t bone_count;
	int32_t first_nonroot_bone_index;
	const int16_t* parents;                   /* model.h m_parents */
	const OLocalRigidTransform* bind_relative; /* Bone::relative_transform */
	const OLocalRigidTransform* inverse_bind;  /* getInverseBindTransform(i) */
} OracleSkeleton;

/* Model::getRelativePose + Animation::getRelativePose (weight=1) + Pose::computeAbsolute for one instance;
 * time_ticks is Time::raw() (1 s = 32768). pos: bone_count Vec3, rot: bone_count Quat. */
void oracle_pose_evaluate(const OracleSkeleton* sk, const OracleClip* clip, uint32_t time_ticks, OVec3* pos, OQuat* rot);
/* Animation::getRelativePose with weight < 0.9999 blending onto an existing relative pose */
void oracle_pose_sample_weighted(const OracleClip* clip, uint32_t bone_count, uint32_t time_ticks, float weight, OVec3* pos, OQuat* rot);
void oracle_pose_compute_absolute(const OracleSkeleton* sk, OVec3* pos, OQuat* rot);
/* pose.cpp:136-146 Pose::computeRelative: absolute -> parent-relative, in place (bones visited from the last one down) */
void oracle_pose_compute_relative(const OracleSkeleton* sk, OVec3* pos, OQuat* rot);
/* pose.cpp:30-41 Pose::blend: a = a * (1 - w) + b * w positions, scalar nlerp rotations; w <= 0.001 leaves a untouched, w clamped to [0,1] */
void oracle_pose_blend(uint32_t bone_count, OVec3* pos_a, OQuat* rot_a, const OVec3* pos_b, const OQuat* rot_b, float weight);
/* pipeline.cpp:2680-2745 */
void oracle_palette_dual_quats(const OracleSkeleton* sk, const OVec3* pos, const OQuat* rot, ODualQuat* out);
/* model.cpp:132-137 */
void oracle_palette_matrices(const OracleSkeleton* sk, const OVec3* pos, const OQuat* rot, OMatrix* out);
/* model.cpp:103-109 evaluateSkin over n vertices */
void oracle_skin_vertices(const OMatrix* matrices, const OVec3* vertices, const float* weights4, const int16_t* indices4, OVec3* out, uint32_t n);
/* animation_module.cpp:458-461 time advance for time_delta > 0 */
uint32_t oracle_time_advance(uint32_t time_ticks, float time_delta, float fps, uint32_t frame_count);
/* batched driver: instances [0,n) each with its own clip index + time; outputs n*bone_count entries */
void oracle_animate_instances(const OracleSkeleton* sk, const OracleClip* clips, const uint32_t* clip_index, const uint32_t* time_ticks,
	uint32_t n, OVec3* out_pos, OQuat* out_rot, ODualQuat* out_dq, OMatrix* out_mtx);

#ifdef __cplusplus
}
#endif
#endif
