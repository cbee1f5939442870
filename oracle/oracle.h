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

/* ---- sort keys, LOD selection, auto-instancing, radix sort (oracle_sortkeys.c; pipeline.cpp:53-143, 452-523, 3789-4144) ---- */
typedef struct {
	float lod_distances[4];  /* Model::m_lod_distances (squared), model.h:234 */
	int32_t lod_from[5];     /* Model::m_lod_indices[].from / .to, model.h:129-133,233 */
	int32_t lod_to[5];
	uint32_t mesh_base;      /* first entry of this model in the mesh table */
	uint32_t mesh_count;
} OracleSkModel; /* 64 B */
typedef struct {
	uint32_t sort_key;       /* MeshMaterial::sort_key (model.h:65; RenderModule::computeSortKey) */
	uint32_t material_index; /* MeshMaterial::material_index */
	float lod;               /* Mesh::lod (model.h:120) */
	uint8_t layer;           /* Material::getLayer() */
	uint8_t skinned;         /* Mesh::type == SKINNED */
	uint16_t pad;
} OracleSkMesh; /* 16 B */
typedef struct {
	double camera_pos[3];    /* view.cp.pos */
	double lod_ref_point[3]; /* m_viewport.pos */
	float time_delta;
	float lod_multiplier;    /* Renderer::getLODMultiplier() */
	uint32_t frame_number;
	uint32_t is_shadow;
	uint32_t max_sort_key;   /* Renderer::getMaxSortKey() */
	uint32_t pad;
	uint32_t bucket_map[256];     /* pipeline.cpp:3803-3812 ([255] unused) */
	uint8_t layer_to_bucket[256]; /* View::layer_to_bucket */
} OracleSkView;

uint32_t oracle_float_flip(uint32_t float_bits_value);
uint64_t oracle_make_mesh_sort_key(uint32_t mesh_sort_key, uint8_t bucket);
uint64_t oracle_make_depth_sort_key(float depth_squared, uint8_t bucket);
uint64_t oracle_make_autoinstanced_sort_key(int32_t instancer_index, uint8_t bucket);
uint64_t oracle_make_decal_sort_key(uint32_t material_sort_key, uint8_t bucket);
uint64_t oracle_make_decal_sort_value(int32_t entity);
uint64_t oracle_make_curve_decal_sort_value(int32_t entity);
uint64_t oracle_make_skinned_sort_value(int32_t entity, uint32_t mesh_idx);
uint64_t oracle_make_mesh_sort_value(int32_t entity, uint32_t mesh_idx);
uint64_t oracle_make_autoinstanced_sort_value(uint32_t batch_idx, uint32_t instancer_idx);
uint32_t oracle_lod_mesh_indices(const float* lod_distances4, float squared_distance);
void oracle_radix_sort(uint64_t* keys_io, uint64_t* values_io, uint32_t size);
void oracle_radix_sort_ex(uint64_t* keys_io, uint64_t* values_io, uint32_t size, int reference_copy_back);
int oracle_create_sort_keys(const uint32_t* visible_ids, const uint8_t* visible_types, uint32_t n_visible, const OTransform* transforms,
	const uint32_t* model_of, float* lod, const uint8_t* flags, uint32_t* pose_frame, const uint32_t* decal_sort_key, const uint8_t* decal_layer,
	const OracleSkModel* models, const OracleSkMesh* meshes, const OracleSkView* view,
	uint64_t* keys, uint64_t* values, uint32_t cap_keys, uint32_t* n_keys,
	uint32_t* group_count, uint32_t* group_offset, uint64_t* group_renderables, uint8_t* instance_data48, uint32_t cap_instances, uint32_t* n_instances,
	uint32_t* pose_list, uint32_t cap_pose, uint32_t* n_pose, uint32_t* dirty_list, uint32_t cap_dirty, uint32_t* n_dirty);

#ifdef __cplusplus
}
#endif
#endif
