/* TEST INFRASTRUCTURE — CPU restatement of the reference's hierarchy propagation (see oracle_math.h).
 *
 *   src/engine/world.cpp:255-282  World::transformEntity — DFS over first_child/next_sibling:
 *                                  child.global = parent.global.compose(child.local_transform)
 *   src/core/math.cpp:801-807     Transform::compose
 *   src/renderer/render_module.cpp:1544-1554 onModelInstanceMoved: radius = bounding_radius * max(scale)
 *
 * The reference is event-driven; the per-node arithmetic does not depend on traversal
 * order (each global is compose(parent global, own local)), so the DFS below yields exactly what any
 * sequence of setTransform() calls on the roots yields.
 */
#include "oracle.h"

#include <stdlib.h>

void oracle_propagate(const int32_t* parents, const OTransform* locals, OTransform* globals, uint32_t n) {
	/* world.h:157-164 Hierarchy{parent, first_child, next_sibling}; built as setParent does (world.cpp:672-676: new child becomes first_child) */
	int32_t* first_child = (int32_t*)malloc(sizeof(int32_t) * n);
	int32_t* next_sibling = (int32_t*)malloc(sizeof(int32_t) * n);
	int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
	for (uint32_t i = 0; i < n; ++i) { first_child[i] = -1; next_sibling[i] = -1; }
	for (uint32_t i = 0; i < n; ++i) {
		const int32_t p = parents[i];
		if (p >= 0) { next_sibling[i] = first_child[p]; first_child[p] = (int32_t)i; }
	}
	for (uint32_t r = 0; r < n; ++r) {
		if (parents[r] >= 0) continue;
		uint32_t sp = 0;
		stack[sp++] = (int32_t)r;
		while (sp) {
			const int32_t e = stack[--sp];
			const OTransform my_transform = globals[e];                    /* :264 */
			for (int32_t child = first_child[e]; child >= 0; child = next_sibling[child]) {
				globals[child] = otransform_compose(&my_transform, &locals[child]); /* :274-276 */
				stack[sp++] = child;                                           /* :277 recurse */
			}
		}
	}
	free(first_child); free(next_sibling); free(stack);
}

void oracle_sphere_radius(const OTransform* globals, const float* bounding_radius, float* out_radius, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) {
		const OVec3 s = globals[i].scale;
		/* render_module.cpp:1552 maximum(scale.x, scale.y, scale.z) — math.h minimum/maximum variadic: max(a, max(b, c)) */
		const float bc = s.y > s.z ? s.y : s.z;
		const float m = s.x > bc ? s.x : bc;
		out_radius[i] = bounding_radius[i] * m;
	}
}

/* World::getRelativeMatrix, world.cpp:370-377:
 *   Matrix mtx = transform.rot.toMatrix();            math.cpp:727-756
 *   mtx.setTranslation(Vec3(transform.pos - base));   DVec3 - DVec3 in fp64 (math.cpp:505), then narrowed per component (math.cpp:443)
 *   mtx.multiply3x3(transform.scale);                 math.cpp:1207-1217: column k scaled by scale[k], rows x,y,z only */
void oracle_relative_matrices(const OTransform* globals, const double* base_pos3, OMatrix* out, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) {
		const OTransform* t = &globals[i];
		OLocalRigidTransform r;
		r.pos.x = (float)(t->pos.x - base_pos3[0]);
		r.pos.y = (float)(t->pos.y - base_pos3[1]);
		r.pos.z = (float)(t->pos.z - base_pos3[2]);
		r.rot = t->rot;
		OMatrix m = olrt_to_matrix(r); /* toMatrix + setTranslation: same stores as Matrix(pos, rot) */
		m.m[0] *= t->scale.x; m.m[1] *= t->scale.x; m.m[2] *= t->scale.x;
		m.m[4] *= t->scale.y; m.m[5] *= t->scale.y; m.m[6] *= t->scale.y;
		m.m[8] *= t->scale.z; m.m[9] *= t->scale.z; m.m[10] *= t->scale.z;
		out[i] = m;
	}
}

void oracle_compute_locals(const int32_t* parents, const OTransform* globals, OTransform* locals, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) {
		if (parents[i] < 0) continue; /* world.cpp:267: only entities with a valid parent */
		locals[i] = otransform_compute_local(&globals[parents[i]], &globals[i]);
	}
}

void oracle_transform_compute_local(const OTransform* parent, const OTransform* child, OTransform* out, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) out[i] = otransform_compute_local(&parent[i], &child[i]);
}

/* render_module.cpp:399-403 */
void oracle_bone_attachments(const OTransform* parent_transforms, const float* bone7, const float* relative7, const float* original_scale3,
	OTransform* out, uint32_t n)
{
	for (uint32_t i = 0; i < n; ++i) {
		const float* b = bone7 + 7 * (size_t)i;
		const float* r = relative7 + 7 * (size_t)i;
		const OLocalRigidTransform bone = {ov3(b[0], b[1], b[2]), oquat(b[3], b[4], b[5], b[6])};
		const OLocalRigidTransform rel = {ov3(r[0], r[1], r[2]), oquat(r[3], r[4], r[5], r[6])};
		const OLocalRigidTransform local = olrt_mul(bone, rel);
		OTransform res = otransform_compose_rigid(&parent_transforms[i], local.pos, local.rot);
		res.scale = ov3(original_scale3[3 * i], original_scale3[3 * i + 1], original_scale3[3 * i + 2]);
		out[i] = res;
	}
}
