/* TEST INFRASTRUCTURE — CPU restatement of the reference's CullingSystem (see oracle_math.h header).
 *
 * Restates, line for line:
 *   src/renderer/culling_system.cpp:23-63    CellIndices / CellIndicesHasher / CellPage
 *   src/renderer/culling_system.cpp:98-258   addToCell / add / remove / setPosition / set / setRadius
 *   src/renderer/culling_system.cpp:260-369  doCulling / cullInternal
 *   src/core/geometry.cpp:99-178             ShiftedFrustum::containsAABB / getRelative / intersectsAABB
 *   src/core/geometry.cpp:311-351,412-418,470-499  frustum construction
 * Cross-checked against the reference's own compiled culling_system.cpp (oracle/_ref) by
 * tests/test_oracle_ref.py: identical sorted visible ids on seeded scenes.
 */
#include "oracle.h"

#include <stdlib.h>

/* ------------------------------------------------------------------------------------------------ */
/* frustum construction (host side in the reference; here so tests can build inputs)                  */
/* ------------------------------------------------------------------------------------------------ */

/* geometry.cpp:421-427 ShiftedFrustum::setPlane(side, normal, point) */
static void sf_set_plane(OShiftedFrustum* f, int side, OVec3 normal, OVec3 point) {
	f->xs[side] = normal.x;
	f->ys[side] = normal.y;
	f->zs[side] = normal.z;
	f->ds[side] = -ov3_dot(point, normal);
}

/* geometry.cpp:412-418 Frustum::setPlane(side, normal, point) */
static void f_set_plane(OFrustum* f, int side, OVec3 normal, OVec3 point) {
	f->xs[side] = normal.x;
	f->ys[side] = normal.y;
	f->zs[side] = normal.z;
	f->ds[side] = -ov3_dot(point, normal);
}

/* geometry.cpp:326-337 ShiftedFrustum::setPlanesFromPoints */
static void sf_set_planes_from_points(OShiftedFrustum* f) {
	const OVec3* p = f->points;
	const OVec3 normal_near = ov3_neg(ov3_normalize(ov3_cross(ov3_sub(p[0], p[1]), ov3_sub(p[0], p[2]))));
	const OVec3 normal_far = ov3_normalize(ov3_cross(ov3_sub(p[4], p[5]), ov3_sub(p[4], p[6])));
	sf_set_plane(f, O_EXTRA0, normal_near, p[0]);
	sf_set_plane(f, O_EXTRA1, normal_near, p[0]);
	sf_set_plane(f, O_NEAR, normal_near, p[0]);
	sf_set_plane(f, O_FAR, normal_far, p[4]);
	sf_set_plane(f, O_LEFT, ov3_normalize(ov3_cross(ov3_sub(p[1], p[2]), ov3_sub(p[1], p[5]))), p[1]);
	sf_set_plane(f, O_RIGHT, ov3_neg(ov3_normalize(ov3_cross(ov3_sub(p[0], p[3]), ov3_sub(p[0], p[4])))), p[0]);
	sf_set_plane(f, O_TOP, ov3_normalize(ov3_cross(ov3_sub(p[0], p[1]), ov3_sub(p[0], p[4]))), p[0]);
	sf_set_plane(f, O_BOTTOM, ov3_normalize(ov3_cross(ov3_sub(p[2], p[3]), ov3_sub(p[2], p[6]))), p[2]);
}

/* geometry.cpp:339-367 setPoints<T> */
static void sf_set_points(OShiftedFrustum* f, OVec3 near_center, OVec3 far_center, OVec3 right_near, OVec3 up_near,
	OVec3 right_far, OVec3 up_far, float vminx, float vminy, float vmaxx, float vmaxy)
{
	OVec3* p = f->points;
	p[0] = ov3_add(ov3_add(near_center, ov3_muls(right_near, vmaxx)), ov3_muls(up_near, vmaxy));
	p[1] = ov3_add(ov3_add(near_center, ov3_muls(right_near, vminx)), ov3_muls(up_near, vmaxy));
	p[2] = ov3_add(ov3_add(near_center, ov3_muls(right_near, vminx)), ov3_muls(up_near, vminy));
	p[3] = ov3_add(ov3_add(near_center, ov3_muls(right_near, vmaxx)), ov3_muls(up_near, vminy));
	p[4] = ov3_add(ov3_add(far_center, ov3_muls(right_far, vmaxx)), ov3_muls(up_far, vmaxy));
	p[5] = ov3_add(ov3_add(far_center, ov3_muls(right_far, vminx)), ov3_muls(up_far, vmaxy));
	p[6] = ov3_add(ov3_add(far_center, ov3_muls(right_far, vminx)), ov3_muls(up_far, vminy));
	p[7] = ov3_add(ov3_add(far_center, ov3_muls(right_far, vmaxx)), ov3_muls(up_far, vminy));
	sf_set_planes_from_points(f);
}

/* geometry.cpp:470-499 ShiftedFrustum::computePerspective (viewport {-1,-1}..{1,1}, :515-525) */
void oracle_frustum_perspective(OShiftedFrustum* f, const double* position, const float* direction, const float* up_,
	float fov, float ratio, float near_distance, float far_distance)
{
	const OVec3 dir = ov3(direction[0], direction[1], direction[2]);
	const OVec3 up = ov3(up_[0], up_[1], up_[2]);
	memset(f, 0, sizeof(*f));
	const float scale = tanf(fov * 0.5f);
	const OVec3 right = ov3_cross(dir, up);
	const OVec3 up_near = ov3_muls(ov3_muls(up, near_distance), scale);
	const OVec3 right_near = ov3_muls(right, near_distance * scale * ratio);
	const OVec3 up_far = ov3_muls(ov3_muls(up, far_distance), scale);
	const OVec3 right_far = ov3_muls(right, far_distance * scale * ratio);
	const OVec3 z = ov3_normalize(dir);
	const OVec3 near_center = ov3_muls(z, near_distance);
	const OVec3 far_center = ov3_muls(z, far_distance);
	f->origin = odv3(position[0], position[1], position[2]);
	sf_set_points(f, near_center, far_center, right_near, up_near, right_far, up_far, -1, -1, 1, 1);
}

/* geometry.cpp:390-409 ShiftedFrustum::computeOrtho */
void oracle_frustum_ortho(OShiftedFrustum* f, const double* position, const float* direction, const float* up_,
	float width, float height, float near_distance, float far_distance)
{
	const OVec3 dir = ov3(direction[0], direction[1], direction[2]);
	const OVec3 up = ov3(up_[0], up_[1], up_[2]);
	memset(f, 0, sizeof(*f));
	const OVec3 z = ov3_normalize(dir);
	f->origin = odv3(position[0], position[1], position[2]);
	const OVec3 near_center = ov3_muls(ov3_neg(z), near_distance);
	const OVec3 far_center = ov3_muls(ov3_neg(z), far_distance);
	const OVec3 x = ov3_muls(ov3_normalize(ov3_cross(up, z)), width);
	const OVec3 y = ov3_muls(ov3_normalize(ov3_cross(z, x)), height);
	sf_set_points(f, near_center, far_center, x, y, x, y, -1, -1, 1, 1);
}

/* geometry.cpp:793-818 Viewport::getFrustum(): the frustum is built at the origin from the camera rotation (Quat * Vec3 = rotate,
 * math.cpp:721-724), ratio = h > 0 ? w / (float)h : 1, and then ret.origin = pos.  The builders never read `position` except to store
 * it, so building at `pos` directly gives the same bytes. */
void oracle_frustum_from_viewport(OShiftedFrustum* f, int is_ortho, float fov, float ortho_size, int w, int h, const double* pos,
	const float* rot4, float near_distance, float far_distance)
{
	const OQuat rot = oquat(rot4[0], rot4[1], rot4[2], rot4[3]);
	const float ratio = h > 0 ? w / (float)h : 1;
	const OVec3 up = oquat_rotate(rot, ov3(0, 1, 0));
	if (is_ortho) {
		const OVec3 dir = oquat_rotate(rot, ov3(0, 0, 1));
		oracle_frustum_ortho(f, pos, &dir.x, &up.x, ortho_size * ratio, ortho_size, near_distance, far_distance);
		return;
	}
	const OVec3 dir = oquat_rotate(rot, ov3(0, 0, -1));
	oracle_frustum_perspective(f, pos, &dir.x, &up.x, fov, ratio, near_distance, far_distance);
}

/* geometry.cpp:99-118 ShiftedFrustum::containsAABB */
int oracle_frustum_contains_aabb(const OShiftedFrustum* f, ODVec3 pos, OVec3 size) {
	const OVec3 rel_pos = ov3_from_d(odv3_sub(pos, f->origin));
	const OVec3 box[2] = {rel_pos, ov3_add(rel_pos, size)};
	for (int i = 0; i < 6; ++i) {
		const int px = (int)(f->xs[i] < 0.0f);
		const int py = (int)(f->ys[i] < 0.0f);
		const int pz = (int)(f->zs[i] < 0.0f);
		const float dp = (f->xs[i] * box[px].x) + (f->ys[i] * box[py].y) + (f->zs[i] * box[pz].z);
		if (dp < -f->ds[i]) return 0;
	}
	return 1;
}

/* geometry.cpp:159-178 ShiftedFrustum::intersectsAABB */
int oracle_frustum_intersects_aabb(const OShiftedFrustum* f, ODVec3 pos, OVec3 size) {
	const OVec3 rel_pos = ov3_from_d(odv3_sub(pos, f->origin));
	const OVec3 box[2] = {rel_pos, ov3_add(rel_pos, size)};
	for (int i = 0; i < 6; ++i) {
		const int px = (int)(f->xs[i] > 0.0f);
		const int py = (int)(f->ys[i] > 0.0f);
		const int pz = (int)(f->zs[i] > 0.0f);
		const float dp = (f->xs[i] * box[px].x) + (f->ys[i] * box[py].y) + (f->zs[i] * box[pz].z);
		if (dp < -f->ds[i]) return 0;
	}
	return 1;
}

/* geometry.cpp:121-149 ShiftedFrustum::getRelative */
void oracle_frustum_get_relative(const OShiftedFrustum* f, ODVec3 origin, OFrustum* res) {
	const OVec3 offset = ov3_from_d(odv3_sub(f->origin, origin));
	memcpy(res->points, f->points, sizeof(f->points));
	const OVec3 n_near = ov3(f->xs[O_NEAR], f->ys[O_NEAR], f->zs[O_NEAR]);
	const OVec3 n_far = ov3(f->xs[O_FAR], f->ys[O_FAR], f->zs[O_FAR]);
	const OVec3 n_left = ov3(f->xs[O_LEFT], f->ys[O_LEFT], f->zs[O_LEFT]);
	const OVec3 n_right = ov3(f->xs[O_RIGHT], f->ys[O_RIGHT], f->zs[O_RIGHT]);
	const OVec3 n_top = ov3(f->xs[O_TOP], f->ys[O_TOP], f->zs[O_TOP]);
	const OVec3 n_bottom = ov3(f->xs[O_BOTTOM], f->ys[O_BOTTOM], f->zs[O_BOTTOM]);
	f_set_plane(res, O_EXTRA0, n_near, ov3_add(f->points[0], offset));
	f_set_plane(res, O_EXTRA1, n_near, ov3_add(f->points[0], offset));
	f_set_plane(res, O_NEAR, n_near, ov3_add(f->points[0], offset));
	f_set_plane(res, O_FAR, n_far, ov3_add(f->points[4], offset));
	f_set_plane(res, O_LEFT, n_left, ov3_add(f->points[1], offset));
	f_set_plane(res, O_RIGHT, n_right, ov3_add(f->points[0], offset));
	f_set_plane(res, O_TOP, n_top, ov3_add(f->points[0], offset));
	f_set_plane(res, O_BOTTOM, n_bottom, ov3_add(f->points[2], offset));
	for (int i = 0; i < 8; ++i) res->points[i] = ov3_add(res->points[i], offset);
}

/* ------------------------------------------------------------------------------------------------ */
/* CullingSystemImpl                                                                                 */
/* ------------------------------------------------------------------------------------------------ */

/* culling_system.cpp:23-38 */
typedef struct {
	OIVec3 pos;
	uint8_t type;
	uint8_t is_big;
} CellIndices;

/* culling_system.cpp:25-31: pos(pos * (1 / cell_size)) — DVec3 * float (math.cpp:496), then IVec3(DVec3) trunc */
static CellIndices cell_indices(ODVec3 pos, float cell_size, uint8_t type, int is_big) {
	CellIndices c;
	c.pos = oiv3_from_d(odv3_muls(pos, 1 / cell_size));
	c.type = type;
	c.is_big = (uint8_t)(is_big ? 1 : 0);
	return c;
}

static int cell_indices_eq(const CellIndices* a, const CellIndices* b) {
	return a->pos.x == b->pos.x && a->pos.y == b->pos.y && a->pos.z == b->pos.z && a->type == b->type && a->is_big == b->is_big;
}

/* culling_system.cpp:41-48 */
static uint32_t cell_hash(const CellIndices* i) {
	return (uint32_t)i->pos.x * 73856093u + (uint32_t)i->pos.y * 19349663u + (uint32_t)i->pos.z * 83492791u;
}

enum { PAGE_SIZE = 4096 };

/* culling_system.cpp:51-63 */
typedef struct CellPage CellPage;
struct CellPageHeader {
	CellPage* next;
	CellPage* prev;
	ODVec3 origin;
	CellIndices indices;
	int count;
};
enum { MAX_COUNT = (PAGE_SIZE - sizeof(struct CellPageHeader)) / (sizeof(OSphere) + sizeof(int32_t)) };
struct CellPage {
	struct CellPageHeader header;
	OSphere spheres[MAX_COUNT];
	int32_t entities[MAX_COUNT];
};

typedef struct {
	CellIndices key;
	CellPage* value;
	uint8_t state; /* 0 empty, 1 used, 2 tombstone */
} MapSlot;

struct OracleCulling {
	MapSlot* map;
	uint32_t map_cap, map_used, map_filled;
	CellPage** cells; /* m_cells: one entry per PAGE */
	uint32_t cells_size, cells_cap;
	OSphere** entity_to_cell;
	uint32_t e2c_size, e2c_cap;
	float cell_size;
};

static CellPage* page_alloc(void) {
	CellPage* p = (CellPage*)aligned_alloc(PAGE_SIZE, PAGE_SIZE);
	memset(p, 0, sizeof(struct CellPageHeader));
	return p;
}

static MapSlot* map_find(const OracleCulling* cs, const CellIndices* k) {
	if (!cs->map_cap) return NULL;
	uint32_t i = cell_hash(k) & (cs->map_cap - 1);
	for (;;) {
		MapSlot* s = &cs->map[i];
		if (s->state == 0) return NULL;
		if (s->state == 1 && cell_indices_eq(&s->key, k)) return s;
		i = (i + 1) & (cs->map_cap - 1);
	}
}

static void map_insert_nogrow(OracleCulling* cs, const CellIndices* k, CellPage* v) {
	uint32_t i = cell_hash(k) & (cs->map_cap - 1);
	for (;;) {
		MapSlot* s = &cs->map[i];
		if (s->state != 1) {
			if (s->state == 0) ++cs->map_filled;
			s->key = *k; s->value = v; s->state = 1;
			++cs->map_used;
			return;
		}
		i = (i + 1) & (cs->map_cap - 1);
	}
}

static void map_insert(OracleCulling* cs, const CellIndices* k, CellPage* v) {
	if ((cs->map_filled + 1) * 2 > cs->map_cap) {
		MapSlot* old = cs->map;
		const uint32_t old_cap = cs->map_cap;
		cs->map_cap = old_cap ? old_cap * 2 : 1024;
		while (cs->map_used * 4 > cs->map_cap) cs->map_cap *= 2;
		cs->map = (MapSlot*)calloc(cs->map_cap, sizeof(MapSlot));
		cs->map_used = cs->map_filled = 0;
		for (uint32_t i = 0; i < old_cap; ++i) if (old[i].state == 1) map_insert_nogrow(cs, &old[i].key, old[i].value);
		free(old);
	}
	map_insert_nogrow(cs, k, v);
}

static void map_erase(OracleCulling* cs, const CellIndices* k) {
	MapSlot* s = map_find(cs, k);
	if (s) { s->state = 2; --cs->map_used; }
}

static void cells_push(OracleCulling* cs, CellPage* p) {
	if (cs->cells_size == cs->cells_cap) {
		cs->cells_cap = cs->cells_cap ? cs->cells_cap * 2 : 256;
		cs->cells = (CellPage**)realloc(cs->cells, sizeof(CellPage*) * cs->cells_cap);
	}
	cs->cells[cs->cells_size++] = p;
}

/* array.h swapAndPopItem: find, move last into its place */
static void cells_swap_and_pop(OracleCulling* cs, CellPage* p) {
	for (uint32_t i = 0; i < cs->cells_size; ++i) {
		if (cs->cells[i] == p) {
			cs->cells[i] = cs->cells[cs->cells_size - 1];
			--cs->cells_size;
			return;
		}
	}
}

OracleCulling* oracle_culling_create(void) {
	OracleCulling* cs = (OracleCulling*)calloc(1, sizeof(OracleCulling));
	cs->cell_size = 300.0f; /* culling_system.cpp:75 */
	return cs;
}

void oracle_culling_destroy(OracleCulling* cs) {
	if (!cs) return;
	for (uint32_t i = 0; i < cs->cells_size; ++i) free(cs->cells[i]);
	free(cs->cells); free(cs->map); free(cs->entity_to_cell); free(cs);
}

/* culling_system.cpp:98-128 addToCell */
static OSphere* add_to_cell(OracleCulling* cs, CellPage* cell, int32_t entity, ODVec3 pos, float radius) {
	const OVec3 rel_pos = ov3_from_d(odv3_sub(pos, cell->header.origin));
	const int count = cell->header.count;
	if (count < MAX_COUNT - 1) {
		cell->spheres[count].position = rel_pos;
		cell->spheres[count].radius = radius;
		cell->entities[count] = entity;
		++cell->header.count;
		return &cell->spheres[count];
	}
	CellPage* new_cell = page_alloc();
	new_cell->header.origin = cell->header.origin;
	new_cell->header.indices = cell->header.indices;
	new_cell->header.next = cell;
	new_cell->header.prev = cell->header.prev;
	new_cell->header.next->header.prev = new_cell;
	if (new_cell->header.prev) new_cell->header.prev->header.next = new_cell;
	cells_push(cs, new_cell);
	if (!new_cell->header.prev) map_find(cs, &new_cell->header.indices)->value = new_cell;
	new_cell->spheres[0].position = rel_pos;
	new_cell->spheres[0].radius = radius;
	new_cell->entities[0] = entity;
	new_cell->header.count = 1;
	return &new_cell->spheres[0];
}

/* culling_system.cpp:131-157 add */
void oracle_culling_add(OracleCulling* cs, int32_t entity, uint8_t type, const double* p, float radius) {
	const ODVec3 pos = odv3(p[0], p[1], p[2]);
	if (cs->e2c_size <= (uint32_t)entity) {
		if (cs->e2c_cap <= (uint32_t)entity) {
			uint32_t cap = cs->e2c_cap ? cs->e2c_cap : 1024;
			while (cap <= (uint32_t)entity) cap *= 2;
			cs->entity_to_cell = (OSphere**)realloc(cs->entity_to_cell, sizeof(OSphere*) * cap);
			cs->e2c_cap = cap;
		}
		while (cs->e2c_size <= (uint32_t)entity) cs->entity_to_cell[cs->e2c_size++] = NULL;
	}
	const CellIndices i = cell_indices(pos, cs->cell_size, type, radius > cs->cell_size);
	MapSlot* s = map_find(cs, &i);
	if (!s) {
		CellPage* new_cell = page_alloc();
		/* :146 i.pos * double(m_cell_size) — math.cpp:149-152 IVec3*double = {i*x, i*y, i*z} */
		const double cs_d = (double)cs->cell_size;
		new_cell->header.origin = odv3(cs_d * i.pos.x, cs_d * i.pos.y, cs_d * i.pos.z);
		new_cell->header.indices = i;
		map_insert(cs, &i, new_cell);
		cells_push(cs, new_cell);
		s = map_find(cs, &i);
	}
	cs->entity_to_cell[entity] = add_to_cell(cs, s->value, entity, pos, radius);
}

/* culling_system.cpp:190-195 getCell */
static CellPage* get_cell(const OSphere* sphere) {
	const intptr_t ptr = (intptr_t)sphere;
	return (CellPage*)(ptr - (ptr % PAGE_SIZE));
}

/* culling_system.cpp:160-187 remove */
void oracle_culling_remove(OracleCulling* cs, int32_t entity) {
	if (cs->e2c_size <= (uint32_t)entity) return;
	const OSphere* sphere = cs->entity_to_cell[entity];
	if (!sphere) return;
	CellPage* cell = get_cell(sphere);
	if (cell->header.count == 1) {
		if (!cell->header.prev) {
			if (!cell->header.next) map_erase(cs, &cell->header.indices);
			else map_find(cs, &cell->header.indices)->value = cell->header.next;
		}
		if (cell->header.prev) cell->header.prev->header.next = cell->header.next;
		if (cell->header.next) cell->header.next->header.prev = cell->header.prev;
		cells_swap_and_pop(cs, cell);
		free(cell);
	}
	else {
		const int idx = (int)(sphere - cell->spheres);
		const int32_t last = cell->entities[cell->header.count - 1];
		cell->entities[idx] = cell->entities[cell->header.count - 1];
		cell->spheres[idx] = cell->spheres[cell->header.count - 1];
		cs->entity_to_cell[last] = &cell->spheres[idx];
		--cell->header.count;
	}
	cs->entity_to_cell[entity] = NULL;
}

static int iv3_eq(OIVec3 a, OIVec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

/* culling_system.cpp:198-214 setPosition */
void oracle_culling_set_position(OracleCulling* cs, int32_t entity, const double* p) {
	const ODVec3 pos = odv3(p[0], p[1], p[2]);
	OSphere* sphere = cs->entity_to_cell[entity];
	CellPage* cell = get_cell(sphere);
	const OIVec3 new_indices = oiv3_from_d(odv3_muls(pos, 1 / cs->cell_size));
	if (iv3_eq(new_indices, cell->header.indices.pos)) {
		sphere->position = ov3_from_d(odv3_sub(pos, cell->header.origin));
		return;
	}
	const float radius = sphere->radius;
	const uint8_t type = cell->header.indices.type;
	oracle_culling_remove(cs, entity);
	oracle_culling_add(cs, entity, type, p, radius);
}

/* culling_system.cpp:217-220 */
float oracle_culling_get_radius(const OracleCulling* cs, int32_t entity) { return cs->entity_to_cell[entity]->radius; }

/* culling_system.cpp:372-375 */
int oracle_culling_is_added(const OracleCulling* cs, int32_t entity) {
	return (uint32_t)entity < cs->e2c_size && cs->entity_to_cell[entity] != NULL;
}

/* culling_system.cpp:222-240 set */
void oracle_culling_set(OracleCulling* cs, int32_t entity, const double* p, float radius) {
	const ODVec3 pos = odv3(p[0], p[1], p[2]);
	OSphere* sphere = cs->entity_to_cell[entity];
	CellPage* cell = get_cell(sphere);
	const OIVec3 new_indices = oiv3_from_d(odv3_muls(pos, 1 / cs->cell_size));
	const int was_big = cell->header.indices.is_big;
	const int is_big = radius > cs->cell_size;
	if (was_big == is_big && iv3_eq(new_indices, cell->header.indices.pos)) {
		sphere->radius = radius;
		sphere->position = ov3_from_d(odv3_sub(pos, cell->header.origin));
		return;
	}
	const uint8_t type = cell->header.indices.type;
	oracle_culling_remove(cs, entity);
	oracle_culling_add(cs, entity, type, p, radius);
}

/* culling_system.cpp:242-258 setRadius */
void oracle_culling_set_radius(OracleCulling* cs, int32_t entity, float radius) {
	OSphere* sphere = cs->entity_to_cell[entity];
	CellPage* cell = get_cell(sphere);
	const int was_big = cell->header.indices.is_big;
	const int is_big = radius > cs->cell_size;
	if (was_big == is_big) {
		sphere->radius = radius;
		return;
	}
	const uint8_t type = cell->header.indices.type;
	/* :254 cell.header.origin + sphere->position (DVec3 + Vec3, math.cpp:512) */
	const ODVec3 pos = odv3_addf(cell->header.origin, sphere->position);
	const double p[3] = {pos.x, pos.y, pos.z};
	oracle_culling_remove(cs, entity);
	oracle_culling_add(cs, entity, type, p, radius);
}

/* f4MoveMask = sign bits (simd.h:119), NOT "< 0": -0.0f counts as culled, NaN by its sign bit */
static inline int sign_bit(float v) {
	uint32_t u;
	memcpy(&u, &v, 4);
	return (int)(u >> 31);
}

/* culling_system.cpp:260-308 doCulling; returns number appended */
static uint32_t do_culling(const CellPage* cell, const OFrustum* fr, uint32_t* out_ids, uint8_t* out_types, uint32_t cap, uint32_t cursor) {
	const uint32_t start = cursor;
	for (int i = 0; i < cell->header.count; ++i) {
		const OSphere* s = &cell->spheres[i];
		const float cx = s->position.x, cy = s->position.y, cz = s->position.z;
		/* :282 f4Splat(-sphere->radius) and :291 t - r are two separate SSE operations in the reference (xorps, subps).  They must stay
		 * separate here: folded into t + radius the value is the same, but a NaN radius would come out with the other sign, and the sign
		 * bit is what movemask reads (a +NaN radius culls the sphere on every plane, a -NaN radius makes it pass).  volatile keeps gcc from
		 * folding. */
		volatile float r = -s->radius;
		int culled = 0;
		for (int g = 0; g < 2 && !culled; ++g) {
			int mask = 0;
			for (int k = 0; k < 4; ++k) {
				const int p = g * 4 + k;
				/* :289 t = cx * px + cy * py + cz * pz + pd;  t = t - r */
				float t = cx * fr->xs[p] + cy * fr->ys[p] + cz * fr->zs[p] + fr->ds[p];
				t = t - r;
				mask |= sign_bit(t);
			}
			if (mask) culled = 1;
		}
		if (culled) continue;
		if (cursor < cap) {
			if (out_ids) out_ids[cursor] = (uint32_t)cell->entities[i];
			if (out_types) out_types[cursor] = cell->header.indices.type;
		}
		++cursor;
	}
	return cursor - start;
}

/* culling_system.cpp:321-369 cullInternal. type < 0 = all (0xff). Visits pages in m_cells order, slots in order:
 * a legal order (the reference's is nondeterministic). */
uint32_t oracle_culling_cull(const OracleCulling* cs, const OShiftedFrustum* frustum, int type, uint32_t* out_ids, uint8_t* out_types,
	uint32_t cap, OracleCullStats* stats)
{
	OracleCullStats st;
	memset(&st, 0, sizeof(st));
	uint32_t cursor = 0;
	const OVec3 v3_cell_size = ov3(cs->cell_size, cs->cell_size, cs->cell_size);
	const OVec3 v3_2_cell_size = ov3(2 * cs->cell_size, 2 * cs->cell_size, 2 * cs->cell_size);
	for (uint32_t ci = 0; ci < cs->cells_size; ++ci) {
		const CellPage* cell = cs->cells[ci];
		++st.pages_total;
		st.entities_total += (uint32_t)cell->header.count;
		if (type >= 0 && cell->header.indices.type != (uint8_t)type) { ++st.pages_filtered; continue; }
		int test_spheres = 0;
		if (cell->header.indices.is_big) {
			test_spheres = 1;
		}
		else if (oracle_frustum_contains_aabb(frustum, odv3_addf(cell->header.origin, v3_cell_size), v3_cell_size)) {
			/* :345-360 memcpy all ids */
			for (int i = 0; i < cell->header.count; ++i) {
				if (cursor < cap) {
					if (out_ids) out_ids[cursor] = (uint32_t)cell->entities[i];
					if (out_types) out_types[cursor] = cell->header.indices.type;
				}
				++cursor;
			}
			++st.pages_inside;
			st.entities_inside += (uint32_t)cell->header.count;
		}
		else if (oracle_frustum_intersects_aabb(frustum, odv3_subf(cell->header.origin, v3_cell_size), v3_2_cell_size)) {
			test_spheres = 1;
		}
		else {
			++st.pages_outside;
		}
		if (test_spheres) {
			OFrustum rel;
			oracle_frustum_get_relative(frustum, cell->header.origin, &rel);
			const uint32_t n = do_culling(cell, &rel, out_ids, out_types, cap, cursor);
			cursor += n;
			++st.pages_tested;
			st.entities_tested += (uint32_t)cell->header.count;
			st.visible_tested += n;
		}
	}
	st.visible = cursor;
	if (stats) *stats = st;
	return cursor;
}

uint32_t oracle_culling_page_count(const OracleCulling* cs) { return cs->cells_size; }

/* Dump page `idx` of m_cells for state comparison with the product's host bookkeeping. */
void oracle_culling_get_page(const OracleCulling* cs, uint32_t idx, double* origin3, int* indices3, uint8_t* type, uint8_t* is_big, int* count,
	float* spheres4, int32_t* entities)
{
	const CellPage* c = cs->cells[idx];
	origin3[0] = c->header.origin.x; origin3[1] = c->header.origin.y; origin3[2] = c->header.origin.z;
	indices3[0] = c->header.indices.pos.x; indices3[1] = c->header.indices.pos.y; indices3[2] = c->header.indices.pos.z;
	*type = c->header.indices.type;
	*is_big = c->header.indices.is_big;
	*count = c->header.count;
	if (spheres4) memcpy(spheres4, c->spheres, sizeof(OSphere) * (size_t)c->header.count);
	if (entities) memcpy(entities, c->entities, sizeof(int32_t) * (size_t)c->header.count);
}

/* batch helpers (ctypes call overhead) */
void oracle_culling_add_many(OracleCulling* cs, const int32_t* entities, const uint8_t* types, const double* pos3, const float* radius, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) oracle_culling_add(cs, entities[i], types[i], pos3 + 3 * i, radius[i]);
}
void oracle_culling_set_many(OracleCulling* cs, const int32_t* entities, const double* pos3, const float* radius, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) oracle_culling_set(cs, entities[i], pos3 + 3 * i, radius[i]);
}
void oracle_culling_set_position_many(OracleCulling* cs, const int32_t* entities, const double* pos3, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) oracle_culling_set_position(cs, entities[i], pos3 + 3 * i);
}
void oracle_culling_set_radius_many(OracleCulling* cs, const int32_t* entities, const float* radius, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) oracle_culling_set_radius(cs, entities[i], radius[i]);
}
void oracle_culling_remove_many(OracleCulling* cs, const int32_t* entities, uint32_t n) {
	for (uint32_t i = 0; i < n; ++i) oracle_culling_remove(cs, entities[i]);
}

/* math.cpp:1333-1341, 1372-1378 Marsaglia RNG, for seeded scenes identical to reference-side harnesses */
void oracle_rng_floats(uint32_t* u_, uint32_t* v_, uint32_t n, float* out) {
	uint32_t u = *u_, v = *v_;
	for (uint32_t i = 0; i < n; ++i) {
		u = 36969 * (u & 65535) + (u >> 16);
		v = 18000 * (v & 65535) + (v >> 16);
		const uint32_t r = (u << 16) + v;
		out[i] = (float)(r * 2.328306435996595e-10);
	}
	*u_ = u; *v_ = v;
}
