// World::setTransformsDeferredB200 / World::propagateHierarchyB200 — the engine-side binding of lb200_hierarchy_* (INTEGRATION.md §2).
//
// This file is the patch a LumixEngine maintainer adds: it is appended to src/engine/world.cpp (it needs World's private hierarchy
// arrays, world.h:157-164,190), together with the declarations of world_b200_decl.inl inside `struct World` and one
// `destroyHierarchyB200();` line in World::~World.  tests/test_integration_compile.py applies exactly that to a temporary copy of the
// reference's world.h / world.cpp and compiles it; oracle/build_ref.sh links the result with the reference's own core objects and
// oracle/ref/ref_world_shim_harness.cpp runs it next to the unpatched recursion (tests/test_engine_boundary_gpu.py).
//
// What it replaces: the recursion World::transformEntity (world.cpp:255-282) for frames in which many hierarchy roots moved —
// one batched level-order pass on the GPU instead of one DFS per moved root.  The set of entities whose `transformed` delegates fire is
// the reference's (every moved entity and all its descendants); the order is moved entities first, then hierarchy nodes in m_hierarchy
// order, not DFS order.

#include "lumix_b200.h"

namespace Lumix {

static_assert(sizeof(Transform) == sizeof(lb200_transform), "Transform is passed to the library as is (math.h:306-327)");

struct World::HierarchyB200 {
	HierarchyB200(IAllocator& allocator) : parents(allocator), locals(allocator), globals(allocator), moved(allocator), state(allocator) {}
	~HierarchyB200() { lb200_hierarchy_destroy(handle); }
	lb200_hierarchy* handle = nullptr;
	Array<i32> parents;         // node i = m_hierarchy[i]; parent node or -1
	Array<Transform> locals;    // Hierarchy::local_transform per node
	Array<Transform> globals;   // m_transforms of the node's entity, in and out
	Array<EntityRef> moved;     // entities set through setTransformsDeferredB200 since the last propagate
	Array<u8> state;            // per node: 0 unknown, 1 under a moved entity, 2 not
	u32 built_for = 0xffFFffFF; // m_hierarchy.size() the topology was built for
};

// World::setTransform (world.cpp:337-342) for a batch, without the recursion: the new transforms are stored, the children follow at
// the next propagateHierarchyB200.  An entity that has a parent takes the reference path — its local transform has to be recomputed
// from the new global one (world.cpp:267-270) before anything below it moves.
void World::setTransformsDeferredB200(const EntityRef* entities, const Transform* transforms, u32 count) {
	if (!m_hierarchy_b200) m_hierarchy_b200 = LUMIX_NEW(m_allocator, HierarchyB200)(m_allocator);
	HierarchyB200& h = *m_hierarchy_b200;
	for (u32 i = 0; i < count; ++i) {
		const EntityRef e = entities[i];
		const i32 hi = m_entities[e.index].hierarchy;
		if (hi >= 0 && m_hierarchy[hi].parent.isValid()) {
			setTransform(e, transforms[i]);
			continue;
		}
		m_transforms[e.index] = transforms[i];
		h.moved.push(e);
	}
}

// Call once per frame after the batch of moves.  `topology_changed`: pass true after setParent / entity destruction
// (world.cpp:619-701 rewires first_child / next_sibling and may swap m_hierarchy slots).
bool World::propagateHierarchyB200(lb200_ctx* ctx, bool topology_changed) {
	if (!m_hierarchy_b200) m_hierarchy_b200 = LUMIX_NEW(m_allocator, HierarchyB200)(m_allocator);
	HierarchyB200& h = *m_hierarchy_b200;
	const u32 n = (u32)m_hierarchy.size();
	auto fire = [this](EntityRef e) { // what transformEntity does first (world.cpp:257-260): tell the modules that own a component of this entity
		const ArchetypeManager::Archetype& archetype = m_archetype_manager->get(m_entities[e.index].archetype);
		for (ComponentType type : archetype.types) m_component_type_map[type.index]->transformed.invoke(e);
	};
	if (n == 0) {
		for (EntityRef e : h.moved) fire(e);
		h.moved.clear();
		return true;
	}
	if (topology_changed || h.built_for != n || !h.handle) {
		h.parents.resize(n);
		for (u32 i = 0; i < n; ++i) {
			const EntityPtr parent = m_hierarchy[i].parent;
			h.parents[i] = parent.isValid() ? m_entities[parent.index].hierarchy : -1;
		}
		lb200_hierarchy_destroy(h.handle);
		h.handle = nullptr;
		if (lb200_hierarchy_create(ctx, h.parents.begin(), n, &h.handle) != LB200_OK) {
			logError("lumix_b200 hierarchy: ", lb200_last_error(ctx));
			return false; // no CPU fallback here: the deferred moves stay recorded, the caller may retry or transformEntity them
		}
		h.built_for = n;
	}
	h.locals.resize(n);
	h.globals.resize(n);
	for (u32 i = 0; i < n; ++i) {
		h.locals[i] = m_hierarchy[i].local_transform;
		h.globals[i] = m_transforms[m_hierarchy[i].entity.index]; // only the roots' entries are read by the library
	}
	if (lb200_hierarchy_set_locals(h.handle, (const lb200_transform*)h.locals.begin()) != LB200_OK
		|| lb200_hierarchy_set_root_globals(h.handle, (const lb200_transform*)h.globals.begin()) != LB200_OK
		|| lb200_hierarchy_propagate(h.handle) != LB200_OK
		|| lb200_hierarchy_get_globals(h.handle, (lb200_transform*)h.globals.begin()) != LB200_OK)
	{
		logError("lumix_b200 hierarchy: ", lb200_last_error(ctx));
		return false;
	}
	// which nodes sit under a moved entity: roots from the list, everyone else inherits from the parent (chains resolved once, then cached)
	h.state.resize(n);
	for (u32 i = 0; i < n; ++i) h.state[i] = h.parents[i] < 0 ? 2 : 0;
	for (EntityRef e : h.moved) {
		const i32 hi = m_entities[e.index].hierarchy;
		if (hi >= 0) h.state[hi] = 1;
	}
	for (EntityRef e : h.moved) fire(e);
	for (u32 i = 0; i < n; ++i) {
		if (h.parents[i] < 0) continue;
		u32 top = i;
		while (h.state[top] == 0) top = (u32)h.parents[top];
		const u8 s = h.state[top];
		for (u32 k = i; h.state[k] == 0; k = (u32)h.parents[k]) h.state[k] = s;
		if (s != 1) continue;
		const EntityRef e = m_hierarchy[i].entity;
		m_transforms[e.index] = h.globals[i]; // world.cpp:275-277
		fire(e);
	}
	h.moved.clear();
	return true;
}

void World::destroyHierarchyB200() {
	LUMIX_DELETE(m_allocator, m_hierarchy_b200);
	m_hierarchy_b200 = nullptr;
}

} // namespace Lumix
