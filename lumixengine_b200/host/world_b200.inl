// World::propagateHierarchyB200 — the engine-side binding of lb200_hierarchy_* (INTEGRATION.md §2).
//
// This file is the patch a LumixEngine maintainer adds: it is appended to src/engine/world.cpp (it needs World's private hierarchy
// arrays, world.h:157-164,190), together with the two declarations of world_b200_decl.inl inside `struct World`.
// tests/test_integration_compile.py applies exactly that to a temporary copy of the reference's world.h / world.cpp and compiles it.
//
// What it replaces: the recursion World::transformEntity (world.cpp:255-282) for frames in which many hierarchy entities moved —
// one batched level-order pass on the GPU instead of one DFS + delegate storm per moved root.

#include "lumix_b200.h"

namespace Lumix {

static_assert(sizeof(Transform) == sizeof(lb200_transform), "Transform is passed to the library as is (math.h:306-327)");

struct World::HierarchyB200 {
	HierarchyB200(IAllocator& allocator) : parents(allocator), locals(allocator), globals(allocator) {}
	~HierarchyB200() { lb200_hierarchy_destroy(handle); }
	lb200_hierarchy* handle = nullptr;
	Array<i32> parents;         // node i = m_hierarchy[i]; parent node or -1
	Array<Transform> locals;    // Hierarchy::local_transform per node
	Array<Transform> globals;   // m_transforms of the node's entity, in and out
	u32 built_for = 0xffFFffFF; // m_hierarchy.size() the topology was built for
};

// Call after any batch of setLocalTransform / root moves, instead of letting every one of them recurse.
// `topology_changed`: pass true after setParent / entity destruction (world.cpp:619-701 rewires first_child / next_sibling).
bool World::propagateHierarchyB200(lb200_ctx* ctx, bool topology_changed) {
	if (m_hierarchy.empty()) return true;
	if (!m_hierarchy_b200) m_hierarchy_b200 = LUMIX_NEW(m_allocator, HierarchyB200)(m_allocator);
	HierarchyB200& h = *m_hierarchy_b200;
	const u32 n = (u32)m_hierarchy.size();
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
			return false; // no CPU fallback here: the caller keeps using transformEntity
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
	for (u32 i = 0; i < n; ++i) {
		const EntityRef e = m_hierarchy[i].entity;
		if (!m_hierarchy[i].parent.isValid()) continue; // roots were inputs
		m_transforms[e.index] = h.globals[i];
		// what transformEntity does per entity (world.cpp:257-260): tell the modules that own a component of this entity
		const ArchetypeManager::Archetype& archetype = m_archetype_manager->get(m_entities[e.index].archetype);
		for (ComponentType type : archetype.types) m_component_type_map[type.index]->transformed.invoke(e);
	}
	return true;
}

void World::destroyHierarchyB200() {
	LUMIX_DELETE(m_allocator, m_hierarchy_b200);
	m_hierarchy_b200 = nullptr;
}

} // namespace Lumix
