	// --- lumix_b200: added inside `struct World`, next to transformEntity (world.h:139-140);
	// world.h also gets `struct lb200_ctx;` in front of `namespace Lumix` (the C handle lives in the global namespace) ---
public:
	void setTransformsDeferredB200(const EntityRef* entities, const Transform* transforms, u32 count);
	bool propagateHierarchyB200(::lb200_ctx* ctx, bool topology_changed);
	void destroyHierarchyB200(); // from ~World
private:
	struct HierarchyB200;
	HierarchyB200* m_hierarchy_b200 = nullptr;
