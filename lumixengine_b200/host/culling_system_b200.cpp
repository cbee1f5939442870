// Engine-side shim: drops the GPU cull into LumixEngine by replacing the body of CullingSystem::create
// (src/renderer/culling_system.cpp:399-402).  Written against the ENGINE'S OWN headers — it is compiled inside the engine tree
// (add this file to src/renderer, remove culling_system.cpp's CullingSystemImpl + create, link liblumix_b200.so), not as part of
// liblumix_b200.so.  tests/test_integration_compile.py syntax-checks it against /root/reference where that tree is present.
//
// Everything the engine sees is unchanged: the CullingSystem vtable (culling_system.h:58-77), results as a CullResult page chain
// allocated from the engine's PageAllocator (one renderable type per 4 KB page, <= 1020 ids, culling_system.h:17-56) and freed by the
// caller with CullResult::free (culling_system.cpp:388-396; pipeline.cpp:1045,3411).
#include "engine/lumix.h"

#include "core/allocator.h"
#include "core/crt.h"
#include "core/geometry.h"
#include "core/job_system.h"
#include "core/log.h"
#include "core/math.h"
#include "core/sync.h"
#include "core/page_allocator.h"
#include "core/string.h"
#include "renderer/culling_system.h"

#include "lumix_b200.h"

#include <string.h>

namespace Lumix {

static_assert(sizeof(ShiftedFrustum) == sizeof(lb200_shifted_frustum), "lb200_shifted_frustum is a byte image of ShiftedFrustum");
static_assert(sizeof(CullResult) == PageAllocator::PAGE_SIZE);

// One lb200_ctx per process, created on first use (the ISystem of INTEGRATION.md §0 owns it in a full integration).
static lb200_ctx* getB200Context() {
	static lb200_ctx* ctx = [] {
		lb200_ctx* c = nullptr;
		if (lb200_init(0, &c) != LB200_OK) logError("lumix_b200: ", lb200_last_error(nullptr));
		return c;
	}();
	return ctx;
}

struct CullingSystemB200 final : CullingSystem {
	CullingSystemB200(IAllocator& allocator, PageAllocator& page_allocator)
		: m_allocator(allocator)
		, m_page_allocator(page_allocator)
	{
		m_ctx = getB200Context();
		lb200_culling_create(m_ctx, &m_cs); // with m_ctx == nullptr this is host bookkeeping only and cull() reports the error
	}

	~CullingSystemB200() override {
		if (m_ids) lb200_host_free(m_ctx, m_ids);
		lb200_culling_destroy(m_cs);
	}

	// culling_system.cpp:131-258 — one C call each
	void add(EntityRef entity, u8 type, const DVec3& pos, float radius) override { lb200_culling_add(m_cs, entity.index, type, &pos.x, radius); }
	void remove(EntityRef entity) override { lb200_culling_remove(m_cs, entity.index); }
	void setPosition(EntityRef entity, const DVec3& pos) override { lb200_culling_set_position(m_cs, entity.index, &pos.x); }
	void setRadius(EntityRef entity, float radius) override { lb200_culling_set_radius(m_cs, entity.index, radius); }
	void set(EntityRef entity, const DVec3& pos, float radius) override { lb200_culling_set(m_cs, entity.index, &pos.x, radius); }
	float getRadius(EntityRef entity) override { return lb200_culling_get_radius(m_cs, entity.index); }
	bool isAdded(EntityRef entity) override { return lb200_culling_is_added(m_cs, entity.index) != 0; }

	CullResult* cull(const ShiftedFrustum& frustum, u8 type) override {
		ASSERT(type != 0xff); // 0xff type is reserved for `all types`, culling_system.cpp:312
		return cullInternal(frustum, type);
	}

	CullResult* cull(const ShiftedFrustum& frustum) override { return cullInternal(frustum, 0xff); }

	bool ensureCapacity(u32 ids) {
		if (ids <= m_capacity) return true;
		u32 cap = m_capacity ? m_capacity : 4096;
		while (cap < ids) cap *= 2;
		if (m_ids) lb200_host_free(m_ctx, m_ids);
		m_ids = (u32*)lb200_host_alloc(m_ctx, sizeof(u32) * size_t(cap)); // page-locked: the D2H copy of the ids runs at full PCIe speed
		m_capacity = m_ids ? cap : 0;
		return m_ids != nullptr;
	}

	CullResult* cullInternal(const ShiftedFrustum& frustum, u8 type) {
		// cull is called from job-system fibers, possibly for several views at once (pipeline.cpp:1036-1041): one GPU cull at a time
		jobs::MutexGuard guard(m_mutex);
		if (!ensureCapacity(lb200_culling_entity_count(m_cs))) return nullptr;
		lb200_cull_result res;
		int rc;
		if (lb200_culling_page_count(m_cs) == 0) return nullptr; // culling_system.cpp:322
		// No OS wait inside a job (docs/job_system.md) and no spinning on jobs::yield() (its own TODO: a yielding fiber can be popped while it
		// is still switching out, job_system.cpp:763): enqueue the cull, let the driver call back when the stream has reached the end of it,
		// and park this fiber on a jobs::Signal meanwhile.  The callback runs on a driver thread, which is not a job worker, so it only
		// schedules a job (jobs::run is legal from any thread) and that job turns the signal green.
		jobs::Signal done;
		jobs::turnRed(&done);
		rc = lb200_culling_cull_begin(m_cs, (const lb200_shifted_frustum*)&frustum, type, m_ids, m_capacity);
		if (rc == LB200_OK) {
			rc = lb200_host_callback(m_ctx, [](void* signal) { jobs::run(signal, [](void* s) { jobs::turnGreen((jobs::Signal*)s); }, nullptr); }, &done);
			if (rc == LB200_OK) jobs::wait(&done);
			else while (lb200_culling_cull_poll(m_cs) == 0) {} // could not register the callback: the cull itself is short
			rc = lb200_culling_cull_end(m_cs, &res);
		}
		if (rc != LB200_OK) { // no CPU fallback: report and return "nothing visible" (the reference's own empty result, :322)
			logError("lumix_b200 cull failed (code ", rc, "): ", lb200_last_error(m_ctx));
			return nullptr;
		}
		CullResult* head = nullptr;
		CullResult** link = &head;
		for (u32 t = 0; t < res.n_types; ++t) {
			const u32* src = m_ids + res.type_offset[t];
			u32 left = res.type_count[t];
			while (left) {
				CullResult* page = new (NewPlaceholder(), m_page_allocator.allocate()) CullResult;
				const u32 n = minimum(left, (u32)lengthOf(page->entities));
				memcpy(page->entities, src, n * sizeof(EntityRef));
				page->header.count = n;
				page->header.type = (u8)t;
				page->header.next = nullptr;
				*link = page;
				link = &page->header.next;
				src += n;
				left -= n;
			}
		}
		return head; // the caller frees it with CullResult::free(PageAllocator&)
	}

	IAllocator& m_allocator;
	PageAllocator& m_page_allocator;
	lb200_ctx* m_ctx = nullptr;
	lb200_culling* m_cs = nullptr;
	u32* m_ids = nullptr;
	u32 m_capacity = 0;
	jobs::Mutex m_mutex;
};

void CullResult::free(PageAllocator& allocator) { // culling_system.cpp:388-396
	CullResult* i = this;
	while (i) {
		CullResult* tmp = i;
		i = i->header.next;
		allocator.deallocate(tmp);
	}
}

UniquePtr<CullingSystem> CullingSystem::create(IAllocator& allocator, PageAllocator& page_allocator) {
	return UniquePtr<CullingSystemB200>::create(allocator, allocator, page_allocator);
}

} // namespace Lumix
