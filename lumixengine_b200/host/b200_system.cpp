// Engine-side plugin entry of liblumix_b200: an ISystem that owns the lb200_ctx for the process.
//
// LUMIX_PLUGIN_ENTRY(b200) is what SystemManager::load looks up — createPlugin via os::getLibrarySymbol for a dynamic plugin,
// createPlugin_b200 from the generated plugins.inl for a static build (src/engine/plugin.h:92-96, plugin.cpp:122-170, 217-223).
// The object is allocated with LUMIX_NEW(engine.getAllocator(), ...) because the engine destroys it with
// LUMIX_DELETE(engine.getAllocator(), system) (plugin.cpp:34-40).  All GPU state is derived from World every frame, so there is
// nothing to serialize.  tests/test_integration_compile.py compiles this file against the reference's headers.
#include "engine/engine.h"
#include "engine/plugin.h"
#include "core/allocator.h"
#include "core/log.h"
#include "core/stream.h"
#include "core/string.h"

#include "lumix_b200.h"

namespace Lumix {

struct B200System final : ISystem {
	explicit B200System(Engine& engine) : m_engine(engine) {
		// no CPU fallback: without a device the context stays null, CullingSystemB200::cull then logs and returns nullptr
		if (lb200_init(0, &m_ctx) != LB200_OK) logError("lumix_b200: ", lb200_last_error(nullptr));
	}
	~B200System() override { lb200_shutdown(m_ctx); }

	const char* getName() const override { return "b200"; }
	void serialize(OutputMemoryStream&) const override {}
	bool deserialize(i32, InputMemoryStream&) override { return true; }
	void shutdownStarted() override { if (m_ctx) lb200_synchronize(m_ctx); } // other systems still exist: let queued work drain

	lb200_ctx* context() const { return m_ctx; }

	Engine& m_engine;
	lb200_ctx* m_ctx = nullptr;
};

// the accessor the other shims use (culling_system_b200.cpp keeps a process-wide fallback for builds without this system)
lb200_ctx* getB200Context(ISystem& system) { return static_cast<B200System&>(system).context(); }

LUMIX_PLUGIN_ENTRY(b200) {
	return LUMIX_NEW(engine.getAllocator(), B200System)(engine);
}

} // namespace Lumix
