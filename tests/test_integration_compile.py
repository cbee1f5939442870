"""The engine-side bindings under lumixengine_b200/host/ compile against the reference's own sources: the CullingSystem shim
(same vtable, same types, same ownership calls) and the World patch (appended to a temporary copy of world.cpp, as INTEGRATION.md
says a maintainer would).  Only where /root/reference is present (this container)."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("LUMIX_REFERENCE", "/root/reference")


def _copy_headers(tmp, subs):
    for sub in subs:
        dst = os.path.join(tmp, "src", sub)
        os.makedirs(dst)
        for dirpath, _, files in os.walk(os.path.join(REF, "src", sub)):
            rel = os.path.relpath(dirpath, os.path.join(REF, "src", sub))
            for f in files:
                if f.endswith((".h", ".inl")):
                    os.makedirs(os.path.join(dst, rel), exist_ok=True)
                    shutil.copy(os.path.join(dirpath, f), os.path.join(dst, rel, f))
    sp = os.path.join(tmp, "src", "core", "sync.h")
    text = open(sp).read().replace('#error "Not implemented"', "pthread_rwlock_t lock;")
    open(sp, "w").write(text)


GXX = ["/usr/bin/g++", "-std=c++20", "-DSTATIC_PLUGINS", "-DNDEBUG", "-fno-exceptions", "-fno-rtti", "-msse2", "-w", "-c"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "engine")), reason="reference tree not present")
def test_world_patch_compiles_inside_world_cpp():
    """host/world_b200_decl.inl goes into `struct World`, host/world_b200.inl is appended to world.cpp: the result must compile with
    the reference's own flags and define World::propagateHierarchyB200 next to the untouched World::transformEntity."""
    host = os.path.join(ROOT, "lumixengine_b200", "host")
    with tempfile.TemporaryDirectory() as tmp:
        _copy_headers(tmp, ("core", "engine"))
        wh = os.path.join(tmp, "src", "engine", "world.h")
        text = open(wh).read()
        anchor = "private:\n\tvoid transformEntity(EntityRef entity, bool update_local);"
        assert anchor in text, "world.h changed: INTEGRATION.md section 2 needs another anchor"
        text = text.replace(anchor, open(os.path.join(host, "world_b200_decl.inl")).read() + anchor)
        assert text.count("namespace Lumix {") >= 1
        text = text.replace("namespace Lumix {", "struct lb200_ctx; // include/lumix_b200.h\nnamespace Lumix {", 1)
        open(wh, "w").write(text)
        wc = os.path.join(tmp, "src", "engine", "world.cpp")
        body = open(os.path.join(REF, "src", "engine", "world.cpp")).read() + "\n" + open(os.path.join(host, "world_b200.inl")).read()
        open(wc, "w").write(body)
        obj = os.path.join(tmp, "world.o")
        cmd = GXX + ["-I", os.path.join(tmp, "src"), "-I", os.path.join(REF, "external"), "-I", os.path.join(ROOT, "include"), wc, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
        assert "Lumix::World::propagateHierarchyB200(lb200_ctx*, bool)" in syms
        assert "Lumix::World::setTransformsDeferredB200(Lumix::EntityRef const*, Lumix::Transform const*, unsigned int)" in syms
        assert "Lumix::World::transformEntity(Lumix::EntityRef, bool)" in syms
        for f in ("lb200_hierarchy_create", "lb200_hierarchy_set_locals", "lb200_hierarchy_propagate", "lb200_hierarchy_get_globals"):
            assert f in syms  # unresolved here, provided by liblumix_b200.so


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "engine")), reason="reference tree not present")
def test_plugin_entry_compiles_against_engine_headers():
    """host/b200_system.cpp: the ISystem + LUMIX_PLUGIN_ENTRY the engine's SystemManager loads (plugin.h:64-96), static-plugin flavour."""
    with tempfile.TemporaryDirectory() as tmp:
        _copy_headers(tmp, ("core", "engine"))
        obj = os.path.join(tmp, "b200_system.o")
        cmd = GXX + ["-I", os.path.join(tmp, "src"), "-I", os.path.join(REF, "external"), "-I", os.path.join(ROOT, "include"),
                     os.path.join(ROOT, "lumixengine_b200", "host", "b200_system.cpp"), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
        assert "createPlugin_b200" in syms  # -DSTATIC_PLUGINS: the name plugins.inl references
        for f in ("lb200_init", "lb200_shutdown", "lb200_synchronize"):
            assert f in syms


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "renderer")), reason="reference tree not present")
def test_shim_compiles_against_engine_headers():
    # the Linux port of the reference lacks an SRWLock body (src/core/sync.h:22-24): give the preprocessor the one-line
    # overlay build_ref.sh uses, through a shadow include dir that only holds the patched sync.h
    with tempfile.TemporaryDirectory() as tmp:
        # headers only, copied to a temp dir so that sync.h can carry the overlay (never into this repository)
        for sub in ("core", "engine", "renderer"):
            dst = os.path.join(tmp, "src", sub)
            os.makedirs(dst)
            for dirpath, _, files in os.walk(os.path.join(REF, "src", sub)):
                rel = os.path.relpath(dirpath, os.path.join(REF, "src", sub))
                for f in files:
                    if f.endswith((".h", ".inl")):
                        os.makedirs(os.path.join(dst, rel), exist_ok=True)
                        shutil.copy(os.path.join(dirpath, f), os.path.join(dst, rel, f))
        sp = os.path.join(tmp, "src", "core", "sync.h")
        text = open(sp).read().replace('#error "Not implemented"', "pthread_rwlock_t lock;")
        open(sp, "w").write(text)
        obj = os.path.join(tmp, "shim.o")
        cmd = ["/usr/bin/g++", "-std=c++20", "-DSTATIC_PLUGINS", "-DNDEBUG", "-fno-exceptions", "-fno-rtti", "-msse2", "-w", "-c",
               "-I", os.path.join(tmp, "src"), "-I", os.path.join(REF, "external"), "-I", os.path.join(ROOT, "include"),
               os.path.join(ROOT, "lumixengine_b200", "host", "culling_system_b200.cpp"), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
        assert "Lumix::CullingSystem::create(Lumix::IAllocator&, Lumix::PageAllocator&)" in syms
        assert "lb200_culling_cull" in syms  # unresolved here, provided by liblumix_b200.so


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "animation")), reason="reference tree not present")
def test_animation_binding_compiles_against_render_module():
    """host/animation_b200.inl instantiated with the engine's own RenderModule (render_module.h:402-403 lockPose / unlockPose,
    getModelInstanceModel), animation.h patched with host/animation_b200_decl.inl as INTEGRATION.md section 3 says."""
    host = os.path.join(ROOT, "lumixengine_b200", "host")
    with tempfile.TemporaryDirectory() as tmp:
        _copy_headers(tmp, ("core", "engine", "renderer", "animation"))
        ah = os.path.join(tmp, "src", "animation", "animation.h")
        text = open(ah).read()
        anchor = "\tconst Array<TranslationTrack>& getTranslations() const { return m_translations; }"
        assert anchor in text, "animation.h changed: INTEGRATION.md section 3 needs another anchor"
        open(ah, "w").write(text.replace(anchor, open(os.path.join(host, "animation_b200_decl.inl")).read() + anchor))
        tu = os.path.join(tmp, "tu.cpp")
        open(tu, "w").write('#include "animation/animation.h"\n#include "animation/animation_module.h"\n#include "core/log.h"\n#include "renderer/model.h"\n'
                            '#include "renderer/pose.h"\n#include "renderer/render_module.h"\n#include <string.h>\n#include "animation_b200.inl"\n'
                            'namespace Lumix { template struct AnimablesB200<RenderModule>; }\n')
        obj = os.path.join(tmp, "tu.o")
        cmd = GXX + ["-I", os.path.join(tmp, "src"), "-I", os.path.join(REF, "external"), "-I", os.path.join(ROOT, "include"), "-I", host, tu, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
        assert "Lumix::AnimablesB200<Lumix::RenderModule>::update(lb200_ctx*, Lumix::RenderModule&, Lumix::Span<Lumix::Animable>, float)" in syms
        for f in ("lb200_animation_create", "lb200_animation_set_instances", "lb200_animation_update", "lb200_animation_get_pose", "lb200_animation_get_times"):
            assert f in syms
