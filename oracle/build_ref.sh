#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds oracle/_ref/libref_lumix.so from the reference's OWN sources where they
# lie under /root/reference (never copied into this repo; the patched overlay copy lives in a mktemp
# dir that is deleted afterwards).  Does not run the reference's build system (GENie/MSBuild).
#
#   unmodified : src/core/{math,geometry,string,hash,log,stream,page_allocator,default_allocator,
#                arena_allocator,path}.cpp, src/core/linux/{thread,atomic,fibers}.cpp, src/engine/resource.cpp,
#                src/renderer/{culling_system,pose}.cpp, src/animation/animation.cpp
#   overlay    : SURVEY.md §8(c) — src/core/sync.h (SRWLock body), src/core/linux/sync.cpp (stale
#                Semaphore::signal signature + SRWLock methods), src/core/simd.h (take the SSE branch
#                on GCC), src/core/job_system.cpp (noinline on getWorker(): the reference's TLS guard
#                is MSVC/clang pragmas only; 1 MB fiber stacks as on Windows)
#   stubs      : oracle/ref/ref_stubs.cpp (os::mem*, profiler no-ops, atomics missing on Linux)
#
# Flags mirror the reference's Linux config (scripts/genie.lua:339-342: -msse2, no FMA, no fast-math)
# plus -msse3 for _mm_hadd_ps (simd_math.h:111-119) and -ffp-contract=off to make "no FMA" explicit.
set -euo pipefail
REF=${LUMIX_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/core" ]; then
	echo "build_ref.sh: $REF not present — keeping prebuilt oracle/_ref as is" >&2
	exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/lumix_ref_overlay.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
S="$TMP/src"
mkdir -p "$S/core" "$S/engine" "$S/renderer" "$S/animation" "$TMP/obj"
cp -r "$REF/src/core/." "$S/core/"
cp "$REF"/src/engine/*.h "$S/engine/"
cp "$REF"/src/renderer/*.h "$S/renderer/"
cp -r "$REF/src/renderer/gpu" "$S/renderer/"
cp "$REF/src/renderer/culling_system.cpp" "$REF/src/renderer/pose.cpp" "$S/renderer/"
cp "$REF"/src/animation/*.h "$S/animation/"
cp "$REF/src/animation/animation.cpp" "$S/animation/"
cp "$REF/src/engine/resource.cpp" "$S/engine/"

# (1) sync.h: SRWLock has no Linux body
sed -i 's|#error "Not implemented"|pthread_rwlock_t lock;|' "$S/core/sync.h"
# (2) linux/sync.cpp: stale signature + missing SRWLock methods
sed -i 's|^void Semaphore::signal()|void Semaphore::signal(u32)|' "$S/core/linux/sync.cpp"
cat >> "$S/core/linux/sync.cpp" <<'EOF'
namespace Lumix {
SRWLock::SRWLock() { pthread_rwlock_init(&lock, nullptr); }
SRWLock::~SRWLock() { pthread_rwlock_destroy(&lock); }
void SRWLock::enterExclusive() { pthread_rwlock_wrlock(&lock); }
void SRWLock::exitExclusive() { pthread_rwlock_unlock(&lock); }
void SRWLock::enterShared() { pthread_rwlock_rdlock(&lock); }
void SRWLock::exitShared() { pthread_rwlock_unlock(&lock); }
}
EOF
# (3) simd.h: use the SSE (MSVC) branch on GCC; GCC's __m128 already has + - * built in, so the
#     overloads of that branch must go (unary minus then differs from the reference's 0-a only in
#     the sign of zero)
python3 - "$S/core/simd.h" <<'EOF'
import sys
p = sys.argv[1]
L = open(p).read().split('\n')
out = []; i = 0; seen_else = False; in_sse = False
while i < len(L):
    line = L[i]
    if 'using float4 = __m128;' in line: in_sse = True
    if in_sse and line.strip().startswith('#else'): seen_else = True
    if line.startswith('#if defined _WIN32 && !defined __clang__'):
        out.append('#if 1'); i += 1; continue
    if '#include <intrin.h>' in line:
        out.append('\t#include <immintrin.h>\n\t#include <math.h>\n\t#include <string.h>'); i += 1; continue
    if not seen_else and 'LUMIX_FORCE_INLINE float4 operator' in line:
        i += 3
        if i < len(L) and L[i].strip() == '': i += 1
        continue
    out.append(line); i += 1
open(p, 'w').write('\n'.join(out))
EOF
# (4) job_system.cpp: keep getWorker() out of line so &g_worker is recomputed after a fiber switch
sed -i 's|^WorkerTask\* getWorker()|__attribute__((noinline)) WorkerTask* getWorker()|' "$S/core/job_system.cpp"
#     and give fibers 1 MB of stack: Windows fibers reserve 1 MB whatever CreateFiber is asked to commit, the Linux port mallocs exactly the
#     64 KB it is asked for, which PipelineImpl::radixSort's two 48 KB histograms overflow (ref_sortkeys_harness.cpp runs it in a job)
sed -i 's|Fiber::create(64 \* 1024, manage, new_fiber)|Fiber::create(1024 * 1024, manage, new_fiber)|' "$S/core/job_system.cpp"

CXX=${LUMIX_REF_CXX:-/usr/bin/g++} # not $CXX: a wrapper there may link libstdc++ statically, which clashes with other C++ libs at exit
FL="-std=c++20 -O2 -DSTATIC_PLUGINS -DNDEBUG -fno-exceptions -fno-rtti -msse2 -msse3 -ffp-contract=off -fPIC -w -fvisibility=hidden -I$S -I$REF/external"
SRCS="renderer/culling_system core/job_system core/page_allocator core/default_allocator core/arena_allocator
      core/linux/thread core/linux/sync core/linux/atomic core/linux/fibers core/math core/geometry core/string
      core/hash core/log core/stream core/path engine/resource renderer/pose animation/animation"
OBJS=""
for f in $SRCS; do
	o="$TMP/obj/$(echo "$f" | tr / _).o"
	$CXX $FL -c "$S/$f.cpp" -o "$o" &
	OBJS="$OBJS $o"
done
$CXX $FL -c "$HERE/ref/ref_stubs.cpp" -o "$TMP/obj/ref_stubs.o" &
$CXX $FL -c "$HERE/ref/ref_harness.cpp" -o "$TMP/obj/ref_harness.o" &
$CXX $FL -c "$HERE/ref/ref_anim_harness.cpp" -o "$TMP/obj/ref_anim_harness.o" &
wait
# (5) PipelineImpl::computeSkeletonDualQuats lives in the DX12-bound pipeline.cpp: cut that one member function out of the reference
#     file into the overlay (never into this repository) and compile it inside oracle/ref/ref_palette_harness.cpp.  Optional: if the
#     reference moves the function, the library is built without ref_skeleton_dual_quats and the tests that need it skip.
EXTRA=""
awk '/void computeSkeletonDualQuats\(const ModelInstance\* mi\) \{/{p=1} /void createCommands\(View& view\)/{p=0} p' \
	"$REF/src/renderer/pipeline.cpp" > "$S/renderer/extracted_compute_skeleton_dual_quats.inl"
# the two file-static helpers of model.cpp (model.cpp itself needs the whole renderer to link): evaluateSkin :103-109, computeSkinMatrices :132-137
awk '/^static Vec3 evaluateSkin\(/{p=1} /^Vec3 Model::evalVertexPose\(/{p=0} p' "$REF/src/renderer/model.cpp" > "$S/renderer/extracted_model_statics.inl"
awk '/^static void computeSkinMatrices\(/{p=1} /^RayCastModelHit Model::castRay\(/{p=0} p' "$REF/src/renderer/model.cpp" >> "$S/renderer/extracted_model_statics.inl"
if [ -s "$S/renderer/extracted_compute_skeleton_dual_quats.inl" ] && grep -q computeSkinMatrices "$S/renderer/extracted_model_statics.inl" \
	&& $CXX $FL -c "$HERE/ref/ref_palette_harness.cpp" -o "$TMP/obj/ref_palette_harness.o" 2> "$TMP/palette.log"; then
	EXTRA="$TMP/obj/ref_palette_harness.o"
else
	echo "build_ref.sh: palette harness not built (see below); continuing without ref_skeleton_dual_quats" >&2
	tail -5 "$TMP/palette.log" >&2 || true
fi
# (6) the sort-key packers (pipeline.cpp:41-143) and Histogram + radixSort (:4020-4144), cut out the same way for oracle/ref/ref_sortkeys_harness.cpp
awk '/^enum class DrawCommandTypes : u8/{p=1} /^struct Indirect \{/{p=0} p' "$REF/src/renderer/pipeline.cpp" > "$S/renderer/extracted_sort_key_packers.inl"
awk '/^\tstruct Histogram \{/{p=1} /^\tvoid viewport\(int x, int y, int w, int h\) override/{p=0} p' "$REF/src/renderer/pipeline.cpp" > "$S/renderer/extracted_radix_sort.inl"
cp "$REF/src/renderer/material.h" "$S/renderer/" 2>/dev/null || true
if grep -q makeAutoInstancedSortValue "$S/renderer/extracted_sort_key_packers.inl" && grep -q "void radixSort" "$S/renderer/extracted_radix_sort.inl" \
	&& $CXX $FL -c "$HERE/ref/ref_sortkeys_harness.cpp" -o "$TMP/obj/ref_sortkeys_harness.o" 2> "$TMP/sortkeys.log"; then
	EXTRA="$EXTRA $TMP/obj/ref_sortkeys_harness.o"
else
	echo "build_ref.sh: sort-key harness not built (see below); continuing without ref_radix_sort / ref_make_*" >&2
	tail -8 "$TMP/sortkeys.log" >&2 || true
fi
$CXX -shared -o "$OUT/libref_lumix.so" $OBJS "$TMP/obj/ref_stubs.o" "$TMP/obj/ref_harness.o" "$TMP/obj/ref_anim_harness.o" $EXTRA -lpthread -Wl,--no-undefined -Wl,--exclude-libs,ALL
# (7) the executed drop-in boundary: the SAME reference objects minus its CullingSystemImpl, plus lumixengine_b200/host/culling_system_b200.cpp
#     (CullingSystem::create / CullResult::free over liblumix_b200.so) and a harness that drives it through the abstract CullingSystem from
#     job-system fibers (oracle/ref/ref_engine_shim_harness.cpp).  Needs the product library to link against; skipped if it is not built yet.
PRODUCT_DIR="$(cd "$HERE/../lumixengine_b200" && pwd)"
if [ -f "$PRODUCT_DIR/liblumix_b200.so" ]; then
	SHIM_OBJS=""
	for o in $OBJS; do case "$o" in */core_*.o|*/engine_resource.o|*/renderer_pose.o|*/animation_animation.o) SHIM_OBJS="$SHIM_OBJS $o";; esac; done # job system, allocators, PageAllocator, math, geometry, log ...; Pose, Animation, Resource for the animation binding
	# the World patch as INTEGRATION.md section 2 describes it, applied to the overlay copy of the reference's own world.h / world.cpp:
	# declarations into `struct World`, host/world_b200.inl appended to world.cpp, one line in ~World
	cp "$REF/src/engine/world.cpp" "$S/engine/world.cpp"
	python3 - "$S/engine" "$PRODUCT_DIR/host" <<'PYEOF'
import sys
eng, host = sys.argv[1], sys.argv[2]
wh = open(eng + "/world.h").read()
anchor = "private:\n\tvoid transformEntity(EntityRef entity, bool update_local);"
assert anchor in wh, "world.h changed: INTEGRATION.md section 2 needs another anchor"
wh = wh.replace(anchor, open(host + "/world_b200_decl.inl").read() + anchor).replace("namespace Lumix {", "struct lb200_ctx; // include/lumix_b200.h\nnamespace Lumix {", 1)
open(eng + "/world.h", "w").write(wh)
wc = open(eng + "/world.cpp").read()
dtor = "World::~World() {\n"
assert dtor in wc
wc = wc.replace(dtor, dtor + "\tdestroyHierarchyB200();\n", 1) + "\n" + open(host + "/world_b200.inl").read()
open(eng + "/world.cpp", "w").write(wc)
PYEOF
	# the animation binding: the accessors of host/animation_b200_decl.inl go into `struct Animation` of a second copy
	# of animation.h that only the binding's harness sees (inline getters: the layout of Animation does not change, the reference's own
	# animation.o above was compiled from the unpatched header)
	mkdir -p "$TMP/patched/animation"
	python3 - "$S/animation/animation.h" "$TMP/patched/animation/animation.h" "$PRODUCT_DIR/host/animation_b200_decl.inl" <<'PYEOF'
import sys
src, dst, decl = sys.argv[1:4]
t = open(src).read()
anchor = "\tconst Array<TranslationTrack>& getTranslations() const { return m_translations; }"
assert anchor in t, "animation.h changed: INTEGRATION.md section 3 needs another anchor"
open(dst, "w").write(t.replace(anchor, open(decl).read() + anchor))
PYEOF
	if $CXX $FL -I"$HERE/../include" -c "$PRODUCT_DIR/host/culling_system_b200.cpp" -o "$TMP/obj/shim.o" 2> "$TMP/shim.log" \
		&& $CXX -I"$TMP/patched" $FL -I"$HERE/../include" -I"$PRODUCT_DIR/host" -c "$HERE/ref/ref_anim_shim_harness.cpp" -o "$TMP/obj/anim_harness.o" 2>> "$TMP/shim.log" \
		&& $CXX $FL -c "$HERE/ref/ref_engine_shim_harness.cpp" -o "$TMP/obj/shim_harness.o" 2>> "$TMP/shim.log" \
		&& $CXX $FL -I"$HERE/../include" -c "$S/engine/world.cpp" -o "$TMP/obj/world_b200.o" 2>> "$TMP/shim.log" \
		&& $CXX $FL -I"$HERE/../include" -c "$HERE/ref/ref_world_shim_harness.cpp" -o "$TMP/obj/world_harness.o" 2>> "$TMP/shim.log" \
		&& $CXX -shared -o "$OUT/libengine_shim_b200.so" $SHIM_OBJS "$TMP/obj/ref_stubs.o" "$TMP/obj/shim.o" "$TMP/obj/shim_harness.o" \
			"$TMP/obj/world_b200.o" "$TMP/obj/world_harness.o" "$TMP/obj/anim_harness.o" \
			-L"$PRODUCT_DIR" -llumix_b200 -Wl,-rpath,'$ORIGIN/../../lumixengine_b200' -lpthread -Wl,--no-undefined -Wl,--exclude-libs,ALL 2>> "$TMP/shim.log"; then
		echo "built $OUT/libengine_shim_b200.so"
	else
		echo "build_ref.sh: engine shim library not built (see below)" >&2
		tail -8 "$TMP/shim.log" >&2 || true
	fi
fi
( cd "$REF" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$OUT/REFERENCE_COMMIT"
echo "built $OUT/libref_lumix.so"
