// AnimablesB200 — the engine-side binding of lb200_animation_* (INTEGRATION.md §3).
//
// This file is the patch a LumixEngine maintainer adds to src/animation/animation_module.cpp (after the includes; AnimationModuleImpl gets
// one `AnimablesB200<RenderModule> m_animables_b200` member and calls update() from updateAnimables (animation_module.cpp:737-749) instead
// of looping over updateAnimable (:439-472)), together with the four accessors of animation_b200_decl.inl inside `struct Animation`.
// tests/test_integration_compile.py applies exactly that to a temporary copy of the reference headers and instantiates the template with the
// engine's RenderModule; oracle/build_ref.sh instantiates it with a two-function stand-in over REAL Model / Animation / Pose objects and
// oracle/ref/ref_anim_shim_harness.cpp runs it next to the reference's own per-animable loop (tests/test_engine_boundary_gpu.py).
//
// What it replaces, per animable: Model::getRelativePose (model.cpp:226-237) -> Animation::getRelativePose (animation.cpp:117-204) ->
// Pose::computeAbsolute (pose.cpp:66-133) -> the time step of :458-469.  Here: animables are grouped by model, one lb200_animation per
// group (the group's clips = the distinct Animation resources its animables play), ONE device pass per group and frame, then the absolute
// poses go back through RenderModule::lockPose / unlockPose (render_module.h:402-403, pose.h:15-31) exactly like the reference writes them.

#include "lumix_b200.h"

namespace Lumix {

template <typename RenderSide> // RenderModule, or anything with getModelInstanceModel / lockPose / unlockPose of the same meaning
struct AnimablesB200 {
	struct Group {
		Group(IAllocator& a) : clips(a), members(a), clip_index(a), time_ticks(a), pos(a), rot(a) {}
		Model* model = nullptr;
		lb200_animation* handle = nullptr;
		u32 capacity = 0;          // max_instances the handle was created with
		u32 clips_built = 0;       // clips.size() the handle was created with
		Array<Animation*> clips;
		Array<u32> members;        // indices into the animables span
		Array<u32> clip_index, time_ticks;
		Array<float> pos, rot;     // read-back: bone_count * 3 / * 4 floats per member
	};

	explicit AnimablesB200(IAllocator& allocator) : m_allocator(allocator), m_groups(allocator), m_reference_path(allocator) {}
	// animables update() left alone because their clip has root-motion tracks (Animation::getRelativePose substitutes its own pose arrays for
	// those tracks, animation.cpp:33-37 — not part of lb200_clip): the module runs them through updateAnimable as before
	const Array<u32>& referencePath() const { return m_reference_path; }
	~AnimablesB200() { clear(); }
	void clear() {
		for (Group* g : m_groups) { lb200_animation_destroy(g->handle); LUMIX_DELETE(m_allocator, g); }
		m_groups.clear();
	}

	// One updateAnimables pass.  Returns false (and logs) if the library reported an error; animables already written back stay written.
	bool update(lb200_ctx* ctx, RenderSide& render, Span<Animable> animables, float time_delta) {
		for (Group* g : m_groups) g->members.clear();
		m_reference_path.clear();
		for (u32 i = 0; i < animables.length(); ++i) { // the early-outs of updateAnimable, animation_module.cpp:440-445
			Animable& a = animables[i];
			if (!a.animation || !a.animation->isReady()) continue;
			if (a.animation->hasRootMotionTracksB200()) { m_reference_path.push(i); continue; }
			Model* model = render.getModelInstanceModel(a.entity);
			if (!model || !model->isReady()) continue;
			Group* g = nullptr;
			for (Group* it : m_groups) if (it->model == model) { g = it; break; }
			if (!g) {
				g = LUMIX_NEW(m_allocator, Group)(m_allocator);
				g->model = model;
				m_groups.push(g);
			}
			i32 clip = g->clips.indexOf(a.animation);
			if (clip < 0) { clip = g->clips.size(); g->clips.push(a.animation); }
			g->members.push(i);
			g->clip_index.resize(g->members.size());
			g->time_ticks.resize(g->members.size());
			g->clip_index[g->members.size() - 1] = (u32)clip;
			g->time_ticks[g->members.size() - 1] = a.time.raw();
		}
		bool ok = true;
		for (Group* g : m_groups) {
			const u32 n = (u32)g->members.size();
			if (n == 0) continue;
			if (!g->handle || g->capacity < n || g->clips_built != (u32)g->clips.size()) {
				if (!build(ctx, *g, n)) { ok = false; continue; }
			}
			const u32 bones = (u32)g->model->getBones().length();
			g->pos.resize(n * bones * 3);
			g->rot.resize(n * bones * 4);
			if (lb200_animation_set_instances(g->handle, g->clip_index.begin(), g->time_ticks.begin(), n) != LB200_OK
				|| lb200_animation_update(g->handle, time_delta, LB200_PALETTE_POSE) != LB200_OK
				|| lb200_animation_get_pose(g->handle, 0, n, g->pos.begin(), g->rot.begin()) != LB200_OK
				|| lb200_animation_get_times(g->handle, 0, n, g->time_ticks.begin()) != LB200_OK)
			{
				logError("lumix_b200 animation: ", lb200_last_error(ctx));
				ok = false;
				continue;
			}
			for (u32 k = 0; k < n; ++k) {
				Animable& a = animables[g->members[k]];
				Pose* pose = render.lockPose(a.entity); // animation_module.cpp:447-448
				if (!pose) continue;
				if (pose->count == bones) {
					memcpy(pose->positions, &g->pos[(size_t)k * bones * 3], sizeof(Vec3) * bones);
					memcpy(pose->rotations, &g->rot[(size_t)k * bones * 4], sizeof(Quat) * bones);
					pose->is_absolute = true; // what Pose::computeAbsolute leaves (pose.cpp:132)
					a.time = Time(g->time_ticks[k]); // :458-469
				}
				render.unlockPose(a.entity, true); // :471
			}
		}
		return ok;
	}

private:
	static lb200_track toTrack(const Animation::TranslationTrack& t) {
		lb200_track o = {};
		o.bone_index = t.bone_index; o.offset_bits = t.offset_bits;
		for (int c = 0; c < 3; ++c) o.bitsizes[c] = t.bitsizes[c];
		o.min[0] = t.min.x; o.min[1] = t.min.y; o.min[2] = t.min.z;
		o.to_range[0] = t.to_range.x; o.to_range[1] = t.to_range.y; o.to_range[2] = t.to_range.z;
		return o;
	}
	static lb200_track toTrack(const Animation::RotationTrack& t) {
		lb200_track o = {};
		o.bone_index = t.bone_index; o.offset_bits = t.offset_bits; o.skipped_channel = t.skipped_channel;
		for (int c = 0; c < 3; ++c) o.bitsizes[c] = t.bitsizes[c];
		o.min[0] = t.min.x; o.min[1] = t.min.y; o.min[2] = t.min.z;
		o.to_range[0] = t.to_range.x; o.to_range[1] = t.to_range.y; o.to_range[2] = t.to_range.z;
		return o;
	}

	// (re)create the group's lb200_animation: skeleton from Model (model.h:154-166, 225-244), clips from the loaded Animation resources
	bool build(lb200_ctx* ctx, Group& g, u32 n) {
		lb200_animation_destroy(g.handle);
		g.handle = nullptr;
		Model& model = *g.model;
		const u32 bones = (u32)model.getBones().length();
		Array<i16> parents(m_allocator);
		Array<float> bind(m_allocator), inv(m_allocator);
		parents.resize(bones); bind.resize(bones * 7); inv.resize(bones * 7);
		const SOATransform& ib = model.getInverseBindPose();
		for (u32 b = 0; b < bones; ++b) {
			parents[b] = model.getBoneParent(b);
			const LocalRigidTransform& r = model.getBone(b).relative_transform;
			float* o = &bind[b * 7];
			o[0] = r.pos.x; o[1] = r.pos.y; o[2] = r.pos.z; o[3] = r.rot.x; o[4] = r.rot.y; o[5] = r.rot.z; o[6] = r.rot.w;
			float* q = &inv[b * 7];
			q[0] = ib.px[b]; q[1] = ib.py[b]; q[2] = ib.pz[b]; q[3] = ib.rx[b]; q[4] = ib.ry[b]; q[5] = ib.rz[b]; q[6] = ib.rw[b];
		}
		lb200_skeleton sk = {};
		sk.bone_count = bones;
		sk.first_nonroot_bone_index = model.getFirstNonrootBoneIndex();
		sk.parents = parents.begin(); sk.bind_relative7 = bind.begin(); sk.inverse_bind7 = inv.begin();

		const u32 nc = (u32)g.clips.size();
		Array<lb200_clip> clips(m_allocator);
		Array<Array<lb200_track>*> tracks(m_allocator); // owners of the converted track tables until the create call returns
		Array<Array<lb200_const_translation>*> cts(m_allocator);
		Array<Array<lb200_const_rotation>*> crs(m_allocator);
		clips.resize(nc);
		for (u32 c = 0; c < nc; ++c) {
			const Animation& a = *g.clips[c];
			auto* tt = LUMIX_NEW(m_allocator, Array<lb200_track>)(m_allocator);
			auto* rt = LUMIX_NEW(m_allocator, Array<lb200_track>)(m_allocator);
			auto* ct = LUMIX_NEW(m_allocator, Array<lb200_const_translation>)(m_allocator);
			auto* cr = LUMIX_NEW(m_allocator, Array<lb200_const_rotation>)(m_allocator);
			tracks.push(tt); tracks.push(rt); cts.push(ct); crs.push(cr);
			for (const Animation::TranslationTrack& t : a.getTranslations()) tt->push(toTrack(t));
			for (const Animation::RotationTrack& t : a.getRotations()) rt->push(toTrack(t));
			for (const Animation::ConstTranslationTrack& t : a.getConstTranslations()) {
				lb200_const_translation o = {};
				o.bone_index = t.bone_index; o.value[0] = t.value.x; o.value[1] = t.value.y; o.value[2] = t.value.z;
				ct->push(o);
			}
			for (const Animation::ConstRotationTrack& t : a.getConstRotations()) {
				lb200_const_rotation o = {};
				o.bone_index = t.bone_index; o.value[0] = t.value.x; o.value[1] = t.value.y; o.value[2] = t.value.z; o.value[3] = t.value.w;
				cr->push(o);
			}
			lb200_clip& k = clips[c];
			k = {};
			k.fps = a.getFPSB200();
			k.frame_count = a.getFramesCount();
			k.translations_frame_size_bits = a.getTranslationFrameSizeBits();
			k.rotations_frame_size_bits = a.getRotationFrameSizeBits();
			k.n_translations = (u32)tt->size(); k.n_const_translations = (u32)ct->size(); k.n_rotations = (u32)rt->size(); k.n_const_rotations = (u32)cr->size();
			k.translations = tt->begin(); k.const_translations = ct->begin(); k.rotations = rt->begin(); k.const_rotations = cr->begin();
			// the streams: from their first byte to the end of Animation::m_mem, which includes the unpacker's padding (animation.cpp:439)
			const u8* end = a.getStreamBaseB200() + a.getStreamEndB200();
			k.translation_stream = a.getTranslationStreamB200();
			k.translation_stream_bytes = k.translation_stream ? u32(end - k.translation_stream) : 0;
			k.rotation_stream = a.getRotationStreamB200();
			k.rotation_stream_bytes = k.rotation_stream ? u32(end - k.rotation_stream) : 0;
		}
		u32 cap = g.capacity ? g.capacity : 64;
		while (cap < n) cap *= 2;
		const int rc = lb200_animation_create(ctx, &sk, clips.begin(), nc, nullptr, cap, &g.handle);
		for (auto* p : tracks) LUMIX_DELETE(m_allocator, p);
		for (auto* p : cts) LUMIX_DELETE(m_allocator, p);
		for (auto* p : crs) LUMIX_DELETE(m_allocator, p);
		if (rc != LB200_OK) {
			logError("lumix_b200 animation: ", lb200_last_error(ctx));
			g.handle = nullptr;
			return false;
		}
		g.capacity = cap;
		g.clips_built = nc;
		return true;
	}

	IAllocator& m_allocator;
	Array<Group*> m_groups;
	Array<u32> m_reference_path;
};

} // namespace Lumix
