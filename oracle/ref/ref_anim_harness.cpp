// TEST INFRASTRUCTURE — not product code.  Runs the REFERENCE'S OWN Animation::getRelativePose
// (src/animation/animation.cpp:117-204,294-311) and Pose::computeAbsolute (src/renderer/pose.cpp:66-133) on
// caller-supplied clip / skeleton data, to pin oracle/oracle_anim.c.
//
// Animation and Model are engine Resources that cannot be constructed without the whole engine, so this file reaches
// their data members through `#define private public` and fills zero-initialised storage by hand; only the member
// functions under test run (they read data members only).  No reference source is copied.
#define private public
#define protected public
#include "animation/animation.h"
#include "renderer/model.h"
#include "renderer/pose.h"
#undef private
#undef protected
#include "core/default_allocator.h"
#include "engine/resource_manager.h"

#include <new>
#include <stdint.h>
#include <string.h>

using namespace Lumix;

#define REF_API extern "C" __attribute__((visibility("default")))

// link-time stubs for the three symbols resource.cpp / animation.cpp reference but this harness never reaches
namespace Lumix {
ResourceManagerHub::LoadHook::Action ResourceManagerHub::onBeforeLoad(Resource&) const { return LoadHook::Action::IMMEDIATE; }
Resource* ResourceManagerHub::load(ResourceType, const Path&) { return nullptr; }
const ResourceType Model::TYPE("model");
}

struct RefTrack { // same layout as OracleTrack / lb200_track
	uint16_t bone_index, offset_bits;
	uint8_t bitsizes[3];
	uint8_t skipped_channel;
	float min[3], to_range[3];
};
struct RefConstT { uint16_t bone_index, pad; float value[3]; };
struct RefConstR { uint16_t bone_index, pad; float value[4]; };
struct RefClip {
	float fps;
	uint32_t frame_count, t_bits, r_bits, n_t, n_ct, n_r, n_cr;
	const RefTrack* t; const RefConstT* ct; const RefTrack* r; const RefConstR* cr;
	const uint8_t* t_stream; const uint8_t* r_stream;
};
struct RefSkeleton {
	uint32_t bone_count; int32_t first_nonroot;
	const int16_t* parents; const float* bind_relative7; const float* inverse_bind7;
};

template <typename T> struct RawStorage {
	RawStorage() { memset(mem, 0, sizeof(mem)); }
	T* get() { return reinterpret_cast<T*>(mem); }
	alignas(alignof(T)) unsigned char mem[sizeof(T)];
};

// pos: bone_count*3, rot: bone_count*4 (in: relative pose to blend onto when weight < 0.9999; out: result)
REF_API void ref_pose_evaluate(const RefSkeleton* sk, const RefClip* clip, uint32_t time_ticks, float weight, int start_from_bind, int compute_absolute,
	float* pos3, float* rot4)
{
	static DefaultAllocator allocator;
	RawStorage<Model> model_mem;
	Model* model = model_mem.get();
	new (&model->m_parents) Array<i16>(allocator);
	for (uint32_t i = 0; i < sk->bone_count; ++i) model->m_parents.push(sk->parents[i]);
	model->m_first_nonroot_bone_index = sk->first_nonroot;

	RawStorage<Animation> anim_mem;
	Animation* anim = anim_mem.get();
	new (&anim->m_translations) Array<Animation::TranslationTrack>(allocator);
	new (&anim->m_const_translations) Array<Animation::ConstTranslationTrack>(allocator);
	new (&anim->m_rotations) Array<Animation::RotationTrack>(allocator);
	new (&anim->m_const_rotations) Array<Animation::ConstRotationTrack>(allocator);
	for (uint32_t i = 0; i < clip->n_t; ++i) {
		Animation::TranslationTrack& t = anim->m_translations.emplace();
		t.bone_index = clip->t[i].bone_index;
		t.min = Vec3(clip->t[i].min[0], clip->t[i].min[1], clip->t[i].min[2]);
		t.to_range = Vec3(clip->t[i].to_range[0], clip->t[i].to_range[1], clip->t[i].to_range[2]);
		t.offset_bits = clip->t[i].offset_bits;
		memcpy(t.bitsizes, clip->t[i].bitsizes, 3);
	}
	for (uint32_t i = 0; i < clip->n_ct; ++i) {
		Animation::ConstTranslationTrack& t = anim->m_const_translations.emplace();
		t.bone_index = clip->ct[i].bone_index;
		t.value = Vec3(clip->ct[i].value[0], clip->ct[i].value[1], clip->ct[i].value[2]);
	}
	for (uint32_t i = 0; i < clip->n_r; ++i) {
		Animation::RotationTrack& t = anim->m_rotations.emplace();
		t.bone_index = clip->r[i].bone_index;
		t.min = Vec3(clip->r[i].min[0], clip->r[i].min[1], clip->r[i].min[2]);
		t.to_range = Vec3(clip->r[i].to_range[0], clip->r[i].to_range[1], clip->r[i].to_range[2]);
		t.offset_bits = clip->r[i].offset_bits;
		memcpy(t.bitsizes, clip->r[i].bitsizes, 3);
		t.skipped_channel = clip->r[i].skipped_channel;
	}
	for (uint32_t i = 0; i < clip->n_cr; ++i) {
		Animation::ConstRotationTrack& t = anim->m_const_rotations.emplace();
		t.bone_index = clip->cr[i].bone_index;
		t.value = Quat(clip->cr[i].value[0], clip->cr[i].value[1], clip->cr[i].value[2], clip->cr[i].value[3]);
	}
	anim->m_translation_stream = clip->t_stream;
	anim->m_rotation_stream = clip->r_stream;
	anim->m_translations_frame_size_bits = clip->t_bits;
	anim->m_rotations_frame_size_bits = clip->r_bits;
	anim->m_frame_count = clip->frame_count;
	anim->m_fps = clip->fps;
	anim->m_max_accessed_bone_index = 0;
	anim->m_root_motion.rotation_track_idx = -1;
	anim->m_root_motion.translation_track_idx = -1;

	{
		Pose pose(allocator);
		pose.resize((int)sk->bone_count);
		for (uint32_t i = 0; i < sk->bone_count; ++i) {
			if (start_from_bind) { // Model::getRelativePose, model.cpp:226-237
				const float* b = sk->bind_relative7 + 7 * i;
				pose.positions[i] = Vec3(b[0], b[1], b[2]);
				pose.rotations[i] = Quat(b[3], b[4], b[5], b[6]);
			}
			else {
				pose.positions[i] = Vec3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]);
				pose.rotations[i] = Quat(rot4[4 * i], rot4[4 * i + 1], rot4[4 * i + 2], rot4[4 * i + 3]);
			}
		}
		pose.is_absolute = false;
		Animation::SampleContext ctx;
		ctx.pose = &pose;
		ctx.model = model;
		ctx.time = Time(time_ticks);
		ctx.weight = weight;
		ctx.mask = nullptr;
		anim->getRelativePose(ctx);
		if (compute_absolute) pose.computeAbsolute(*model);
		for (uint32_t i = 0; i < sk->bone_count; ++i) {
			pos3[3 * i] = pose.positions[i].x; pos3[3 * i + 1] = pose.positions[i].y; pos3[3 * i + 2] = pose.positions[i].z;
			rot4[4 * i] = pose.rotations[i].x; rot4[4 * i + 1] = pose.rotations[i].y; rot4[4 * i + 2] = pose.rotations[i].z; rot4[4 * i + 3] = pose.rotations[i].w;
		}
	}
	anim->m_translations.~Array();
	anim->m_const_translations.~Array();
	anim->m_rotations.~Array();
	anim->m_const_rotations.~Array();
	model->m_parents.~Array();
}

static void fillPose(Pose& pose, uint32_t n, const float* pos3, const float* rot4) {
	pose.resize((int)n);
	for (uint32_t i = 0; i < n; ++i) {
		pose.positions[i] = Vec3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]);
		pose.rotations[i] = Quat(rot4[4 * i], rot4[4 * i + 1], rot4[4 * i + 2], rot4[4 * i + 3]);
	}
}
static void readPose(const Pose& pose, uint32_t n, float* pos3, float* rot4) {
	for (uint32_t i = 0; i < n; ++i) {
		pos3[3 * i] = pose.positions[i].x; pos3[3 * i + 1] = pose.positions[i].y; pos3[3 * i + 2] = pose.positions[i].z;
		rot4[4 * i] = pose.rotations[i].x; rot4[4 * i + 1] = pose.rotations[i].y; rot4[4 * i + 2] = pose.rotations[i].z; rot4[4 * i + 3] = pose.rotations[i].w;
	}
}

// Pose::blend (pose.cpp:30-41): a = a.blend(b, weight), in place in pos_a / rot_a
REF_API void ref_pose_blend(uint32_t bone_count, float* pos_a, float* rot_a, const float* pos_b, const float* rot_b, float weight) {
	static DefaultAllocator allocator;
	Pose a(allocator), b(allocator);
	fillPose(a, bone_count, pos_a, rot_a);
	fillPose(b, bone_count, pos_b, rot_b);
	a.blend(b, weight);
	readPose(a, bone_count, pos_a, rot_a);
}

// Pose::computeRelative (pose.cpp:136-146) on an absolute pose, in place
REF_API void ref_pose_compute_relative(const RefSkeleton* sk, float* pos3, float* rot4) {
	static DefaultAllocator allocator;
	RawStorage<Model> model_mem;
	Model* model = model_mem.get();
	new (&model->m_parents) Array<i16>(allocator);
	for (uint32_t i = 0; i < sk->bone_count; ++i) model->m_parents.push(sk->parents[i]);
	model->m_first_nonroot_bone_index = sk->first_nonroot;
	{
		Pose pose(allocator);
		fillPose(pose, sk->bone_count, pos3, rot4);
		pose.is_absolute = true;
		pose.computeRelative(*model);
		readPose(pose, sk->bone_count, pos3, rot4);
	}
	model->m_parents.~Array();
}

REF_API uint32_t ref_clip_length_ticks(float fps, uint32_t frame_count) {
	// Animation::getLength, animation.h:128
	return Time::fromSeconds(frame_count / fps).raw();
}

// Animation::load (animation.cpp:397-493) on a caller-supplied compiled .ani image, version SKELETON (7: no skeleton path in the file; the
// harness pre-sets m_skeleton so the loader does not go to the resource system).  Reports what the loader parsed: scalar fields, the
// four track arrays (name hash instead of the bone index, which onBeforeReady resolves later against the model) and where the two
// bit streams start inside its copy of the file body.
struct RefLoaded {
	int32_t ok;
	float fps;
	uint32_t frame_count, flags, t_bits, r_bits, n_t, n_ct, n_r, n_cr, t_stream_offset, r_stream_offset, mem_size;
};
REF_API void ref_animation_load(const uint8_t* image, uint32_t size, RefLoaded* out, uint64_t* t_hash, RefTrack* t, uint64_t* ct_hash, float* ct_value3,
	uint64_t* r_hash, RefTrack* r, uint64_t* cr_hash, float* cr_value4, uint32_t cap)
{
	static DefaultAllocator allocator;
	memset(out, 0, sizeof(*out));
	RawStorage<Animation> anim_mem;
	Animation* anim = anim_mem.get();
	new (&anim->m_translations) Array<Animation::TranslationTrack>(allocator);
	new (&anim->m_const_translations) Array<Animation::ConstTranslationTrack>(allocator);
	new (&anim->m_rotations) Array<Animation::RotationTrack>(allocator);
	new (&anim->m_const_rotations) Array<Animation::ConstRotationTrack>(allocator);
	new (&anim->m_mem) Array<u8>(allocator);
	RawStorage<Model> model_mem;
	anim->m_skeleton = model_mem.get(); // only tested for null by load()
	out->ok = anim->load(Span<const u8>(image, size)) ? 1 : 0;
	if (out->ok && anim->m_translations.size() <= (int)cap && anim->m_const_translations.size() <= (int)cap && anim->m_rotations.size() <= (int)cap
		&& anim->m_const_rotations.size() <= (int)cap)
	{
		out->fps = anim->m_fps; out->frame_count = anim->m_frame_count; out->flags = (uint32_t)anim->m_flags;
		out->t_bits = anim->m_translations_frame_size_bits; out->r_bits = anim->m_rotations_frame_size_bits;
		out->n_t = anim->m_translations.size(); out->n_ct = anim->m_const_translations.size();
		out->n_r = anim->m_rotations.size(); out->n_cr = anim->m_const_rotations.size();
		out->mem_size = anim->m_mem.size();
		out->t_stream_offset = (uint32_t)(anim->m_translation_stream - anim->m_mem.begin());
		out->r_stream_offset = (uint32_t)(anim->m_rotation_stream - anim->m_mem.begin());
		for (uint32_t i = 0; i < out->n_t; ++i) {
			const Animation::TranslationTrack& k = anim->m_translations[i];
			t_hash[i] = k.bone_name.getHashValue();
			t[i].bone_index = 0; t[i].offset_bits = k.offset_bits; memcpy(t[i].bitsizes, k.bitsizes, 3); t[i].skipped_channel = 0;
			t[i].min[0] = k.min.x; t[i].min[1] = k.min.y; t[i].min[2] = k.min.z;
			t[i].to_range[0] = k.to_range.x; t[i].to_range[1] = k.to_range.y; t[i].to_range[2] = k.to_range.z;
		}
		for (uint32_t i = 0; i < out->n_ct; ++i) {
			const Animation::ConstTranslationTrack& k = anim->m_const_translations[i];
			ct_hash[i] = k.bone_name.getHashValue();
			ct_value3[3 * i] = k.value.x; ct_value3[3 * i + 1] = k.value.y; ct_value3[3 * i + 2] = k.value.z;
		}
		for (uint32_t i = 0; i < out->n_r; ++i) {
			const Animation::RotationTrack& k = anim->m_rotations[i];
			r_hash[i] = k.bone_name.getHashValue();
			r[i].bone_index = 0; r[i].offset_bits = k.offset_bits; memcpy(r[i].bitsizes, k.bitsizes, 3); r[i].skipped_channel = k.skipped_channel;
			r[i].min[0] = k.min.x; r[i].min[1] = k.min.y; r[i].min[2] = k.min.z;
			r[i].to_range[0] = k.to_range.x; r[i].to_range[1] = k.to_range.y; r[i].to_range[2] = k.to_range.z;
		}
		for (uint32_t i = 0; i < out->n_cr; ++i) {
			const Animation::ConstRotationTrack& k = anim->m_const_rotations[i];
			cr_hash[i] = k.bone_name.getHashValue();
			cr_value4[4 * i] = k.value.x; cr_value4[4 * i + 1] = k.value.y; cr_value4[4 * i + 2] = k.value.z; cr_value4[4 * i + 3] = k.value.w;
		}
	}
	else out->ok = out->ok ? -1 : 0; // -1: parsed but over the caller's capacity
	anim->m_translations.~Array();
	anim->m_const_translations.~Array();
	anim->m_rotations.~Array();
	anim->m_const_rotations.~Array();
	anim->m_mem.~Array();
}

// The time step of AnimationModuleImpl::updateAnimable (animation_module.cpp:458-469) through the reference's own Time operators
// (animation.h:17-43) and Animation::getLength's expression (animation.h:128)
REF_API uint32_t ref_time_advance(uint32_t time_ticks, float time_delta, float fps, uint32_t frame_count) {
	const Time now(time_ticks);
	const Time l = Time::fromSeconds(frame_count / fps);
	if (time_delta > 0) return ((now + Time::fromSeconds(time_delta)) % l).raw();
	const Time dt = Time::fromSeconds(-time_delta) % l;
	return ((now + l - dt) % l).raw();
}

REF_API uint32_t ref_time_from_seconds(float s) { return Time::fromSeconds(s).raw(); }
