	// --- lumix_b200: added inside `struct Animation` (animation.h:124-142), next to getTranslations(): the frame rate
	// and the two packed key-frame streams Animation::load keeps private (animation.h:165-172) ---
	float getFPSB200() const { return m_fps; }
	const u8* getTranslationStreamB200() const { return m_translation_stream; }
	const u8* getRotationStreamB200() const { return m_rotation_stream; }
	u32 getStreamEndB200() const { return u32(m_mem.size()); } // both streams live in m_mem, which ends with the unpacker's 8 bytes of padding (animation.cpp:439)
	const u8* getStreamBaseB200() const { return m_mem.empty() ? nullptr : &m_mem[0]; }
	bool hasRootMotionTracksB200() const { return m_root_motion.rotation_track_idx >= 0 || m_root_motion.translation_track_idx >= 0; } // animation.cpp:33-37, 321
