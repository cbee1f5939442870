// TEST INFRASTRUCTURE — not product code.
//
// Link-time stubs that let the reference's own culling_system.cpp / job_system.cpp /
// page_allocator.cpp link on Linux without the (bit-rotted, X11-dependent) os.cpp and the
// Windows-only profiler.cpp.  See SURVEY.md §8(c) "With an overlay?" for why each is needed.
// Nothing here does arithmetic on the hot path: profiler calls are no-ops, the os:: calls are
// mmap, the atomics are the ones missing from the reference's src/core/linux/atomic.cpp.
#include "core/atomic.h"
#include "core/os.h"
#include "core/profiler.h"

#include <pthread.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

namespace Lumix {

namespace os {
void* memReserve(size_t size) {
	void* p = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	return p == MAP_FAILED ? nullptr : p;
}
void memCommit(void*, size_t) {}
void memRelease(void* ptr, size_t size) { munmap(ptr, size); }
u32 getMemPageAlignment() { return (u32)sysconf(_SC_PAGESIZE); }
ThreadID getCurrentThreadID() { return pthread_self(); }
u64 Timer::getRawTimestamp() {
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return u64(ts.tv_sec) * 1000000000ull + u64(ts.tv_nsec);
}
} // namespace os

namespace profiler {
void setThreadName(const char*) {}
void showInProfiler(bool) {}
void beginBlock(const char*) {}
void beginJob(i32) {}
void blockColor(u32) {}
void endBlock() {}
void pushInt(const char*, int) {}
void beforeFiberSwitch() {}
void signalTriggered(i32) {}
FiberSwitchData beginFiberWait(i32) { return {}; }
void endFiberWait(const FiberSwitchData&) {}
} // namespace profiler

i32 AtomicI32::setBits(i32 v) { return __atomic_fetch_or(&value, v, __ATOMIC_ACQ_REL); }
i32 AtomicI32::clearBits(i32 v) { return __atomic_fetch_and(&value, ~v, __ATOMIC_ACQ_REL); }
bool AtomicI32::compareExchange(volatile i32* value, i32 exchange, i32 comperand) {
	return __sync_bool_compare_and_swap(value, comperand, exchange);
}
i64 AtomicI64::exchange(i64 new_value) { return __atomic_exchange_n(&value, new_value, __ATOMIC_ACQ_REL); }
i64 AtomicI64::setBits(i64 v) { return __atomic_fetch_or(&value, v, __ATOMIC_ACQ_REL); }
i64 AtomicI64::clearBits(i64 v) { return __atomic_fetch_and(&value, ~v, __ATOMIC_ACQ_REL); }
bool AtomicI64::bitTestAndSet(u32 bit_position) {
	const i64 mask = i64(1) << bit_position;
	return (__atomic_fetch_or(&value, mask, __ATOMIC_ACQ_REL) & mask) == 0; // true = the bit was clear and is now ours (win/atomic.cpp:47: !_interlockedbittestandset64)
}
void* exchangePtr(void* volatile* value, void* exchange) { return __atomic_exchange_n(value, exchange, __ATOMIC_ACQ_REL); }
bool compareExchangePtr(void* volatile* value, void* exchange, void* comperand) {
	return __sync_bool_compare_and_swap(value, comperand, exchange);
}
void readBarrier() { __atomic_thread_fence(__ATOMIC_ACQUIRE); }
void writeBarrier() { __atomic_thread_fence(__ATOMIC_RELEASE); }
void cpuRelax() { __builtin_ia32_pause(); }

} // namespace Lumix
