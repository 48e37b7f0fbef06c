// Intra-kernel event profiler.  Parity: reference include/flashinfer/profiler.cuh:22-149 (PROFILER_INIT /
// EVENT_START / EVENT_END / EVENT_INSTANT writing (tag, globaltimer) pairs into a user buffer).
//
// Usage inside a kernel compiled with -DFIB200_ENABLE_PROFILER:
//   FIB_PROFILER_INIT(buf, group_id, num_groups, is_writer_thread);   // once, e.g. one writer lane per warp role
//   FIB_PROFILER_EVENT_START(kEventLoadK); ... FIB_PROFILER_EVENT_END(kEventLoadK);
// Buffer layout (uint64): [0] = (num_blocks << 32 | num_groups); then per (block, group) a ring of entries
//   entry = (tag << 32) | globaltimer_lo,  tag = event_id << 2 | type (0 start, 1 end, 2 instant).
// Without the define every macro compiles to nothing (zero overhead in production builds).
#pragma once
#include <cstdint>

namespace fib200 {
namespace profiler {

constexpr uint32_t kStart = 0, kEnd = 1, kInstant = 2;

struct Ctx {
  uint64_t* base;
  uint32_t cursor, stride, sm_block;
  bool writer;
};

__device__ __forceinline__ uint32_t globaltimer_lo() {
  uint32_t t;
  asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(t));
  return t;
}
__device__ __forceinline__ uint32_t smid() {
  uint32_t s;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  return s;
}

__device__ __forceinline__ void init(Ctx& c, uint64_t* buf, uint32_t group, uint32_t num_groups, bool writer,
                                     uint32_t max_events_per_group) {
  const uint32_t block = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const uint32_t nblocks = gridDim.x * gridDim.y * gridDim.z;
  if (block == 0 && group == 0 && writer) buf[0] = (uint64_t(nblocks) << 32) | num_groups;
  c.base = buf + 1 + (uint64_t(block) * num_groups + group) * (max_events_per_group + 1);
  c.cursor = 0;
  c.stride = max_events_per_group;
  c.writer = writer;
  c.sm_block = (smid() << 16) | (block & 0xffff);
  if (writer) c.base[0] = uint64_t(c.sm_block) << 32;  // header: sm id | block, count patched by emit()
}

__device__ __forceinline__ void emit(Ctx& c, uint32_t event, uint32_t type) {
  if (!c.writer || c.cursor >= c.stride) return;
  c.base[1 + c.cursor] = (uint64_t((event << 2) | type) << 32) | globaltimer_lo();
  ++c.cursor;
  c.base[0] = (uint64_t(c.sm_block) << 32) | c.cursor;
}

}  // namespace profiler
}  // namespace fib200

#ifdef FIB200_ENABLE_PROFILER
#define FIB_PROFILER_DECL fib200::profiler::Ctx __fib_prof_ctx;
#define FIB_PROFILER_INIT(buf, group, num_groups, writer, max_events) \
  fib200::profiler::init(__fib_prof_ctx, (buf), (group), (num_groups), (writer), (max_events))
#define FIB_PROFILER_EVENT_START(ev) fib200::profiler::emit(__fib_prof_ctx, (ev), fib200::profiler::kStart)
#define FIB_PROFILER_EVENT_END(ev) fib200::profiler::emit(__fib_prof_ctx, (ev), fib200::profiler::kEnd)
#define FIB_PROFILER_EVENT_INSTANT(ev) fib200::profiler::emit(__fib_prof_ctx, (ev), fib200::profiler::kInstant)
#else
#define FIB_PROFILER_DECL
#define FIB_PROFILER_INIT(buf, group, num_groups, writer, max_events)
#define FIB_PROFILER_EVENT_START(ev)
#define FIB_PROFILER_EVENT_END(ev)
#define FIB_PROFILER_EVENT_INSTANT(ev)
#endif
