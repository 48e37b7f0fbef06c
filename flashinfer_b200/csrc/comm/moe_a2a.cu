// MoE expert-parallel all-to-all over NVLink peer memory: dispatch (push) and combine (pull + top-k reduce).
//
// Parity: reference MoeAlltoAll (flashinfer/comm/trtllm_moe_alltoall.py:202-743) and its kernels
// moeA2ADispatchKernel / moeA2ACombineKernel (csrc/nv_internal/.../moeAlltoAllKernels.cu:308,823).
//
// B200 design: every rank owns one symmetric workspace (same layout on all ranks, peers mapped through NVLink):
//     ctrl      : recv_count[world] | dispatch_flag[world] | combine_flag[world]            (int32 each)
//     payload p : [world (source rank)][max_tokens][bytes_p]       p < num_payloads (hidden, scales, ids, weights)
//     combine   : [world (source rank)][max_tokens][hidden]        expert outputs, written by the local MoE
// dispatch: one warp per token de-duplicates the target ranks of its top-k experts, reserves a slot per target
// with a local atomic and PUSHES the payload rows with 16-byte stores straight into the peers' payload regions
// (NVSwitch gives every peer full bandwidth, so there is no staging or FIFO).  The last CTA publishes the per-peer
// counts and an epoch flag with st.release.sys and waits for the flags of all sources.
// combine: every CTA acquires the peers' "outputs ready" flags, then one warp per token PULLS the rows of its
// distinct target ranks with L1-bypassing 16-byte loads and sums them (the expert-weighted top-k reduction for
// experts on the same rank already happened there).  Epochs live in device memory: CUDA-graph replay safe.
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

constexpr int kMaxWorld = 16;
constexpr int kMaxPayloads = 4;
constexpr int kMaxTopK = 32;

struct Layout {
  int64_t ctrl_off;
  int64_t payload_off[kMaxPayloads];
  int64_t combine_off;
  int32_t payload_bytes[kMaxPayloads];
  int32_t num_payloads;
  int32_t max_tokens;
};

struct Peers {
  uint8_t* base[kMaxWorld];
};

__device__ __forceinline__ int4 ld_peer16(const void* p) {
  int4 v;
  asm volatile("ld.global.relaxed.sys.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_i32(int32_t* p, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_sys_i32(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// state (local device memory, int32): [0] dispatch epoch, [1] combine epoch, [2] done counter (dispatch),
// [3] done counter (combine), [4 .. 4+world) send counters
__global__ void __launch_bounds__(256)
moe_a2a_dispatch_kernel(const int32_t* __restrict__ topk_ids, int T, int K, int experts_per_rank, int rank, int world,
                        Peers peers, Layout L, const uint8_t* p0, const uint8_t* p1, const uint8_t* p2, const uint8_t* p3,
                        int32_t* __restrict__ token_slot, int32_t* __restrict__ state, int32_t* __restrict__ recv_counts_out) {
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int gw = blockIdx.x * warps + (threadIdx.x >> 5), nw = gridDim.x * warps;
  const uint8_t* src[kMaxPayloads] = {p0, p1, p2, p3};
  int32_t* send_cnt = state + 4;
  ptx::grid_dep_wait();
  for (int t = gw; t < T; t += nw) {
    int r = -1;
    if (lane < K) {
      const int e = topk_ids[int64_t(t) * K + lane];
      if (e >= 0 && e < experts_per_rank * world) r = e / experts_per_rank;
    }
    // first occurrence of each target rank among the K lanes
    const uint32_t same = __match_any_sync(0xffffffffu, r);
    const int first_lane = __ffs(same) - 1;
    const bool first = (r >= 0) && (first_lane == lane);
    int slot = -1;
    if (first) slot = atomicAdd(&send_cnt[r], 1);
    slot = __shfl_sync(0xffffffffu, slot, first_lane);
    if (lane < K) token_slot[int64_t(t) * K + lane] = (r >= 0 && slot < L.max_tokens) ? slot : -1;
    uint32_t todo = __ballot_sync(0xffffffffu, first && slot < L.max_tokens);
    while (todo) {
      const int l = __ffs(todo) - 1;
      todo &= todo - 1;
      const int dst_rank = __shfl_sync(0xffffffffu, r, l);
      const int dst_slot = __shfl_sync(0xffffffffu, slot, l);
#pragma unroll
      for (int p = 0; p < kMaxPayloads; ++p) {
        if (p >= L.num_payloads) break;
        const int bytes = L.payload_bytes[p];
        const int4* s = reinterpret_cast<const int4*>(src[p] + int64_t(t) * bytes);
        int4* d = reinterpret_cast<int4*>(peers.base[dst_rank] + L.payload_off[p] +
                                          (int64_t(rank) * L.max_tokens + dst_slot) * bytes);
        for (int v = lane; v < bytes / 16; v += 32) d[v] = s[v];
      }
    }
  }
  // ---- completion: publish counts + flag to every peer, then wait for every source
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(&state[2], 1) == int(gridDim.x) - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    const int epoch = state[0] + 1;
    if (int(threadIdx.x) < world) {
      const int peer = threadIdx.x;
      int32_t* ctrl = reinterpret_cast<int32_t*>(peers.base[peer] + L.ctrl_off);
      int c = send_cnt[peer];
      if (c > L.max_tokens) c = L.max_tokens;
      ctrl[rank] = c;                                   // recv_count[src = me] on the peer
      st_release_sys_i32(ctrl + kMaxWorld + rank, epoch);  // dispatch_flag[src = me]
      send_cnt[peer] = 0;
    }
    __syncthreads();
    if (int(threadIdx.x) < world) {
      const int32_t* ctrl = reinterpret_cast<const int32_t*>(peers.base[rank] + L.ctrl_off);
      while (ld_acquire_sys_i32(ctrl + kMaxWorld + threadIdx.x) < epoch) {
      }
      recv_counts_out[threadIdx.x] = ctrl[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      state[0] = epoch;
      state[2] = 0;
    }
  }
  ptx::grid_dep_launch();
}

template <typename T>
__global__ void __launch_bounds__(256)
moe_a2a_combine_kernel(const int32_t* __restrict__ topk_ids, const int32_t* __restrict__ token_slot, T* __restrict__ out,
                       int Tn, int K, int hidden, int experts_per_rank, int rank, int world, Peers peers, Layout L,
                       int32_t* __restrict__ state) {
  constexpr int VN = 16 / sizeof(T);
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int gw = blockIdx.x * warps + (threadIdx.x >> 5), nw = gridDim.x * warps;
  ptx::grid_dep_wait();
  const int epoch = state[1] + 1;
  // "my expert outputs are in my combine region" -> every peer (once per kernel), then wait for all peers
  if (blockIdx.x == 0 && int(threadIdx.x) < world) {
    __threadfence_system();
    int32_t* ctrl = reinterpret_cast<int32_t*>(peers.base[threadIdx.x] + L.ctrl_off);
    st_release_sys_i32(ctrl + 2 * kMaxWorld + rank, epoch);
  }
  if (int(threadIdx.x) < world) {
    const int32_t* ctrl = reinterpret_cast<const int32_t*>(peers.base[rank] + L.ctrl_off);
    while (ld_acquire_sys_i32(ctrl + 2 * kMaxWorld + threadIdx.x) < epoch) {
    }
  }
  __syncthreads();
  const int64_t row_bytes = int64_t(hidden) * sizeof(T);
  for (int t = gw; t < Tn; t += nw) {
    int r = -1, slot = -1;
    if (lane < K) {
      const int e = topk_ids[int64_t(t) * K + lane];
      slot = token_slot[int64_t(t) * K + lane];
      if (e >= 0 && e < experts_per_rank * world && slot >= 0) r = e / experts_per_rank;
    }
    const uint32_t same = __match_any_sync(0xffffffffu, r);
    const bool first = (r >= 0) && (__ffs(same) - 1 == lane);
    const uint32_t todo0 = __ballot_sync(0xffffffffu, first);
    for (int v = lane; v < hidden / VN; v += 32) {
      float acc[VN];
#pragma unroll
      for (int e = 0; e < VN; ++e) acc[e] = 0.f;
      uint32_t todo = todo0;
      while (todo) {
        const int l = __ffs(todo) - 1;
        todo &= todo - 1;
        const int sr = __shfl_sync(0xffffffffu, r, l);
        const int ss = __shfl_sync(0xffffffffu, slot, l);
        const uint8_t* p = peers.base[sr] + L.combine_off + (int64_t(rank) * L.max_tokens + ss) * row_bytes + int64_t(v) * 16;
        const int4 raw = ld_peer16(p);
        const T* vals = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[e] += to_f32(vals[e]);
      }
      Vec16<T> o;
#pragma unroll
      for (int e = 0; e < VN; ++e) o.v[e] = from_f32<T>(acc[e]);
      st16(out + int64_t(t) * hidden + int64_t(v) * VN, o);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(&state[3], 1) == int(gridDim.x) - 1) {
      state[1] = epoch;
      state[3] = 0;
    }
  }
  ptx::grid_dep_launch();
}

// Marks the expert ids of rows that were not received (row >= recv_count[src]) as invalid.
__global__ void moe_a2a_sanitize_kernel(int32_t* __restrict__ ids, const int32_t* __restrict__ recv_counts, int world,
                                        int max_tokens, int K, int invalid) {
  const int64_t n = int64_t(world) * max_tokens * K;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int src = int(i / (int64_t(max_tokens) * K));
    const int row = int((i / K) % max_tokens);
    if (row >= recv_counts[src]) ids[i] = invalid;
  }
}

Layout make_layout(const int64_t* lay) {
  Layout L;
  L.ctrl_off = lay[0];
  L.combine_off = lay[1];
  L.max_tokens = (int)lay[2];
  L.num_payloads = (int)lay[3];
  for (int p = 0; p < kMaxPayloads; ++p) {
    L.payload_off[p] = lay[4 + p];
    L.payload_bytes[p] = (int)lay[8 + p];
  }
  return L;
}

}  // namespace

// peer_ptrs: host int64[world]; layout: host int64[12] = {ctrl_off, combine_off, max_tokens, num_payloads,
// payload_off[4], payload_bytes[4]}.
extern "C" int moe_a2a_dispatch(void* topk_ids, int64_t T, int64_t K, int64_t experts_per_rank, int64_t rank, int64_t world,
                                void* peer_ptrs_host, void* layout_host, void* p0, void* p1, void* p2, void* p3,
                                void* token_slot, void* state, void* recv_counts, int64_t pdl, int64_t stream_) {
  FIB_CHECK(world >= 1 && world <= kMaxWorld, "moe_a2a: world size must be in [1,16]");
  FIB_CHECK(K >= 1 && K <= kMaxTopK, "moe_a2a: top_k must be in [1,32]");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  Peers peers;
  for (int i = 0; i < world; ++i) peers.base[i] = reinterpret_cast<uint8_t*>(((const int64_t*)peer_ptrs_host)[i]);
  const Layout L = make_layout((const int64_t*)layout_host);
  for (int p = 0; p < L.num_payloads; ++p) FIB_CHECK(L.payload_bytes[p] % 16 == 0, "moe_a2a: payload rows must be multiples of 16 bytes");
  int grid = (int)((T + 7) / 8);
  if (grid < 1) grid = 1;
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
  LaunchCfg lc(dim3(grid), dim3(256), 0, stream, pdl != 0);
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, moe_a2a_dispatch_kernel, (const int32_t*)topk_ids, (int)T, (int)K,
                                    (int)experts_per_rank, (int)rank, (int)world, peers, L, (const uint8_t*)p0,
                                    (const uint8_t*)p1, (const uint8_t*)p2, (const uint8_t*)p3, (int32_t*)token_slot,
                                    (int32_t*)state, (int32_t*)recv_counts));
  return 0;
}

extern "C" int moe_a2a_combine(void* topk_ids, void* token_slot, void* out, int64_t T, int64_t K, int64_t hidden,
                               int64_t experts_per_rank, int64_t rank, int64_t world, void* peer_ptrs_host,
                               void* layout_host, void* state, int64_t dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(world >= 1 && world <= kMaxWorld, "moe_a2a: world size must be in [1,16]");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  Peers peers;
  for (int i = 0; i < world; ++i) peers.base[i] = reinterpret_cast<uint8_t*>(((const int64_t*)peer_ptrs_host)[i]);
  const Layout L = make_layout((const int64_t*)layout_host);
  int grid = (int)((T + 7) / 8);
  if (grid < 1) grid = 1;
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
  return FIB_DISPATCH_HALF(dtype, Tt, [&]() -> int {
    FIB_CHECK(hidden % (16 / (int)sizeof(Tt)) == 0, "moe_a2a: hidden must be a multiple of 8");
    LaunchCfg lc(dim3(grid), dim3(256), 0, stream, pdl != 0);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, moe_a2a_combine_kernel<Tt>, (const int32_t*)topk_ids,
                                      (const int32_t*)token_slot, (Tt*)out, (int)T, (int)K, (int)hidden,
                                      (int)experts_per_rank, (int)rank, (int)world, peers, L, (int32_t*)state));
    return 0;
  });
}

extern "C" int moe_a2a_sanitize(void* ids, void* recv_counts, int64_t world, int64_t max_tokens, int64_t K,
                                int64_t invalid, int64_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ++launch_counter();
  moe_a2a_sanitize_kernel<<<num_sms(), 256, 0, stream>>>((int32_t*)ids, (const int32_t*)recv_counts, (int)world,
                                                         (int)max_tokens, (int)K, (int)invalid);
  FIB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
