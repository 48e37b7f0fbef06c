// NVLS / NVLink collectives that are not all-reduce: all-gather, reduce-scatter and the all-to-all used by
// decode context parallelism.  No NCCL on these paths.
//
// Parity: reference mixed_comm kernels (include/flashinfer/comm/mixed_comm.cuh: allgather_kernel :1842,
// reducescatter_kernel :2276 — multimem.st / multimem.ld_reduce), decode_cp_a2a_alltoall
// (csrc/nv_internal/tensorrt_llm/kernels/helixAllToAll.cu:287) and the NCCL fallbacks they replace.
//
//   all_gather     : every rank multicast-stores (multimem.st) its shard into slot `rank` of the symmetric output;
//                    one 16-byte store is replicated to all GPUs by the switch.  P2P fallback: store to each peer.
//   reduce_scatter : inputs live in the symmetric heap; rank r pulls the in-switch sum (multimem.ld_reduce) of
//                    shard r only.  P2P fallback: 16-byte loads from every peer, fp32 accumulate.
//   all_to_all     : in [rows, world, bytes] -> peer j's out[rows, me, bytes] with plain 16-byte peer stores
//                    (DCP: partial attention outputs + softmax statistics travel in one launch).
// Barriers: per-CTA epoch slots (st.release.sys / ld.acquire.sys), epochs in device memory (graph-replay safe).
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

constexpr int kMaxRanks = 16;

struct CollParams {
  uint8_t* peer_buf[kMaxRanks];   // symmetric data region of every rank
  uint32_t* peer_sig[kMaxRanks];  // signal pads: [2][max_blocks][world]
  uint8_t* mc_buf;                // multicast alias of the data region (or null)
  uint32_t* epochs;               // local [2][max_blocks]
  int rank, world, max_blocks;
};

__device__ __forceinline__ void coll_barrier(const CollParams& p, int phase) {
  __threadfence_system();
  __syncthreads();
  if (int(threadIdx.x) < p.world) {
    const int peer = threadIdx.x;
    const int slot_base = (phase * p.max_blocks + blockIdx.x) * p.world;
    const uint32_t epoch = p.epochs[phase * p.max_blocks + blockIdx.x] + 1;
    ptx::st_release_sys(p.peer_sig[peer] + slot_base + p.rank, epoch);
    const uint32_t* mine = p.peer_sig[p.rank] + slot_base + peer;
    while (int32_t(ptx::ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) p.epochs[phase * p.max_blocks + blockIdx.x] += 1;
}

__device__ __forceinline__ int4 ld_sys16(const void* ptr) {
  int4 v;
  asm volatile("ld.global.relaxed.sys.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
  return v;
}

// out region: [world][shard_bytes] at out_off; src: local shard
__global__ void __launch_bounds__(512) all_gather_kernel(const CollParams p, const int4* __restrict__ src, int64_t out_off,
                                                          int64_t shard_vecs) {
  ptx::grid_dep_wait();
  coll_barrier(p, 0);  // peers finished reading the previous contents of the output region
  const int64_t base = out_off / 16 + int64_t(p.rank) * shard_vecs;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < shard_vecs; i += int64_t(gridDim.x) * blockDim.x) {
    const int4 v = src[i];
    if (p.mc_buf) {
      ptx::multimem_st_v4(reinterpret_cast<int4*>(p.mc_buf) + base + i, v);
    } else {
      for (int r = 0; r < p.world; ++r) reinterpret_cast<int4*>(p.peer_buf[r])[base + i] = v;
    }
  }
  coll_barrier(p, 1);
  ptx::grid_dep_launch();
}

// in region: [world][shard] at in_off on every rank; out: local [shard]
template <typename T>
__global__ void __launch_bounds__(512) reduce_scatter_kernel(const CollParams p, T* __restrict__ out, int64_t in_off,
                                                             int64_t shard_vecs) {
  constexpr int VN = 16 / sizeof(T);
  ptx::grid_dep_wait();
  coll_barrier(p, 0);  // every rank has written its input
  const int64_t base = in_off / 16 + int64_t(p.rank) * shard_vecs;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < shard_vecs; i += int64_t(gridDim.x) * blockDim.x) {
    float acc[VN];
    if (p.mc_buf) {
      const void* a = reinterpret_cast<const int4*>(p.mc_buf) + base + i;
      if constexpr (sizeof(T) == 4) {
        const float4 f = ptx::multimem_ld_reduce_f32x4(a);
        acc[0] = f.x; acc[1] = f.y; acc[2] = f.z; acc[3] = f.w;
      } else {
        int4 v;
        if constexpr (std::is_same<T, __half>::value) v = ptx::multimem_ld_reduce_f16x8(a);
        else v = ptx::multimem_ld_reduce_bf16x8(a);
        const T* h = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[e] = to_f32(h[e]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < VN; ++e) acc[e] = 0.f;
      for (int r = 0; r < p.world; ++r) {
        const int4 v = ld_sys16(reinterpret_cast<const int4*>(p.peer_buf[r]) + base + i);
        const T* h = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[e] += to_f32(h[e]);
      }
    }
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < VN; ++e) o.v[e] = from_f32<T>(acc[e]);
    st16(out + i * VN, o);
  }
  coll_barrier(p, 1);  // nobody overwrites its input while a peer still reads it
  ptx::grid_dep_launch();
}

// src local [rows][world][vecs]; peer j's out region (at out_off) gets [rows][me][vecs]
__global__ void __launch_bounds__(512) all_to_all_kernel(const CollParams p, const int4* __restrict__ src, int64_t out_off,
                                                          int64_t rows, int64_t vecs) {
  ptx::grid_dep_wait();
  coll_barrier(p, 0);
  const int64_t total = rows * p.world * vecs;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t v = i % vecs, j = (i / vecs) % p.world, r = i / (vecs * p.world);
    int4* dst = reinterpret_cast<int4*>(p.peer_buf[j] + out_off) + (r * p.world + p.rank) * vecs + v;
    *dst = src[i];
  }
  coll_barrier(p, 1);
  ptx::grid_dep_launch();
}

int fill(CollParams& p, const int64_t* peer_buf, const int64_t* peer_sig, void* mc, void* epochs, int64_t rank, int64_t world,
         int64_t max_blocks) {
  FIB_CHECK(world >= 1 && world <= kMaxRanks, "collectives: world size must be in [1,16]");
  for (int i = 0; i < world; ++i) {
    p.peer_buf[i] = reinterpret_cast<uint8_t*>(peer_buf[i]);
    p.peer_sig[i] = reinterpret_cast<uint32_t*>(peer_sig[i]);
  }
  p.mc_buf = reinterpret_cast<uint8_t*>(mc);
  p.epochs = reinterpret_cast<uint32_t*>(epochs);
  p.rank = (int)rank;
  p.world = (int)world;
  p.max_blocks = (int)max_blocks;
  return 0;
}

inline int grid_for(int64_t vecs, int64_t max_blocks) {
  int64_t g = (vecs + 511) / 512;
  if (g > max_blocks) g = max_blocks;
  return g < 1 ? 1 : (int)g;
}

}  // namespace

extern "C" int nvls_all_gather(void* peer_buf_host, void* peer_sig_host, void* mc_buf, void* epochs, void* src, int64_t out_off,
                               int64_t shard_bytes, int64_t rank, int64_t world, int64_t max_blocks, int64_t pdl, int64_t stream_) {
  FIB_CHECK(shard_bytes % 16 == 0 && out_off % 16 == 0, "all_gather: sizes must be multiples of 16 bytes");
  CollParams p;
  if (fill(p, (const int64_t*)peer_buf_host, (const int64_t*)peer_sig_host, mc_buf, epochs, rank, world, max_blocks)) return 1;
  LaunchCfg lc(dim3(grid_for(shard_bytes / 16, max_blocks)), dim3(512), 0, reinterpret_cast<cudaStream_t>(stream_), pdl != 0);
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, all_gather_kernel, p, (const int4*)src, out_off, shard_bytes / 16));
  return 0;
}

extern "C" int nvls_reduce_scatter(void* peer_buf_host, void* peer_sig_host, void* mc_buf, void* epochs, void* out, int64_t in_off,
                                   int64_t shard_bytes, int64_t dtype, int64_t rank, int64_t world, int64_t max_blocks, int64_t pdl,
                                   int64_t stream_) {
  FIB_CHECK(shard_bytes % 16 == 0 && in_off % 16 == 0, "reduce_scatter: sizes must be multiples of 16 bytes");
  CollParams p;
  if (fill(p, (const int64_t*)peer_buf_host, (const int64_t*)peer_sig_host, mc_buf, epochs, rank, world, max_blocks)) return 1;
  LaunchCfg lc(dim3(grid_for(shard_bytes / 16, max_blocks)), dim3(512), 0, reinterpret_cast<cudaStream_t>(stream_), pdl != 0);
  if (dtype == kF32) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, reduce_scatter_kernel<float>, p, (float*)out, in_off, shard_bytes / 16));
  } else if (dtype == kF16) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, reduce_scatter_kernel<__half>, p, (__half*)out, in_off, shard_bytes / 16));
  } else if (dtype == kBF16) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, reduce_scatter_kernel<__nv_bfloat16>, p, (__nv_bfloat16*)out, in_off, shard_bytes / 16));
  } else {
    FIB_CHECK(false, "reduce_scatter: dtype must be f32/f16/bf16");
  }
  return 0;
}

extern "C" int p2p_all_to_all(void* peer_buf_host, void* peer_sig_host, void* epochs, void* src, int64_t out_off, int64_t rows,
                              int64_t row_bytes, int64_t rank, int64_t world, int64_t max_blocks, int64_t pdl, int64_t stream_) {
  FIB_CHECK(row_bytes % 16 == 0 && out_off % 16 == 0, "all_to_all: row size must be a multiple of 16 bytes");
  CollParams p;
  if (fill(p, (const int64_t*)peer_buf_host, (const int64_t*)peer_sig_host, nullptr, epochs, rank, world, max_blocks)) return 1;
  LaunchCfg lc(dim3(grid_for(rows * world * (row_bytes / 16), max_blocks)), dim3(512), 0, reinterpret_cast<cudaStream_t>(stream_), pdl != 0);
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, all_to_all_kernel, p, (const int4*)src, out_off, rows, row_bytes / 16));
  return 0;
}

// =====================================================================================================================
// Indexed all-to-all-v of the legacy MoE "prepare + comm" protocol (reference flashinfer/comm/trtllm_alltoall.py moe_comm,
// include/flashinfer/comm/trtllm_alltoall.cuh: moeAllToAllKernel).  Entry e of the send list goes to the rank j with
// send_cumsum[j - 1] <= e < send_cumsum[j] and carries row send_idx[e] of `src`; the k-th row received from rank i lands in row
// recv_idx[recv_cumsum[i - 1] + k] of `out`.  Two launches over a symmetric staging region [world][cap_rows][row]:
//   push: cross-rank barrier (every rank has left its previous pull), rows are stored straight into the target's staging slot
//         [me][k] with 16-byte peer stores over NVLink;
//   pull: cross-rank barrier (every rank's push kernel has completed), staging rows are scattered into `out`.
// Both grids have the same fixed size on every rank (the barriers are per CTA index); spins carry the watchdog.
// =====================================================================================================================
namespace {

struct MoeCommParams {
  CollParams c;
  const uint8_t* src;
  int64_t src_pitch;
  const int32_t* send_cumsum;
  const int32_t* send_idx;
  uint8_t* out;
  int64_t out_pitch;
  const int32_t* recv_cumsum;
  const int32_t* recv_idx;
  int vecs;
  int64_t cap_rows, src_rows, out_rows;
};

__device__ __forceinline__ void coll_barrier_wd(const CollParams& p, int phase) {
  __threadfence_system();
  __syncthreads();
  if (int(threadIdx.x) < p.world) {
    const int peer = threadIdx.x;
    const int slot_base = (phase * p.max_blocks + blockIdx.x) * p.world;
    const uint32_t epoch = p.epochs[phase * p.max_blocks + blockIdx.x] + 1;
    ptx::st_release_sys(p.peer_sig[peer] + slot_base + p.rank, epoch);
    ptx::spin_until_ge_sys(p.peer_sig[p.rank] + slot_base + peer, epoch);
  }
  __syncthreads();
  if (threadIdx.x == 0) p.epochs[phase * p.max_blocks + blockIdx.x] += 1;
}

// rank owning entry e of an inclusive cumulative count list, and the entry's offset inside that rank's run
__device__ __forceinline__ int owner_of(const int32_t* cumsum, int world, int e, int* k) {
  int j = 0, lo = 0;
  while (j < world - 1 && e >= cumsum[j]) {
    lo = cumsum[j];
    ++j;
  }
  *k = e - lo;
  return j;
}

__global__ void __launch_bounds__(512) moe_comm_push_kernel(const MoeCommParams p) {
  ptx::grid_dep_wait();
  coll_barrier_wd(p.c, 0);
  const int total = p.send_cumsum[p.c.world - 1];
  const int64_t work = int64_t(total) * p.vecs;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < work; i += int64_t(gridDim.x) * blockDim.x) {
    const int e = int(i / p.vecs), v = int(i - int64_t(e) * p.vecs);
    int k;
    const int j = owner_of(p.send_cumsum, p.c.world, e, &k);
    const int row = p.send_idx[e];
    if (k >= p.cap_rows || row < 0 || row >= p.src_rows) {
      if (v == 0) printf("fib200: moe_comm: rank %d entry %d (row %d, slot %d of peer %d) outside the workspace / input -> trap\n", p.c.rank, e, row, k, j);
      __trap();
    }
    const int4 val = *(reinterpret_cast<const int4*>(p.src + int64_t(row) * p.src_pitch) + v);
    *(reinterpret_cast<int4*>(p.c.peer_buf[j]) + (int64_t(p.c.rank) * p.cap_rows + k) * p.vecs + v) = val;
  }
  ptx::grid_dep_launch();
}

__global__ void __launch_bounds__(512) moe_comm_pull_kernel(const MoeCommParams p) {
  ptx::grid_dep_wait();
  coll_barrier_wd(p.c, 1);
  const int total = p.recv_cumsum[p.c.world - 1];
  const int64_t work = int64_t(total) * p.vecs;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < work; i += int64_t(gridDim.x) * blockDim.x) {
    const int e = int(i / p.vecs), v = int(i - int64_t(e) * p.vecs);
    int k;
    const int src_rank = owner_of(p.recv_cumsum, p.c.world, e, &k);
    const int row = p.recv_idx[e];
    if (row < 0 || row >= p.out_rows || k >= p.cap_rows) continue;
    const int4 val = ld_sys16(reinterpret_cast<const int4*>(p.c.peer_buf[p.c.rank]) + (int64_t(src_rank) * p.cap_rows + k) * p.vecs + v);
    *(reinterpret_cast<int4*>(p.out + int64_t(row) * p.out_pitch) + v) = val;
  }
  ptx::grid_dep_launch();
}

}  // namespace

extern "C" int moe_comm_run(void* peer_buf_host, void* peer_sig_host, void* epochs, void* src, int64_t src_rows, int64_t src_pitch,
                            void* send_cumsum, void* send_idx, void* out, int64_t out_rows, int64_t out_pitch, void* recv_cumsum,
                            void* recv_idx, int64_t row_bytes, int64_t cap_rows, int64_t rank, int64_t world, int64_t max_blocks,
                            int64_t stream_) {
  FIB_CHECK(row_bytes % 16 == 0 && src_pitch % 16 == 0 && out_pitch % 16 == 0, "moe_comm: rows / pitches must be multiples of 16 bytes");
  FIB_CHECK(cap_rows >= 1, "moe_comm: the workspace is too small for one row per rank pair");
  MoeCommParams p;
  memset(&p, 0, sizeof(p));
  if (fill(p.c, (const int64_t*)peer_buf_host, (const int64_t*)peer_sig_host, nullptr, epochs, rank, world, max_blocks)) return 1;
  p.src = reinterpret_cast<const uint8_t*>(src); p.src_pitch = src_pitch; p.src_rows = src_rows;
  p.send_cumsum = reinterpret_cast<const int32_t*>(send_cumsum); p.send_idx = reinterpret_cast<const int32_t*>(send_idx);
  p.out = reinterpret_cast<uint8_t*>(out); p.out_pitch = out_pitch; p.out_rows = out_rows;
  p.recv_cumsum = reinterpret_cast<const int32_t*>(recv_cumsum); p.recv_idx = reinterpret_cast<const int32_t*>(recv_idx);
  p.vecs = int(row_bytes / 16); p.cap_rows = cap_rows;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LaunchCfg lc(dim3((unsigned)max_blocks), dim3(512), 0, stream, false);
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, moe_comm_push_kernel, p));
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, moe_comm_pull_kernel, p));
  return 0;
}
