// In-kernel tensor-parallel all-reduce fused with residual-add + RMSNorm (+ optional fp8 quant) over
// NVLink 5 / NVSwitch peer memory.  No NCCL on this path.
//
// Parity: reference trtllm_allreduce_fusion (flashinfer/comm/trtllm_ar.py:951-1060, kernels
// include/flashinfer/comm/trtllm_allreduce_fusion.cuh:1336-1466), trtllm_mnnvl_allreduce
// (include/flashinfer/comm/trtllm_mnnvl_allreduce.cuh:514-890) and vllm custom all-reduce.
//
// B200-first design:
//  * inputs live in a symmetric heap (same offset on every rank) that is also mapped through an NVLS
//    multicast address: the N-way reduction happens INSIDE the NVSwitch via
//    `multimem.ld_reduce.add.acc::f32.v4.bf16x2` (one 16 B request returns the sum over all ranks);
//    the two-shot variant broadcasts results with `multimem.st`.  Without multicast the same kernels
//    fall back to direct P2P 16 B loads from every peer.
//  * one-shot (small token counts): every rank reduces every row -> a single cross-rank barrier, the
//    residual stream stays replicated, no result traffic at all.
//  * two-shot (large): rows are partitioned over ranks (whole rows, so the RMSNorm epilogue sees a
//    full hidden vector), reduced in-switch, normalised and multicast-stored to all ranks.
//  * cross-rank barriers are per-CTA epoch counters in the symmetric heap:
//    `st.release.sys` to the peer's slot, `ld.acquire.sys` spin on the local slot; epochs are kept in
//    device memory so the kernels are CUDA-graph replay safe and never need a reset.
//  * PDL: griddepcontrol.wait precedes the first signal (the producer GEMM may still be running).
#include <cuda_fp4.h>
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

constexpr int kMaxRanks = 16;

struct ARParams {
  const void* peer_in[kMaxRanks];   // input buffer of every rank (same symmetric offset)
  uint32_t* peer_sig[kMaxRanks];    // signal pad of every rank: [2][max_blocks][world] uint32
  void* peer_out[kMaxRanks];        // two-shot without multicast: output buffer of every rank
  const void* mc_in;                // multicast alias of the input buffer (or null)
  void* mc_out;                     // multicast alias of the symmetric output buffer (two-shot)
  void* out;                        // local output (normed if gamma else plain sum)
  void* residual;                   // local residual in/out (may be null)
  const void* gamma;                // RMSNorm weight (may be null -> plain all-reduce)
  uint32_t* epochs;                 // local device memory [2][max_blocks]
  void* quant_out;                  // optional fp8 output of the normed value
  int64_t tokens, hidden;
  int rank, world, max_blocks;
  float eps, weight_bias, quant_scale_inv;
};

template <typename T>
struct Pack16;
template <>
struct Pack16<__nv_bfloat16> {
  static __device__ __forceinline__ void load_reduce_mc(const void* mc, float* f) {
    int4 v = ptx::multimem_ld_reduce_bf16x8(mc);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
};
template <>
struct Pack16<__half> {
  static __device__ __forceinline__ void load_reduce_mc(const void* mc, float* f) {
    int4 v = ptx::multimem_ld_reduce_f16x8(mc);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
};
template <>
struct Pack16<float> {
  static __device__ __forceinline__ void load_reduce_mc(const void* mc, float* f) {
    float4 v = ptx::multimem_ld_reduce_f32x4(mc);
    f[0] = v.x;
    f[1] = v.y;
    f[2] = v.z;
    f[3] = v.w;
  }
};

// Per-CTA cross-rank barrier. `phase` selects one of two independent flag sets (start / end).
__device__ __forceinline__ void cta_rank_barrier(const ARParams& p, int phase, bool release_only_after_sync) {
  (void)release_only_after_sync;
  __syncthreads();
  if (threadIdx.x < p.world) {
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

template <typename T>
__device__ __forceinline__ float block_sum_ar(float v, float* smem) {
  v = warp_reduce_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = (lane < nwarps) ? smem[lane] : 0.f;
  return warp_reduce_sum(r);
}

// kTwoShot=false: every rank reduces all rows (single barrier).
// kTwoShot=true : rank r reduces rows r, r+world, ...; results are multicast-stored (or P2P-stored) to
//                 the symmetric output of every rank; second barrier before exit.
template <typename T, bool kNvls, bool kTwoShot, int kMaxVec>
__global__ void __launch_bounds__(1024) allreduce_fusion_kernel(const ARParams p) {
  constexpr int VN = 16 / sizeof(T);
  __shared__ float red[32];
  const int nvec = int(p.hidden / VN);
  const T* gamma = reinterpret_cast<const T*>(p.gamma);

  ptx::grid_dep_wait();
  cta_rank_barrier(p, 0, false);

  const int64_t row_step = kTwoShot ? int64_t(gridDim.x) * p.world : gridDim.x;
  const int64_t row0 = kTwoShot ? int64_t(blockIdx.x) * p.world + p.rank : blockIdx.x;
  for (int64_t row = row0; row < p.tokens; row += row_step) {
    float acc[kMaxVec][VN];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < kMaxVec; ++it) {
      const int v = threadIdx.x + it * blockDim.x;
      if (v < nvec) {
        const int64_t off = row * p.hidden + int64_t(v) * VN;
        if constexpr (kNvls) {
          Pack16<T>::load_reduce_mc(reinterpret_cast<const T*>(p.mc_in) + off, acc[it]);
        } else {
#pragma unroll
          for (int e = 0; e < VN; ++e) acc[it][e] = 0.f;
          for (int r = 0; r < p.world; ++r) {
            const int peer = (p.rank + r) % p.world;  // spread the first access over different links
            const int4 raw = ptx::ld_volatile_v4(reinterpret_cast<const T*>(p.peer_in[peer]) + off);
            const T* x = reinterpret_cast<const T*>(&raw);
#pragma unroll
            for (int e = 0; e < VN; ++e) acc[it][e] += to_f32(x[e]);
          }
        }
        if (p.residual) {
          T* res = reinterpret_cast<T*>(p.residual) + off;
          Vec16<T> rv = ld16(res);
#pragma unroll
          for (int e = 0; e < VN; ++e) {
            rv.v[e] = from_f32<T>(acc[it][e] + to_f32(rv.v[e]));
            acc[it][e] = to_f32(rv.v[e]);
          }
          st16(res, rv);
        }
#pragma unroll
        for (int e = 0; e < VN; ++e) ss += acc[it][e] * acc[it][e];
      }
    }
    float rstd = 1.f;
    if (gamma) {
      ss = block_sum_ar<T>(ss, red);
      rstd = rsqrtf(ss / float(p.hidden) + p.eps);
    }
#pragma unroll
    for (int it = 0; it < kMaxVec; ++it) {
      const int v = threadIdx.x + it * blockDim.x;
      if (v < nvec) {
        const int64_t off = row * p.hidden + int64_t(v) * VN;
        Vec16<T> o;
        if (gamma) {
          const Vec16<T> g = ldg16(gamma + int64_t(v) * VN);
#pragma unroll
          for (int e = 0; e < VN; ++e) o.v[e] = from_f32<T>(acc[it][e] * rstd * (to_f32(g.v[e]) + p.weight_bias));
        } else {
#pragma unroll
          for (int e = 0; e < VN; ++e) o.v[e] = from_f32<T>(acc[it][e]);
        }
        if constexpr (kTwoShot) {
          if constexpr (kNvls) {
            ptx::multimem_st_v4(reinterpret_cast<T*>(p.mc_out) + off, *reinterpret_cast<const int4*>(&o));
          } else {
            for (int r = 0; r < p.world; ++r)
              st16(reinterpret_cast<T*>(p.peer_out[(p.rank + r) % p.world]) + off, o);
          }
        } else {
          st16(reinterpret_cast<T*>(p.out) + off, o);
        }
        if (p.quant_out) {
          __nv_fp8_e4m3* q = reinterpret_cast<__nv_fp8_e4m3*>(p.quant_out) + off;
#pragma unroll
          for (int e = 0; e < VN; ++e)
            q[e] = __nv_fp8_e4m3(fminf(fmaxf(to_f32(o.v[e]) * p.quant_scale_inv, -448.f), 448.f));
        }
      }
    }
  }
  if constexpr (kTwoShot) {
    __threadfence_system();
    cta_rank_barrier(p, 1, true);
  }
  ptx::grid_dep_launch();
}

template <typename T>
int launch_ar(const ARParams& p, bool nvls, bool two_shot, int blocks, int threads, bool pdl, cudaStream_t stream) {
  LaunchCfg lc(dim3(blocks), dim3(threads), 0, stream, pdl);
#define FIB_AR(NVLS, TWO)                                                                              \
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, allreduce_fusion_kernel<T, NVLS, TWO, 4>, p))
  if (nvls && two_shot) {
    FIB_AR(true, true);
  } else if (nvls) {
    FIB_AR(true, false);
  } else if (two_shot) {
    FIB_AR(false, true);
  } else {
    FIB_AR(false, false);
  }
#undef FIB_AR
  return 0;
}

}  // namespace

// peer_in / peer_sig / peer_out: host arrays of `world` device pointers (int64 values).
extern "C" int allreduce_fusion_run(void* peer_in_host, void* peer_sig_host, void* peer_out_host, void* mc_in,
                                    void* mc_out, void* out, void* residual, void* gamma, void* epochs, void* quant_out,
                                    int64_t tokens, int64_t hidden, int64_t rank, int64_t world, int64_t max_blocks,
                                    double eps, double weight_bias, double quant_scale, int64_t two_shot, int64_t dtype,
                                    int64_t pdl, int64_t stream_) {
  FIB_CHECK(world >= 1 && world <= kMaxRanks, "allreduce: world size must be in [1,16]");
  const int esz = dtype_size(dtype);
  const int vn = 16 / esz;
  FIB_CHECK(hidden % vn == 0, "allreduce: hidden must be a multiple of the 16B vector width");
  if (tokens == 0) return 0;
  ARParams p;
  memset(&p, 0, sizeof(p));
  const int64_t* pin = reinterpret_cast<const int64_t*>(peer_in_host);
  const int64_t* psig = reinterpret_cast<const int64_t*>(peer_sig_host);
  const int64_t* pout = reinterpret_cast<const int64_t*>(peer_out_host);
  for (int i = 0; i < world; ++i) {
    p.peer_in[i] = reinterpret_cast<const void*>(pin[i]);
    p.peer_sig[i] = reinterpret_cast<uint32_t*>(psig[i]);
    p.peer_out[i] = pout ? reinterpret_cast<void*>(pout[i]) : nullptr;
  }
  p.mc_in = mc_in;
  p.mc_out = mc_out;
  p.out = out;
  p.residual = residual;
  p.gamma = gamma;
  p.epochs = reinterpret_cast<uint32_t*>(epochs);
  p.quant_out = quant_out;
  p.tokens = tokens;
  p.hidden = hidden;
  p.rank = (int)rank;
  p.world = (int)world;
  p.max_blocks = (int)max_blocks;
  p.eps = (float)eps;
  p.weight_bias = (float)weight_bias;
  p.quant_scale_inv = quant_scale != 0.0 ? (float)(1.0 / quant_scale) : 1.f;
  const int nvec = int(hidden / vn);
  int threads = (nvec + 31) / 32 * 32;
  if (threads > 1024) threads = 1024;
  FIB_CHECK(nvec <= threads * 4, "allreduce: hidden too large for this kernel (max 4 vectors per thread)");
  if (threads < world) threads = (int(world) + 31) / 32 * 32;
  int64_t rows_mine = two_shot ? (tokens + world - 1) / world : tokens;
  int blocks = (int)(rows_mine < max_blocks ? rows_mine : max_blocks);
  if (blocks < 1) blocks = 1;
  const bool nvls = mc_in != nullptr && (!two_shot || mc_out != nullptr);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (dtype == kBF16) return launch_ar<__nv_bfloat16>(p, nvls, two_shot != 0, blocks, threads, pdl != 0, stream);
  if (dtype == kF16) return launch_ar<__half>(p, nvls, two_shot != 0, blocks, threads, pdl != 0, stream);
  if (dtype == kF32) return launch_ar<float>(p, nvls, two_shot != 0, blocks, threads, pdl != 0, stream);
  return set_error("allreduce: unsupported dtype");
}

// =====================================================================================================================
// One-shot PUSH all-reduce (Lamport style) with fused prologue and epilogue - the small-message path (<= 128-256 tokens).
//
// Parity: reference include/flashinfer/comm/trtllm_allreduce_fusion.cuh:1336-1403 (one-shot Lamport kernel),
// :487-740 (fp8 / nvfp4 quantisation epilogues with linear / 128x4 / 8x4 scale layouts) and
// include/flashinfer/comm/trtllm_moe_allreduce_fusion.cuh:938,1239 (MoE reduction / finalize fused in front of the all-reduce).
//
// One CTA per token.  Prologue (registers only): plain input row | MoE reduction  sum_e scale[e, t] * act[e, t, :] + token[t, :]
// | MoE finalize  sum_k w[t, k] * permuted[e2p[t, k], :] (+ shared[t, :]).  The bf16 / fp16 row is multicast with
// `multimem.st` into slot [rank] of EVERY rank's receive buffer (per-peer stores without NVLS); -0.0 is the "not yet
// arrived" sentinel, the sender flushes real -0.0 to +0.0.  Every rank then polls its own copy of all slots, sums them in rank
// order (bitwise identical everywhere), and runs the epilogue on the fp32 row: raw sum out | + residual -> residual out |
// RMSNorm -> norm out | e4m3 quantisation | NVFP4 quantisation with the scale factors written in the requested layout.
// No flags, no system fences.  Three rotating buffers, selected by a device-side epoch (CUDA-graph replay safe); the buffer
// of the NEXT call is reset to the sentinel while the pushes are in flight.  Polls carry the 20 s watchdog.
// =====================================================================================================================
namespace {

struct PushParams {
  // prologue
  int mode;                       // 0 plain | 1 MoE reduction | 2 MoE finalize
  const void* in;                 // mode 0: [T, H];  mode 1: active-expert outputs [E, T, H];  mode 2: permuted rows [P, H]
  const float* moe_scale;         // mode 1: [E, T];  mode 2: expert weights [T, K] (null = 1)
  const void* moe_token_in;       // mode 1: [T, H] added to the reduction;  mode 2: shared-expert output [T, H] (or null)
  const int32_t* e2p;             // mode 2: expanded (t * K + k) -> permuted row
  int moe_n;                      // mode 1: experts E;  mode 2: top-k K
  // exchange
  void* recv;                     // local [3][world][max_tokens][hidden]
  void* mc_recv;                  // multicast alias (or null)
  void* peer_recv[kMaxRanks];
  uint32_t* epoch;                // local device word
  int64_t slot_elems, buf_elems;
  int rank, world;
  int tokens, hidden, max_tokens;
  // epilogue
  void* ar_out;                   // raw all-reduce sum [T, H] (or null)
  const void* residual_in;        // (or null)
  void* residual_out;             // (or null)
  const void* gamma;              // RMSNorm weight (or null: no norm)
  void* norm_out;                 // (or null)
  float eps, weight_bias;
  int quant;                      // 0 none | 1 e4m3 (quant_out [T, H]) | 2 nvfp4 (quant_out [T, H/2] + scale_out)
  void* quant_out;
  uint8_t* scale_out;
  const float* scale_factor;      // device scalar: fp8: y / scale;  nvfp4: global scale (448 * 6 / amax)
  int sf_layout;                  // 0 = 128x4 swizzled, 1 = 8x4 swizzled, 2 = linear
};

__device__ __forceinline__ int64_t sf_offset(int layout, int row, int c, int ncols_sf) {
  const int pad4 = (ncols_sf + 3) / 4 * 4;
  if (layout == 2) return int64_t(row) * ncols_sf + c;
  if (layout == 1) return (int64_t(row / 8) * (pad4 / 4) + c / 4) * 32 + (row % 8) * 4 + (c % 4);
  return (int64_t(row / 128) * (pad4 / 4) + c / 4) * 512 + (row % 32) * 16 + ((row % 128) / 32) * 4 + (c % 4);
}

template <typename T, int kVecPerThread>
__global__ void __launch_bounds__(1024) allreduce_push_kernel(const PushParams p) {
  constexpr int VN = 8;  // 16-bit elements per 16-byte vector
  __shared__ float red[32];
  __shared__ uint32_t s_epoch;
  const int row = blockIdx.x;
  const int nvec = p.hidden / VN;
  ptx::grid_dep_wait();
  __shared__ int s_dirty;
  if (threadIdx.x == 0) {
    s_epoch = *reinterpret_cast<volatile uint32_t*>(p.epoch);
    s_dirty = int(*reinterpret_cast<volatile uint32_t*>(p.epoch + 1 + (s_epoch + 1u) % 3u));  // rows the NEXT buffer still holds
  }
  __syncthreads();
  const uint32_t ep = s_epoch;
  const int dirty_next = s_dirty;
  T* recv = reinterpret_cast<T*>(p.recv);
  const int64_t cur = int64_t(ep % 3u) * p.buf_elems, nxt = int64_t((ep + 1u) % 3u) * p.buf_elems;

  // ---------------- prologue + push
  int4 mine[kVecPerThread];
#pragma unroll
  for (int it = 0; it < kVecPerThread; ++it) {
    const int v = threadIdx.x + it * blockDim.x;
    if (v >= nvec) continue;
    const int64_t off = int64_t(row) * p.hidden + int64_t(v) * VN;
    float x[VN];
    if (p.mode == 0) {
      const Vec16<T> a = ld16(reinterpret_cast<const T*>(p.in) + off);
#pragma unroll
      for (int e = 0; e < VN; ++e) x[e] = to_f32(a.v[e]);
    } else if (p.mode == 1) {
      const Vec16<T> t0 = ld16(reinterpret_cast<const T*>(p.moe_token_in) + off);
#pragma unroll
      for (int e = 0; e < VN; ++e) x[e] = to_f32(t0.v[e]);
      for (int ex = 0; ex < p.moe_n; ++ex) {
        const float sc = p.moe_scale[int64_t(ex) * p.tokens + row];
        const Vec16<T> a = ld16(reinterpret_cast<const T*>(p.in) + int64_t(ex) * p.tokens * p.hidden + off);
#pragma unroll
        for (int e = 0; e < VN; ++e) x[e] += sc * to_f32(a.v[e]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < VN; ++e) x[e] = 0.f;
      if (p.moe_token_in) {
        const Vec16<T> t0 = ld16(reinterpret_cast<const T*>(p.moe_token_in) + off);
#pragma unroll
        for (int e = 0; e < VN; ++e) x[e] = to_f32(t0.v[e]);
      }
      for (int k = 0; k < p.moe_n; ++k) {
        const int pr = p.e2p[int64_t(row) * p.moe_n + k];
        if (pr < 0) continue;
        const float w = p.moe_scale ? p.moe_scale[int64_t(row) * p.moe_n + k] : 1.f;
        const Vec16<T> a = ld16(reinterpret_cast<const T*>(p.in) + int64_t(pr) * p.hidden + int64_t(v) * VN);
#pragma unroll
        for (int e = 0; e < VN; ++e) x[e] += w * to_f32(a.v[e]);
      }
    }
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < VN; ++e) o.v[e] = from_f32<T>(x[e]);
    uint32_t* w32 = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {  // -0.0 is the sentinel: never send it
      if ((w32[e] & 0xffffu) == 0x8000u) w32[e] &= 0xffff0000u;
      if ((w32[e] >> 16) == 0x8000u) w32[e] &= 0x0000ffffu;
    }
    mine[it] = *reinterpret_cast<const int4*>(&o);
    const int64_t dst = cur + int64_t(p.rank) * p.slot_elems + off;
    if (p.mc_recv) {
      ptx::multimem_st_v4(reinterpret_cast<T*>(p.mc_recv) + dst, mine[it]);
    } else {
      for (int r = 0; r < p.world; ++r) ptx::st_na_v4(reinterpret_cast<T*>(p.peer_recv[(p.rank + r) % p.world]) + dst, mine[it]);
    }
  }
  // ---------------- while the pushes fly: reset the rows the NEXT call's buffer still holds from its last use (two calls ago);
  //                  the per-buffer dirty row count lives next to the epoch, so the work is what was actually written
  {
    const int4 sent = make_int4(int(0x80008000u), int(0x80008000u), int(0x80008000u), int(0x80008000u));
    for (int rr = row; rr < dirty_next; rr += gridDim.x)
      for (int r = 0; r < p.world; ++r)
        for (int v = threadIdx.x; v < nvec; v += blockDim.x)
          *reinterpret_cast<int4*>(recv + nxt + int64_t(r) * p.slot_elems + int64_t(rr) * p.hidden + int64_t(v) * VN) = sent;
  }
  // ---------------- gather + reduce (rank order) + epilogue
  float acc[kVecPerThread][VN];
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < kVecPerThread; ++it) {
    const int v = threadIdx.x + it * blockDim.x;
    if (v >= nvec) continue;
    const int64_t off = int64_t(row) * p.hidden + int64_t(v) * VN;
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[it][e] = 0.f;
    for (int r = 0; r < p.world; ++r) {
      int4 x;
      if (r == p.rank) {
        x = mine[it];  // my own contribution is already in registers (its multicast copy is never waited for)
      } else {
        const T* src = recv + cur + int64_t(r) * p.slot_elems + off;
        uint32_t polls = 0;
        uint64_t t0 = 0;
        bool done;
        do {
          x = ptx::ld_volatile_v4(src);
          const uint32_t* w = reinterpret_cast<const uint32_t*>(&x);
          done = true;
#pragma unroll
          for (int e = 0; e < 4; ++e) done = done && ((w[e] & 0xffffu) != 0x8000u) && ((w[e] >> 16) != 0x8000u);
          if (!done && (++polls & 0x3ffu) == 0) {
            if (t0 == 0) t0 = ptx::globaltimer();
            else if (ptx::globaltimer() - t0 > ptx::kSpinTimeoutNs) {
              printf("fib200: allreduce_push watchdog: rank %d row %d waited 20 s for rank %d -> trap\n", p.rank, row, r);
              __trap();
            }
          }
        } while (!done);
      }
      const T* h = reinterpret_cast<const T*>(&x);
#pragma unroll
      for (int e = 0; e < VN; ++e) acc[it][e] += to_f32(h[e]);
    }
    if (p.ar_out) {
      Vec16<T> o;
#pragma unroll
      for (int e = 0; e < VN; ++e) o.v[e] = from_f32<T>(acc[it][e]);
      st16(reinterpret_cast<T*>(p.ar_out) + off, o);
    }
    if (p.residual_in) {
      Vec16<T> rv = ld16(reinterpret_cast<const T*>(p.residual_in) + off);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        rv.v[e] = from_f32<T>(acc[it][e] + to_f32(rv.v[e]));
        acc[it][e] = to_f32(rv.v[e]);
      }
      if (p.residual_out) st16(reinterpret_cast<T*>(p.residual_out) + off, rv);
    }
#pragma unroll
    for (int e = 0; e < VN; ++e) ss += acc[it][e] * acc[it][e];
  }
  float rstd = 1.f;
  if (p.gamma) {
    ss = block_sum_ar<T>(ss, red);
    rstd = rsqrtf(ss / float(p.hidden) + p.eps);
  }
  const float qs = p.scale_factor ? *p.scale_factor : 1.f;
#pragma unroll
  for (int it = 0; it < kVecPerThread; ++it) {
    const int v = threadIdx.x + it * blockDim.x;
    const bool live = v < nvec;
    const int64_t off = int64_t(row) * p.hidden + int64_t(live ? v : 0) * VN;
    float y[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) y[e] = 0.f;
    if (live) {
      if (p.gamma) {
        const Vec16<T> g = ldg16(reinterpret_cast<const T*>(p.gamma) + int64_t(v) * VN);
#pragma unroll
        for (int e = 0; e < VN; ++e) y[e] = to_f32(from_f32<T>(acc[it][e] * rstd * (to_f32(g.v[e]) + p.weight_bias)));
      } else {
#pragma unroll
        for (int e = 0; e < VN; ++e) y[e] = acc[it][e];
      }
      if (p.norm_out) {
        Vec16<T> o;
#pragma unroll
        for (int e = 0; e < VN; ++e) o.v[e] = from_f32<T>(y[e]);
        st16(reinterpret_cast<T*>(p.norm_out) + off, o);
      }
      if (p.quant == 1) {
        uint8_t q8[VN];
        const float inv = 1.f / qs;
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const __nv_fp8_e4m3 q(fminf(fmaxf(y[e] * inv, -448.f), 448.f));
          q8[e] = *reinterpret_cast<const uint8_t*>(&q);
        }
        *reinterpret_cast<int2*>(reinterpret_cast<uint8_t*>(p.quant_out) + off) = *reinterpret_cast<const int2*>(q8);
      }
    }
    if (p.quant == 2) {
      // NVFP4: 16-element blocks = two neighbouring vectors (lanes 2j, 2j + 1); every lane takes part in the shuffle
      float amax = 0.f;
#pragma unroll
      for (int e = 0; e < VN; ++e) amax = fmaxf(amax, fabsf(y[e]));
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
      const __nv_fp8_e4m3 s8(qs * (amax * (1.f / 6.f)));
      const float sfv = float(s8);
      const float out_scale = sfv != 0.f ? qs / sfv : 0.f;
      if (live) {
        uint8_t pk[4];
#pragma unroll
        for (int e = 0; e < VN; e += 2)
          pk[e / 2] = (uint8_t)__nv_cvt_float2_to_fp4x2(make_float2(y[e] * out_scale, y[e + 1] * out_scale), __NV_E2M1, cudaRoundNearest);
        *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(p.quant_out) + (int64_t(row) * p.hidden + int64_t(v) * VN) / 2) =
            *reinterpret_cast<const uint32_t*>(pk);
        if ((v & 1) == 0 && p.scale_out)
          p.scale_out[sf_offset(p.sf_layout, row, v / 2, p.hidden / 16)] = *reinterpret_cast<const uint8_t*>(&s8);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    volatile uint32_t* e = reinterpret_cast<volatile uint32_t*>(p.epoch);
    // rows [0, tokens) of `cur` are dirty now; with fewer CTAs than dirty rows of `nxt` the strided loop above covered them all
    e[1 + ep % 3u] = uint32_t(p.tokens);
    e[1 + (ep + 1u) % 3u] = 0u;
    e[0] = ep + 1u;
  }
  ptx::grid_dep_launch();
}

}  // namespace

extern "C" int allreduce_push_run(int64_t mode, void* in, void* moe_scale, void* moe_token_in, void* e2p, int64_t moe_n, void* recv,
                                  void* mc_recv, void* peer_recv_host, void* epoch, int64_t slot_elems, int64_t rank, int64_t world,
                                  int64_t tokens, int64_t hidden, int64_t max_tokens, void* ar_out, void* residual_in,
                                  void* residual_out, void* gamma, void* norm_out, double eps, double weight_bias, int64_t quant,
                                  void* quant_out, void* scale_out, void* scale_factor, int64_t sf_layout, int64_t dtype, int64_t pdl,
                                  int64_t stream_) {
  FIB_CHECK(world >= 1 && world <= kMaxRanks, "allreduce_push: world size must be in [1, 16]");
  FIB_CHECK(dtype == kF16 || dtype == kBF16, "allreduce_push: f16 / bf16 only");
  FIB_CHECK(hidden % 16 == 0 && hidden <= 16384, "allreduce_push: hidden must be a multiple of 16 and <= 16384");
  FIB_CHECK(tokens >= 1 && tokens <= max_tokens && slot_elems >= max_tokens * hidden, "allreduce_push: token count exceeds the workspace");
  FIB_CHECK(mode >= 0 && mode <= 2 && quant >= 0 && quant <= 2 && sf_layout >= 0 && sf_layout <= 2, "allreduce_push: bad mode / quant / layout");
  FIB_CHECK(recv && epoch && (mc_recv || peer_recv_host || world == 1), "allreduce_push: receive buffers required");
  if (mode == 1) FIB_CHECK(moe_scale && moe_token_in && moe_n >= 1, "allreduce_push (MoE reduction): scale / token input required");
  if (mode == 2) FIB_CHECK(e2p && moe_n >= 1, "allreduce_push (MoE finalize): expanded_idx_to_permuted_idx required");
  if (quant) FIB_CHECK(quant_out != nullptr, "allreduce_push: quant_out required");
  PushParams p;
  memset(&p, 0, sizeof(p));
  p.mode = int(mode); p.in = in; p.moe_scale = reinterpret_cast<const float*>(moe_scale); p.moe_token_in = moe_token_in;
  p.e2p = reinterpret_cast<const int32_t*>(e2p); p.moe_n = int(moe_n);
  p.recv = recv; p.mc_recv = mc_recv; p.epoch = reinterpret_cast<uint32_t*>(epoch);
  if (peer_recv_host) {
    const int64_t* ps = reinterpret_cast<const int64_t*>(peer_recv_host);
    for (int i = 0; i < world; ++i) p.peer_recv[i] = reinterpret_cast<void*>(ps[i]);
  }
  p.slot_elems = slot_elems; p.buf_elems = slot_elems * world; p.rank = int(rank); p.world = int(world);
  p.tokens = int(tokens); p.hidden = int(hidden); p.max_tokens = int(max_tokens);
  p.ar_out = ar_out; p.residual_in = residual_in; p.residual_out = residual_out; p.gamma = gamma; p.norm_out = norm_out;
  p.eps = float(eps); p.weight_bias = float(weight_bias); p.quant = int(quant); p.quant_out = quant_out;
  p.scale_out = reinterpret_cast<uint8_t*>(scale_out); p.scale_factor = reinterpret_cast<const float*>(scale_factor);
  p.sf_layout = int(sf_layout);
  const int nvec = int(hidden / 8);
  int threads = (nvec + 31) / 32 * 32;
  int vpt = 1;
  if (threads > 1024) {
    vpt = 2;
    threads = ((nvec + 1) / 2 + 31) / 32 * 32;
  }
  if (threads < 32) threads = 32;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LaunchCfg lc(dim3((unsigned)tokens), dim3(threads), 0, stream, pdl != 0);
#define FIB_PUSH(TT, V) FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, allreduce_push_kernel<TT, V>, p))
  if (dtype == kBF16) {
    if (vpt == 1) { FIB_PUSH(__nv_bfloat16, 1); } else { FIB_PUSH(__nv_bfloat16, 2); }
  } else {
    if (vpt == 1) { FIB_PUSH(__half, 1); } else { FIB_PUSH(__half, 2); }
  }
#undef FIB_PUSH
  return 0;
}

// =====================================================================================================================
// Greedy sampling over a vocabulary-sharded LM head in ONE kernel: per-row argmax of the local logits shard, (value, global
// index) pushed to every rank (8-byte {payload, tag} stores, LL style: the tag is the call epoch, so there is nothing to reset),
// poll the `world` pairs of the row, pick the winner (largest value, lowest index on ties: identical on all ranks).
// Replaces torch.max + a small all-reduce + a barrier kernel + gather glue at the end of a tensor-parallel decode step.
// Parity: the reference leaves this to the serving engine (vLLM / SGLang gather the vocab-parallel logits with NCCL).
// =====================================================================================================================
namespace {

struct ArgmaxParams {
  const void* logits;        // [rows, ld] local shard
  int64_t ld;
  int shard;                 // valid columns of the shard
  int64_t index_offset;      // global index of column 0
  uint2* inbox;              // local [2 parity][world][max_rows][2]
  uint2* peer_inbox[kMaxRanks];
  uint32_t* epoch;           // local device word
  int rank, world, rows, max_rows;
  int64_t* out;              // [rows] winning global index
  float* out_val;            // [rows] winning value (or null)
};

__device__ __forceinline__ void st_volatile_v2(uint2* p, uint2 v) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint2 ld_volatile_v2(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(1024) argmax_push_kernel(const ArgmaxParams p) {
  constexpr int VN = 16 / sizeof(T);
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  __shared__ float r_val[kMaxRanks];
  __shared__ int64_t r_idx[kMaxRanks];
  __shared__ uint32_t s_epoch;
  const int row = blockIdx.x;
  ptx::grid_dep_wait();
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(p.epoch);
  const T* src = reinterpret_cast<const T*>(p.logits) + int64_t(row) * p.ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  const int nfull = vec_ok ? p.shard / VN * VN : 0;
  for (int c = threadIdx.x * VN; c < nfull; c += blockDim.x * VN) {
    const Vec16<T> v = ldg16(src + c);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      const float x = to_f32(v.v[e]);
      if (x > best) {  // ascending scan: the first occurrence of the maximum wins
        best = x;
        bi = c + e;
      }
    }
  }
  for (int c = nfull + threadIdx.x; c < p.shard; c += blockDim.x) {
    const float x = to_f32(src[c]);
    if (x > best || (x == best && c < bi)) {
      best = x;
      bi = c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    s_val[threadIdx.x >> 5] = best;
    s_idx[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  const uint32_t tag = s_epoch + 1u;
  const int par = int(s_epoch & 1u);
  if (threadIdx.x < p.world) {
    float v = s_val[0];
    int i = s_idx[0];
    for (int w = 1; w < int(blockDim.x >> 5); ++w)
      if (s_val[w] > v || (s_val[w] == v && s_idx[w] < i)) {
        v = s_val[w];
        i = s_idx[w];
      }
    const int64_t gi = p.index_offset + (i == 0x7fffffff ? 0 : i);
    // push my pair into slot [rank] of peer `threadIdx.x`, then wait for that peer's pair in my inbox
    const int peer = threadIdx.x;
    uint2* dst = p.peer_inbox[peer] + ((int64_t(par) * p.world + p.rank) * p.max_rows + row) * 2;
    st_volatile_v2(dst, make_uint2(__float_as_uint(v), tag));
    st_volatile_v2(dst + 1, make_uint2(uint32_t(gi), tag));
    const uint2* in = p.inbox + ((int64_t(par) * p.world + peer) * p.max_rows + row) * 2;
    uint2 a, b;
    uint32_t polls = 0;
    uint64_t t0 = 0;
    for (;;) {
      a = ld_volatile_v2(in);
      b = ld_volatile_v2(in + 1);
      if (a.y == tag && b.y == tag) break;
      if ((++polls & 0x3ffu) == 0) {
        if (t0 == 0) t0 = ptx::globaltimer();
        else if (ptx::globaltimer() - t0 > ptx::kSpinTimeoutNs) {
          printf("fib200: argmax_push watchdog: rank %d row %d waited 20 s for rank %d -> trap\n", p.rank, row, peer);
          __trap();
        }
      }
    }
    r_val[peer] = __uint_as_float(a.x);
    r_idx[peer] = int64_t(b.x);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = r_val[0];
    int64_t i = r_idx[0];
    for (int r = 1; r < p.world; ++r)
      if (r_val[r] > v || (r_val[r] == v && r_idx[r] < i)) {
        v = r_val[r];
        i = r_idx[r];
      }
    p.out[row] = i;
    if (p.out_val) p.out_val[row] = v;
    if (blockIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(p.epoch) = tag;  // every block read the epoch before its scan
  }
  ptx::grid_dep_launch();
}

}  // namespace

extern "C" int argmax_push_run(void* logits, int64_t ld, int64_t shard, int64_t index_offset, void* inbox, void* peer_inbox_host,
                               void* epoch, int64_t rank, int64_t world, int64_t rows, int64_t max_rows, void* out, void* out_val,
                               int64_t dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(world >= 1 && world <= kMaxRanks, "argmax_push: world size must be in [1, 16]");
  FIB_CHECK(rows >= 1 && rows <= max_rows && shard >= 1 && shard < (int64_t(1) << 31) && index_offset + shard < (int64_t(1) << 32),
            "argmax_push: bad shape (rows <= max_rows, global vocabulary index must fit 32 bits)");
  FIB_CHECK(inbox && peer_inbox_host && epoch && out, "argmax_push: inbox / peer table / epoch / out required");
  ArgmaxParams p;
  memset(&p, 0, sizeof(p));
  p.logits = logits; p.ld = ld; p.shard = int(shard); p.index_offset = index_offset;
  p.inbox = reinterpret_cast<uint2*>(inbox);
  const int64_t* ps = reinterpret_cast<const int64_t*>(peer_inbox_host);
  for (int i = 0; i < world; ++i) p.peer_inbox[i] = reinterpret_cast<uint2*>(ps[i]);
  p.epoch = reinterpret_cast<uint32_t*>(epoch);
  p.rank = int(rank); p.world = int(world); p.rows = int(rows); p.max_rows = int(max_rows);
  p.out = reinterpret_cast<int64_t*>(out); p.out_val = reinterpret_cast<float*>(out_val);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  // few long rows (a decode batch over a 128 K vocabulary shard): 1024 threads per row keep enough loads in flight
  LaunchCfg lc(dim3((unsigned)rows), dim3(shard >= 16384 ? 1024 : 256), 0, stream, pdl != 0);
  if (dtype == kBF16) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, argmax_push_kernel<__nv_bfloat16>, p));
  } else if (dtype == kF16) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, argmax_push_kernel<__half>, p));
  } else if (dtype == kF32) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, argmax_push_kernel<float>, p));
  } else {
    return set_error("argmax_push: f16 / bf16 / f32 logits only");
  }
  return 0;
}
