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
