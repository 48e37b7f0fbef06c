// Decode-layer linear for sm_100a:  out = epilogue( A[M <= 64, K] * W[N, K]^T ), one launch per projection of a decode step.
//
// This is the small-M ("weight streaming") GEMM of the flagship decode step with everything that used to sit between two
// GEMMs folded into its prologue / epilogue, so that a transformer layer is 5 launches (QKV, attention, O, gate/up, down):
//   * RMSNorm is algebraically folded: the norm weight is multiplied into W offline and the per-token 1/rms is applied to the
//     fp32 accumulators (`row_sumsq`), the sum of squares itself is produced by the epilogue of the PREVIOUS residual GEMM;
//   * kRope:  QKV projection -> RoPE on the accumulators (rotation pairs are adjacent columns: weight rows are permuted at load
//     time) -> Q to the query buffer, K / V straight into their paged-KV-cache slots (no rope / append kernels);
//   * kGated: gate/up projection -> SwiGLU on the accumulators (row-interleaved gate / up weights);
//   * kResid: O / down projection -> (tensor-parallel: in-kernel all-reduce over NVLink, see below) -> residual += sum,
//     per-token sum of squares of the new residual accumulated for the next folded RMSNorm.
// Tensor parallelism (world > 1, kResid): the all-reduce is a one-shot PUSH inside the epilogue, Lamport style.  Every rank
// owns `world` receive slots per rotating buffer in a symmetric heap; the epilogue multicasts its bf16 partial tile straight from
// registers into slot [rank] of EVERY rank with `multimem.st` through the NVSwitch (per-peer stores when the fabric has no
// multicast), then polls its own columns of all slots until no sentinel (-0.0, which the sender never emits) is left, sums them
// in rank order (bitwise identical on all ranks: the replicated residual stream cannot drift), adds the residual and
// accumulates the norm statistics.  No flags, no system-scope fences, no staging round trip: the data is its own arrival signal,
// the cost over the local GEMM is one NVLink one-way latency plus world x tile bytes of ingress.  Three rotating buffers selected
// by a device-side epoch (CUDA-graph replay safe); the buffer of the next call is reset to the sentinel here (it was last read
// two calls ago).  Polls carry a watchdog.  GEMM, all-reduce, residual add and norm statistics are ONE kernel; no NCCL call or
// separate all-reduce launch exists on the path.
//
// Parity targets: reference include/flashinfer/gemm/tgv_gemm.cuh (low-latency small-M GEMM, PDL), the fused
// allreduce + residual + RMSNorm patterns of include/flashinfer/comm/trtllm_allreduce_fusion.cuh:1336-1466, the
// gemm + all-reduce of flashinfer/cute_dsl/gemm_allreduce_two_shot.py and rope_quantize_fp8_append_paged_kv_cache
// (flashinfer/rope.py:1500-1691).
//
// Kernel shape: activations are the M = 64 UMMA A operand, a tile of BN weight rows the B operand (`tcgen05.mma
// cta_group::1 kind::f16`, fp32 accumulators in TMEM), one output tile per cluster, split-K over a 2-CTA cluster.  The K
// halves are combined SYMMETRICALLY: each CTA owns half of the tile's columns and pushes the other half of its partial
// into the peer's shared memory with `st.async` that credits the peer's mbarrier (tx bytes) - no hand-shake, no release
// fence, no exit cluster barrier (a CTA leaves once the bytes addressed to it have landed).  Weights are prefetched into
// the TMA ring before `griddepcontrol.wait` (they do not depend on the previous kernel).
#include <stdlib.h>
#include <math.h>
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>
#include <fib200/profiler.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

constexpr int BM = 64, BK = 64, kMaxRanks = 16;
constexpr int kABytes = BM * BK * 2;  // 8 KB activation tile per stage

enum Epi : int { kPlain = 0, kGated = 1, kResid = 2, kRope = 3 };
// intra-kernel profiler: groups = warp roles (0 TMA producer, 1 MMA issuer, 2 epilogue), events below
enum ProfEvent : int { kEvSetup = 0, kEvWaitPrevGrid = 1, kEvWeightPrefetch = 2, kEvMainLoop = 3, kEvSplitKExchange = 4, kEvEpilogue = 5, kEvAllReduce = 6 };

struct DLP {
  int M, N, K, BN, S, kblocks, stages, epi;
  uint64_t* prof;  // intra-kernel profiler buffer (FIB200_ENABLE_PROFILER builds; see fib200/profiler.cuh)
  int prof_events;
  int w_blockk;  // weights stored BlockMajorK: [K / 64, N, 64] (every TMA box is one contiguous BN x 128 B chunk)
  void* out;
  int64_t ldo;
  const void* bias;
  const float* row_sumsq;  // folded RMSNorm: rstd[m] = rsqrt(row_sumsq[m] * inv_dim + eps)
  float inv_dim, eps;
  void* resid;             // kResid: residual stream in / out
  int64_t ldr;
  float* sumsq_out;        // kResid: += sum_n resid_new[m, n]^2
  int world, rank;         // kResid all-reduce (Lamport push)
  int ar_algo;             // 1 one-shot (gather all partials), 2 two-shot (reduce-scatter push + multicast all-gather)
  void* recv;              // local receive buffers [3][world][rows][lds] (symmetric heap)
  int64_t lds;             // row pitch of a slot (elements)
  int64_t slot_elems;      // elements per rank slot
  int64_t buf_elems;       // elements per rotating buffer (= world * slot_elems)
  void* mc_recv;           // multicast alias of `recv` (or null)
  uint32_t* epoch;         // local device word: number of all-reduce calls so far (selects the rotating buffer)
  void* peer_recv[kMaxRanks];
  const float* cos_sin;      // kRope: [M, head_dim] fp32 = cos[0:hd/2] | sin[0:hd/2] of the token's position
  const int64_t* cache_row;  // kRope: [M] element offset of the token's (page, slot) row in k_cache / v_cache
  void* k_cache;
  void* v_cache;
  int64_t c_sh;              // head stride inside a cache row (elements)
  int hq, hkv, head_dim, interleave;
};

__host__ __device__ inline int dl_stage_bytes(int BN) { return kABytes + BN * BK * 2; }
__host__ __device__ inline int dl_xbuf_bytes(int BN, int S) { return S > 1 ? (S - 1) * (BN / S) * BM * 4 : 0; }

template <typename T>
__device__ __forceinline__ void store16(T* dst, const float* v) {  // 16 values -> two 16-byte stores
  Vec16<T> a, b;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    a.v[e] = from_f32<T>(v[e]);
    b.v[e] = from_f32<T>(v[8 + e]);
  }
  st16(dst, a);
  st16(dst + 8, b);
}

template <typename T>
__global__ void __launch_bounds__(256, 1)
dlinear_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DLP p, uint32_t idesc) {
  const int BN = p.BN, S = p.S, kStages = p.stages;
  const int stage_bytes = dl_stage_bytes(BN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xbuf = smem + kStages * stage_bytes;  // split-K: the peer's partial of MY columns lands here
  const int xbytes = dl_xbuf_bytes(BN, S);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(xbuf + xbytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* xbar = tmem_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(xbar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = S > 1 ? int(ptx::cluster_ctarank()) : 0;
  const int tb = blockIdx.x / S;  // output tile (N direction)

  FIB_PROFILER_DECL
#ifdef FIB200_ENABLE_PROFILER
  {
    const int grp = warp == 0 ? 0 : (warp == 1 ? 1 : 2);
    const bool wr = p.prof != nullptr && lane == 0 && (warp == 0 || warp == 1 || warp == 4);
    FIB_PROFILER_INIT(p.prof ? p.prof : reinterpret_cast<uint64_t*>(smem_raw), grp, 3, wr, p.prof_events);
  }
#endif
  FIB_PROFILER_EVENT_START(kEvSetup);
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    ptx::mbar_init(tmem_full, 1);
    ptx::mbar_init(xbar, 1);
    ptx::fence_mbar_init();
    if (S > 1) ptx::mbar_arrive_expect_tx(xbar, uint32_t(xbytes));  // armed before the peer can possibly send
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < uint32_t(BN)) tmem_cols <<= 1;
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_ptr, tmem_cols);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (S > 1) ptx::cluster_sync();  // the peer's barrier / exchange buffer exist before anything is pushed
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  FIB_PROFILER_EVENT_END(kEvSetup);
  // PDL: only the activation loads and the epilogue inputs depend on the previous kernel
  if (warp != 0) {
    FIB_PROFILER_EVENT_START(kEvWaitPrevGrid);
    ptx::grid_dep_wait();
    FIB_PROFILER_EVENT_END(kEvWaitPrevGrid);
  }
  ptx::grid_dep_launch();

  const int kb0 = (crank * p.kblocks) / S, kb1 = ((crank + 1) * p.kblocks) / S;

  if (warp == 0) {
    if (ptx::elect_one()) {
      const int nkb = kb1 - kb0;
      const int npre = nkb < kStages ? nkb : kStages;
      for (int i = 0; i < npre; ++i) {  // weights first: independent of the previous kernel
        ptx::mbar_arrive_expect_tx(&full_bar[i], uint32_t(stage_bytes));
        if (p.w_blockk) ptx::tma_load_3d(smem + i * stage_bytes + kABytes, &tmB, &full_bar[i], 0, tb * BN, kb0 + i, ptx::kEvictFirst);
        else ptx::tma_load_2d(smem + i * stage_bytes + kABytes, &tmB, &full_bar[i], (kb0 + i) * BK, tb * BN, ptx::kEvictFirst);
      }
      FIB_PROFILER_EVENT_INSTANT(kEvWeightPrefetch);
      FIB_PROFILER_EVENT_START(kEvWaitPrevGrid);
      ptx::grid_dep_wait();
      FIB_PROFILER_EVENT_END(kEvWaitPrevGrid);
      FIB_PROFILER_EVENT_START(kEvMainLoop);
      for (int i = 0; i < npre; ++i)
        ptx::tma_load_2d(smem + i * stage_bytes, &tmA, &full_bar[i], (kb0 + i) * BK, 0, ptx::kEvictLast);
      int stage = npre == kStages ? 0 : npre;
      uint32_t phase = npre == kStages ? 1 : 0;
      for (int kb = kb0 + npre; kb < kb1; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * stage_bytes;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], uint32_t(stage_bytes));
        ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, 0, ptx::kEvictLast);
        if (p.w_blockk) ptx::tma_load_3d(sa + kABytes, &tmB, &full_bar[stage], 0, tb * BN, kb, ptx::kEvictFirst);
        else ptx::tma_load_2d(sa + kABytes, &tmB, &full_bar[stage], kb * BK, tb * BN, ptx::kEvictFirst);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      FIB_PROFILER_EVENT_END(kEvMainLoop);
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    FIB_PROFILER_EVENT_START(kEvMainLoop);
    for (int kb = kb0; kb < kb1; ++kb) {
      ptx::mbar_wait(&full_bar[stage], phase);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t sa = ptx::smem_u32(smem + stage * stage_bytes);
        const uint64_t da = ptx::make_smem_desc(sa, 16, 1024, ptx::kSwz128);
        const uint64_t db = ptx::make_smem_desc(sa + kABytes, 16, 1024, ptx::kSwz128);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          ptx::mma_f16_ss<1>(tmem_base, ptx::desc_advance(da, k * 32), ptx::desc_advance(db, k * 32), idesc,
                             (kb > kb0 || k > 0) ? 1u : 0u);
        ptx::mma_commit(&empty_bar[stage]);
        if (kb == kb1 - 1) ptx::mma_commit(tmem_full);
      }
      __syncwarp();
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
    FIB_PROFILER_EVENT_END(kEvMainLoop);
  } else if (warp >= 4) {
    // ===================== epilogue: thread = token row (UMMA M = 64: rows 16q..16q+15 in lanes 0-15 of quadrant q) ==========
    const int q = warp - 4, etid = threadIdx.x - 128;
    const bool row_ok = lane < 16;
    const int m = q * 16 + (lane & 15);
    const bool m_ok = row_ok && m < p.M;
    const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
    const int own_w = BN / S;
    const int own_lo = crank * own_w;
    const int xsw = (m >> 1) & 3;  // 16-byte slot swizzle of the exchange rows (bank spread)
    // all-reduce: the rotating-buffer epoch is read here, right after griddepcontrol.wait and long before any CTA of this grid can
    // finish its all-reduce and bump it (published to the other epilogue threads by the named barrier after the tile staging)
    if (p.epi == kResid && p.world > 1 && etid == 0) tmem_ptr[1] = *reinterpret_cast<volatile uint32_t*>(p.epoch);
    // epilogue inputs that do not depend on the accumulators are fetched while the main loop runs: the folded-RMSNorm scale of
    // the row, and (single-GPU residual epilogue, strips of <= 64 columns) the residual values this thread will update
    float rs_pre = 1.f;
    if (p.row_sumsq != nullptr && m_ok) rs_pre = rsqrtf(p.row_sumsq[m] * p.inv_dim + p.eps);
    constexpr int kResPre = 4;  // 16-column chunks of the residual held in registers
    Vec16<T> res_pre[2 * kResPre];
    const bool res_prefetched = p.epi == kResid && p.world <= 1 && own_w <= 16 * kResPre;
    if (res_prefetched && m_ok) {
      const T* rp0 = reinterpret_cast<const T*>(p.resid) + int64_t(m) * p.ldr + tb * BN + own_lo;
#pragma unroll
      for (int c = 0; c < kResPre; ++c)
        if (c * 16 < own_w && tb * BN + own_lo + c * 16 < p.N) {
          res_pre[2 * c] = ld16(rp0 + c * 16);
          res_pre[2 * c + 1] = ld16(rp0 + c * 16 + 8);
        }
    }
    FIB_PROFILER_EVENT_START(kEvMainLoop);
    ptx::mbar_wait(tmem_full, 0);
    ptx::tc_fence_after();
    FIB_PROFILER_EVENT_END(kEvMainLoop);
    FIB_PROFILER_EVENT_START(kEvSplitKExchange);
    if (S > 1) {
      // push my partial of every peer's columns into that peer's exchange slot (slot index = my rank, skipping the owner)
      for (int pi = 1; pi < S; ++pi) {
        const int peer = (crank + pi) % S;
        const int slot = crank < peer ? crank : crank - 1;
        const uint32_t rx = ptx::mapa(ptx::smem_u32(xbuf), uint32_t(peer)) + uint32_t(slot * own_w * BM * 4);
        const uint32_t rbar = ptx::mapa(ptx::smem_u32(xbar), uint32_t(peer));
        for (int c = 0; c < own_w; c += 16) {
          uint32_t r[16];
          ptx::tmem_ld_x16(taddr + peer * own_w + c, r);
          ptx::tmem_ld_wait();
          if (row_ok) {
            const uint32_t base = rx + uint32_t(((c >> 4) * BM + m) * 64);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ptx::st_async_v4(base + uint32_t((j ^ xsw) * 16),
                               make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                           __uint_as_float(r[4 * j + 3])),
                               rbar);
          }
        }
      }
      ptx::mbar_wait(xbar, 0);  // tx-count completion (like a TMA write): all peers' partials of my columns have landed
    }
    FIB_PROFILER_EVENT_END(kEvSplitKExchange);
    FIB_PROFILER_EVENT_START(kEvEpilogue);
    const float rs = rs_pre;

    // fp32 values of 16 owned columns starting at tile column own_lo + c (warp-collective TMEM load)
    auto load_chunk = [&](int c, float* v) {
      uint32_t r[16];
      ptx::tmem_ld_x16(taddr + own_lo + c, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      if (S > 1 && row_ok) {
        for (int sl = 0; sl < S - 1; ++sl) {
          const float4* src = reinterpret_cast<const float4*>(xbuf + sl * own_w * BM * 4 + ((c >> 4) * BM + m) * 64);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 x = src[j ^ xsw];
            v[4 * j] += x.x;
            v[4 * j + 1] += x.y;
            v[4 * j + 2] += x.z;
            v[4 * j + 3] += x.w;
          }
        }
      }
    };
    // 32 columns per TMEM round trip (tcgen05.wait::ld waits for every outstanding load of the thread, so the way to have fewer
    // serialized round trips is wider loads): the wide epilogues (gate/up: 208 columns per thread) are bound by that latency chain
    auto load_chunk32 = [&](int c, float* v) {
      uint32_t r[32];
      ptx::tmem_ld_x32(taddr + own_lo + c, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (S > 1 && row_ok) {
        for (int sl = 0; sl < S - 1; ++sl) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4* src = reinterpret_cast<const float4*>(xbuf + sl * own_w * BM * 4 + (((c >> 4) + h) * BM + m) * 64);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 x = src[j ^ xsw];
              v[16 * h + 4 * j] += x.x;
              v[16 * h + 4 * j + 1] += x.y;
              v[16 * h + 4 * j + 2] += x.z;
              v[16 * h + 4 * j + 3] += x.w;
            }
          }
        }
      }
    };
    const int n_base = tb * BN + own_lo;

    if (p.epi == kPlain) {
      const T* bias = reinterpret_cast<const T*>(p.bias);
      for (int c = 0; c < own_w; c += 16) {
        float v[16];
        load_chunk(c, v);
        const int n0 = n_base + c;
        if (m_ok && n0 < p.N) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= rs;
          T* dst = reinterpret_cast<T*>(p.out) + int64_t(m) * p.ldo + n0;
          if (n0 + 16 <= p.N && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            if (bias) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += to_f32(bias[n0 + j]);
            }
            store16<T>(dst, v);
          } else {
            for (int j = 0; j < 16 && n0 + j < p.N; ++j) dst[j] = from_f32<T>(v[j] + (bias ? to_f32(bias[n0 + j]) : 0.f));
          }
        }
      }
    } else if (p.epi == kGated) {
      int c = 0;
      for (; c + 32 <= own_w; c += 32) {
        float v[32];
        load_chunk32(c, v);
        const int n0 = n_base + c;
        if (m_ok && n0 < p.N) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (n0 + 16 * h < p.N) {
              Vec16<T> o;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float g = v[16 * h + 2 * j] * rs, u = v[16 * h + 2 * j + 1] * rs;
                o.v[j] = from_f32<T>(g / (1.f + __expf(-g)) * u);
              }
              st16(reinterpret_cast<T*>(p.out) + int64_t(m) * p.ldo + (n0 + 16 * h) / 2, o);
            }
          }
        }
      }
      for (; c < own_w; c += 16) {
        float v[16];
        load_chunk(c, v);
        const int n0 = n_base + c;
        if (m_ok && n0 < p.N) {
          Vec16<T> o;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float g = v[2 * j] * rs, u = v[2 * j + 1] * rs;
            o.v[j] = from_f32<T>(g / (1.f + __expf(-g)) * u);
          }
          st16(reinterpret_cast<T*>(p.out) + int64_t(m) * p.ldo + n0 / 2, o);
        }
      }
    } else if (p.epi == kResid) {
      float ss = 0.f;
      T* resid = reinterpret_cast<T*>(p.resid);
      if (p.world <= 1) {
        for (int c = 0; c < own_w; c += 16) {
          float v[16];
          load_chunk(c, v);
          const int n0 = n_base + c;
          if (m_ok && n0 < p.N) {
            T* rp = resid + int64_t(m) * p.ldr + n0;
            Vec16<T> a, b;
            if (res_prefetched) {
              // (compile-time indices: the chunk loop is at most kResPre long on this path)
              a = res_pre[0];
              b = res_pre[1];
#pragma unroll
              for (int q2 = 1; q2 < kResPre; ++q2)
                if (c == 16 * q2) {
                  a = res_pre[2 * q2];
                  b = res_pre[2 * q2 + 1];
                }
            } else {
              a = ld16(rp);
              b = ld16(rp + 8);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              a.v[e] = from_f32<T>(to_f32(a.v[e]) + v[e]);
              b.v[e] = from_f32<T>(to_f32(b.v[e]) + v[8 + e]);
              const float x = to_f32(a.v[e]), y = to_f32(b.v[e]);
              ss += x * x + y * y;
            }
            st16(rp, a);
            st16(rp + 8, b);
          }
        }
      } else {
        // ---- in-kernel all-reduce of this CTA's column strip over NVLink; the data is its own arrival signal (Lamport) ----
        // (0) fp32 accumulators -> 16-bit partial (never the sentinel) -> shared-memory tile [64][own_w] in the idle TMA ring, so
        //     that the exchange below runs with 16-byte pieces of one row in neighbouring lanes: every NVLink packet carries
        //     own_w * 2 contiguous bytes of a row instead of one 16-byte piece per token row.
        constexpr uint32_t kSent = 0x80008000u;  // two -0.0 halves
        uint32_t* s_epoch = tmem_ptr + 1;        // read right after griddepcontrol.wait (top of this warp role)
        uint8_t* tile = smem;
        const int P = own_w >> 3;                // 16-byte pieces per strip row
        const int pmask = (P & (P - 1)) == 0 ? P - 1 : 0;
        auto stage16 = [&](const float* v, int c) {  // 16 accumulator columns -> two 16-byte pieces of the staged strip
          Vec16<T> a, b;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            a.v[e] = from_f32<T>(v[e]);
            b.v[e] = from_f32<T>(v[8 + e]);
          }
          uint32_t* wa = reinterpret_cast<uint32_t*>(&a);
          uint32_t* wb = reinterpret_cast<uint32_t*>(&b);
#pragma unroll
          for (int e = 0; e < 4; ++e) {  // -0.0 is the sentinel: never send it
            if ((wa[e] & 0xffffu) == 0x8000u) wa[e] &= 0xffff0000u;
            if ((wa[e] >> 16) == 0x8000u) wa[e] &= 0x0000ffffu;
            if ((wb[e] & 0xffffu) == 0x8000u) wb[e] &= 0xffff0000u;
            if ((wb[e] >> 16) == 0x8000u) wb[e] &= 0x0000ffffu;
          }
          const int sw = (m >> 1) & pmask, pc = c >> 3;
          *reinterpret_cast<int4*>(tile + ((m * P + (pc ^ sw)) << 4)) = *reinterpret_cast<const int4*>(&a);
          *reinterpret_cast<int4*>(tile + ((m * P + ((pc + 1) ^ sw)) << 4)) = *reinterpret_cast<const int4*>(&b);
        };
        {
          int c = 0;
          for (; c + 32 <= own_w; c += 32) {
            float v[32];
            load_chunk32(c, v);
            if (row_ok) {
              stage16(v, c);
              stage16(v + 16, c + 16);
            }
          }
          for (; c < own_w; c += 16) {
            float v[16];
            load_chunk(c, v);
            if (row_ok) stage16(v, c);
          }
        }
        ptx::named_bar_sync(1, 128);
        const uint32_t ep = *s_epoch;
        T* recv = reinterpret_cast<T*>(p.recv);
        const int W = p.world;
        const int64_t cur = int64_t(ep % 3u) * p.buf_elems, nxt = int64_t((ep + 1u) % 3u) * p.buf_elems;
        const int nvec = BM * P;  // vectors of the strip; thread -> (row = vi / P, piece = vi % P)
        const int4 sent = make_int4(int(kSent), int(kSent), int(kSent), int(kSent));
        auto tile_vec = [&](int row, int pc) { return *reinterpret_cast<const int4*>(tile + ((row * P + (pc ^ ((row >> 1) & pmask))) << 4)); };
        // poll `n` 16-byte vectors (src + r * stride) until none holds a sentinel half; all loads of a round are in flight together
        auto poll = [&](const T* src, int64_t stride, int n, int4* x) {
          uint32_t polls = 0;
          uint64_t t0 = 0;
          bool done;
          do {
            done = true;
#pragma unroll
            for (int r = 0; r < 8; ++r)
              if (r < n) x[r] = ptx::ld_volatile_v4(src + int64_t(r) * stride);
#pragma unroll
            for (int r = 0; r < 8; ++r)
              if (r < n) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(&x[r]);
#pragma unroll
                for (int e = 0; e < 4; ++e) done = done && ((w[e] & 0xffffu) != 0x8000u) && ((w[e] >> 16) != 0x8000u);
              }
            if (!done && (++polls & 0x3ffu) == 0) {  // watchdog: a peer that never sends traps this kernel instead of hanging the GPU
              if (t0 == 0) t0 = ptx::globaltimer();
              else if (ptx::globaltimer() - t0 > ptx::kSpinTimeoutNs) {
                printf("fib200: decode_linear all-reduce watchdog: rank %d cta %d waited 20 s for peer data -> trap\n", p.rank, int(blockIdx.x));
                __trap();
              }
            }
          } while (!done);
        };
        // residual += sum (16-bit rounding of the new residual), returns the sum of squares of the 8 new values
        auto add_resid = [&](int row, int col, const float* sum, bool have, Vec16<T> a) {
          T* rp = resid + int64_t(row) * p.ldr + col;
          if (!have) a = ld16(rp);
          float s2 = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            a.v[e] = from_f32<T>(to_f32(a.v[e]) + sum[e]);
            const float x = to_f32(a.v[e]);
            s2 += x * x;
          }
          st16(rp, a);
          return s2;
        };
        // per-row sum of squares: the P lanes of a row are neighbours when P is a power of two
        auto row_ss = [&](float s2, int row, int pc, bool ok) {
          if (pmask) {
            for (int o = 1; o < P && o < 32; o <<= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            if (ok && pc == 0 && p.sumsq_out) atomicAdd(p.sumsq_out + row, s2);
          } else if (ok && p.sumsq_out) {
            atomicAdd(p.sumsq_out + row, s2);
          }
        };
        if (p.ar_algo != 2) {
          // ======== one-shot: multicast my strip into slot [rank] of every rank, gather all `world` slots, reduce in rank order ========
          for (int vi = etid; vi < nvec; vi += 128) {
            const int row = vi / P, pc = vi - row * P, col = n_base + pc * 8;
            if (row < p.M && col < p.N) {
              const int4 x = tile_vec(row, pc);
              const int64_t off = cur + int64_t(p.rank) * p.slot_elems + int64_t(row) * p.lds + col;
              if (p.mc_recv) {
                ptx::multimem_st_v4(reinterpret_cast<T*>(p.mc_recv) + off, x);
              } else {
                for (int r = 0; r < W; ++r) ptx::st_na_v4(reinterpret_cast<T*>(p.peer_recv[(p.rank + r) % W]) + off, x);
              }
            }
          }
          // the residual values this thread will update do not depend on the peers: fetch them now, under the NVLink latency
          constexpr int kPre = 4;
          Vec16<T> rpre[kPre];
          const bool pre_ok = nvec <= kPre * 128;
          if (pre_ok) {
#pragma unroll
            for (int q2 = 0; q2 < kPre; ++q2) {
              const int vi = etid + q2 * 128;
              const int row = vi / P, pc = vi - row * P, col = n_base + pc * 8;
              if (vi < nvec && row < p.M && col < p.N) rpre[q2] = ld16(resid + int64_t(row) * p.ldr + col);
            }
          }
          // while the pushes fly: reset my strip (all 64 rows: the next call may carry more tokens) of the NEXT call's buffer, last read two calls ago
          for (int vi = etid; vi < nvec; vi += 128) {
            const int row = vi / P, pc = vi - row * P, col = n_base + pc * 8;
            if (col < p.N)
              for (int r = 0; r < W; ++r) *reinterpret_cast<int4*>(recv + nxt + int64_t(r) * p.slot_elems + int64_t(row) * p.lds + col) = sent;
          }
          for (int vi = etid; vi < nvec; vi += 128) {
            const int row = vi / P, pc = vi - row * P, col = n_base + pc * 8;
            const bool ok = row < p.M && col < p.N;
            float s2 = 0.f;
            if (ok) {
              int4 x[8];
              poll(recv + cur + int64_t(row) * p.lds + col, p.slot_elems, W, x);
              float sum[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) sum[e] = 0.f;
#pragma unroll
              for (int r = 0; r < 8; ++r)
                if (r < W) {
                  const T* h = reinterpret_cast<const T*>(&x[r]);
#pragma unroll
                  for (int e = 0; e < 8; ++e) sum[e] += to_f32(h[e]);
                }
              Vec16<T> rv = rpre[0];
#pragma unroll
              for (int q2 = 1; q2 < kPre; ++q2)
                if (vi == etid + q2 * 128) rv = rpre[q2];
              s2 = add_resid(row, col, sum, pre_ok, rv);
            }
            row_ss(s2, row, pc, ok);
          }
        } else {
          // ======== two-shot: reduce-scatter by push (row m is owned by rank m % world), the owner adds the residual and multicasts
          //          the NEW residual rows to every rank.  Two one-way NVLink hops, (1 + 1/world) x strip bytes of ingress per rank
          //          instead of world x: the choice for world >= 8 (measured equal at 4).  Slot 0 of the rotating buffer is the reduce-scatter inbox
          //          [src rank][owned row][hidden], slot 1 the all-gather inbox [row][hidden]. ========
          const int R = BM / W;  // rows per owner (world is 2, 4 or 8)
          for (int vi = etid; vi < nvec; vi += 128) {
            const int row = vi / P, pc = vi - row * P, col = n_base + pc * 8;
            if (row < p.M && col < p.N) {
              const int dst = row % W, lr = row / W;
              ptx::st_na_v4(reinterpret_cast<T*>(p.peer_recv[dst]) + cur + int64_t(p.rank * R + lr) * p.lds + col, tile_vec(row, pc));
            }
          }
          for (int vi = etid; vi < nvec; vi += 128) {  // reset both inboxes of the NEXT call's buffer (my strip, all 64 rows)
            const int row = vi / P, pc = vi - row * P, col = n_base + pc * 8;
            if (col < p.N) {
              *reinterpret_cast<int4*>(recv + nxt + int64_t(row) * p.lds + col) = sent;
              *reinterpret_cast<int4*>(recv + nxt + p.slot_elems + int64_t(row) * p.lds + col) = sent;
            }
          }
          for (int it = etid; it < R * P; it += 128) {  // owner: gather the `world` partials of my rows, reduce, + residual, broadcast
            const int lr = it / P, pc = it - lr * P, col = n_base + pc * 8, row = lr * W + p.rank;
            if (row < p.M && col < p.N) {
              int4 x[8];
              const Vec16<T> old = ld16(resid + int64_t(row) * p.ldr + col);  // issued before the poll: hidden under the NVLink hop
              poll(recv + cur + int64_t(lr) * p.lds + col, int64_t(R) * p.lds, W, x);
              float sum[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) sum[e] = 0.f;
#pragma unroll
              for (int r = 0; r < 8; ++r)
                if (r < W) {
                  const T* h = reinterpret_cast<const T*>(&x[r]);
#pragma unroll
                  for (int e = 0; e < 8; ++e) sum[e] += to_f32(h[e]);
                }
              Vec16<T> nw;
#pragma unroll
              for (int e = 0; e < 8; ++e) nw.v[e] = from_f32<T>(to_f32(old.v[e]) + sum[e]);
              uint32_t* w = reinterpret_cast<uint32_t*>(&nw);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if ((w[e] & 0xffffu) == 0x8000u) w[e] &= 0xffff0000u;
                if ((w[e] >> 16) == 0x8000u) w[e] &= 0x0000ffffu;
              }
              const int64_t off = cur + p.slot_elems + int64_t(row) * p.lds + col;
              if (p.mc_recv) {
                ptx::multimem_st_v4(reinterpret_cast<T*>(p.mc_recv) + off, *reinterpret_cast<const int4*>(&nw));
              } else {
                for (int r = 0; r < W; ++r)
                  ptx::st_na_v4(reinterpret_cast<T*>(p.peer_recv[(p.rank + r) % W]) + off, *reinterpret_cast<const int4*>(&nw));
              }
            }
          }
          for (int vi = etid; vi < nvec; vi += 128) {  // everyone: the new residual rows of my strip arrive from their owners
            const int row = vi / P, pc = vi - row * P, col = n_base + pc * 8;
            const bool ok = row < p.M && col < p.N;
            float s2 = 0.f;
            if (ok) {
              int4 x[8];
              poll(recv + cur + p.slot_elems + int64_t(row) * p.lds + col, 0, 1, x);
              const T* h = reinterpret_cast<const T*>(&x[0]);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float f = to_f32(h[e]);
                s2 += f * f;
              }
              *reinterpret_cast<int4*>(resid + int64_t(row) * p.ldr + col) = x[0];
            }
            row_ss(s2, row, pc, ok);
          }
        }
        // every CTA read the epoch right after griddepcontrol.wait, long before any CTA can get here: bump it for the next call
        if (blockIdx.x == 0 && etid == 0) *reinterpret_cast<volatile uint32_t*>(p.epoch) = ep + 1u;
      }
      if (p.world <= 1) {
        if (m_ok && p.sumsq_out) atomicAdd(p.sumsq_out + m, ss);
      }
    } else {  // kRope
      const int hd = p.head_dim, half = hd / 2;
      T* k_cache = reinterpret_cast<T*>(p.k_cache);
      T* v_cache = reinterpret_cast<T*>(p.v_cache);
      const int64_t crow = m_ok ? p.cache_row[m] : 0;
      const float* cs = p.cos_sin + int64_t(m_ok ? m : 0) * hd;
      for (int c = 0; c < own_w; c += 16) {
        float v[16];
        load_chunk(c, v);
        const int n0 = n_base + c;
        const int head = n0 / hd, within = n0 - head * hd;
        if (!(m_ok && n0 < p.N)) {
          // nothing to store (padding row / column); no `continue`: the warp stays converged for the next TMEM load
        } else if (head >= p.hq + p.hkv) {  // V: straight into the cache slot
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= rs;
          store16<T>(v_cache + crow + int64_t(head - p.hq - p.hkv) * p.c_sh + within, v);
        } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= rs;
        T* dst = head < p.hq ? reinterpret_cast<T*>(p.out) + int64_t(m) * p.ldo + int64_t(head) * hd
                             : k_cache + crow + int64_t(head - p.hq) * p.c_sh;
        const int j0 = within >> 1;  // first rotation pair of this chunk
        float o1[8], o2[8];
        const float4 c0 = *reinterpret_cast<const float4*>(cs + j0), c1 = *reinterpret_cast<const float4*>(cs + j0 + 4);
        const float4 s0 = *reinterpret_cast<const float4*>(cs + half + j0), s1 = *reinterpret_cast<const float4*>(cs + half + j0 + 4);
        const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x1 = v[2 * e], x2 = v[2 * e + 1];
          o1[e] = x1 * cc[e] - x2 * sn[e];
          o2[e] = x2 * cc[e] + x1 * sn[e];
        }
        if (p.interleave) {  // pairs stay where they are: (2j, 2j + 1)
          float w[16];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            w[2 * e] = o1[e];
            w[2 * e + 1] = o2[e];
          }
          store16<T>(dst + within, w);
        } else {  // NeoX halves: weight rows were permuted (j, j + hd/2) -> adjacent columns; un-permute on the way out
          Vec16<T> a, b;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            a.v[e] = from_f32<T>(o1[e]);
            b.v[e] = from_f32<T>(o2[e]);
          }
          st16(dst + j0, a);
          st16(dst + half + j0, b);
        }
        }
      }
    }
  }

  if (warp >= 4) { FIB_PROFILER_EVENT_END(kEvEpilogue); }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Per-step preparation (one launch): embedding gather -> residual stream + its sum of squares, RoPE cos / sin of every token's
// position, the element offset of every token's KV-cache slot (same page table for all layers), and the reset of the
// per-layer sum-of-squares accumulators.
// ---------------------------------------------------------------------------------------------------------------------
struct PrepParams {
  const int64_t* tokens;
  const void* embed;
  void* resid;
  int64_t lde, ldr;
  int hidden;
  float* sumsq;       // [n_sumsq][64]; slice 0 <- embedding rows, the rest <- 0
  int n_sumsq;
  const int32_t* positions;
  const int32_t* batch_indices;
  const int32_t* kv_indptr;
  const int32_t* kv_indices;
  int page_size;
  int64_t c_sp, c_sn;
  float* cos_sin;     // [M, head_dim]
  int64_t* cache_row; // [M]
  int head_dim, interleave;
  float rope_rcp_scale, rope_theta_log2, smooth_a, smooth_b;
  int llama31;
  int M;
};

template <typename T>
__global__ void __launch_bounds__(256) decode_prep_kernel(const PrepParams p) {
  constexpr int VN = 16 / sizeof(T);
  __shared__ float red[8];
  const int m = blockIdx.x;
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();
  // reset the accumulators of the other layers (slice 0 is written below)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (p.n_sumsq - 1) * 64; i += gridDim.x * blockDim.x) p.sumsq[64 + i] = 0.f;
  if (m >= p.M) return;
  const int64_t tok = p.tokens[m];
  const T* src = reinterpret_cast<const T*>(p.embed) + tok * p.lde;
  T* dst = reinterpret_cast<T*>(p.resid) + int64_t(m) * p.ldr;
  float ss = 0.f;
  for (int c = threadIdx.x * VN; c < p.hidden; c += blockDim.x * VN) {
    const Vec16<T> v = ldg16(src + c);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      const float x = to_f32(v.v[e]);
      ss += x * x;
    }
    st16(dst + c, v);
  }
  ss = warp_reduce_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    p.sumsq[m] = t;
    const int pos = p.positions[m];
    const int b = p.batch_indices ? p.batch_indices[m] : m;
    const int page = p.kv_indices[p.kv_indptr[b] + pos / p.page_size];
    p.cache_row[m] = int64_t(page) * p.c_sp + int64_t(pos % p.page_size) * p.c_sn;
  }
  const int half = p.head_dim / 2;
  if (threadIdx.x < half) {
    const int i = threadIdx.x;
    float inv = exp2f(-p.rope_theta_log2 * float(2 * i) / float(p.head_dim));
    if (p.llama31) {
      float smooth = inv * p.smooth_a + p.smooth_b;
      smooth = fminf(fmaxf(smooth, 0.f), 1.f);
      inv = (1.f - smooth) * (inv * p.rope_rcp_scale) + smooth * inv;
    } else {
      inv *= p.rope_rcp_scale;
    }
    float sn, cs;
    sincosf(float(p.positions[m]) * inv, &sn, &cs);
    p.cos_sin[int64_t(m) * p.head_dim + i] = cs;
    p.cos_sin[int64_t(m) * p.head_dim + half + i] = sn;
  }
}

struct ProfArm {
  uint64_t* buf = nullptr;
  int max_events = 0;
};
thread_local ProfArm g_prof;  // consumed (and cleared) by the next dlinear_run of this thread

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace

// A [M, K] (lda), W [N, K] (ldw), both K-major 16-bit.  `epi`: 0 plain, 1 gated SiLU (out [M, N/2]), 2 residual (+ all-reduce),
// 3 RoPE + paged-KV append.  peer_recv_host: host int64 array of `world` device pointers (or null when multicast is used).
extern "C" int dlinear_run(void* A, void* W, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldw, int64_t epi, void* out,
                           int64_t ldo, void* bias, void* row_sumsq, double inv_dim, double eps, void* resid, int64_t ldr,
                           void* sumsq_out, int64_t world, int64_t rank, void* recv, int64_t lds, int64_t slot_elems,
                           void* mc_recv, void* epoch, void* peer_recv_host, int64_t ar_algo, void* cos_sin,
                           void* cache_row, void* k_cache, void* v_cache, int64_t c_sh, int64_t hq, int64_t hkv, int64_t head_dim,
                           int64_t interleave, int64_t force_bn, int64_t force_s, int64_t smem_kb, int64_t w_blockk, int64_t dtype,
                           int64_t pdl, int64_t stream_) {
  FIB_CHECK(M >= 1 && M <= BM, "dlinear: 1 <= M <= 64 (decode batch); larger batches use gemm_nt");
  FIB_CHECK(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "dlinear: K / lda / ldw must be multiples of 8 (16 B TMA alignment)");
  FIB_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "dlinear: A / W must be 16 B aligned");
  FIB_CHECK(epi >= 0 && epi <= 3, "dlinear: bad epilogue id");
  if (epi != kPlain) FIB_CHECK(N % 16 == 0, "dlinear: fused epilogues need N % 16 == 0");
  if (epi == kGated) FIB_CHECK(ldo % 8 == 0, "dlinear (gated): ldo % 8");
  if (epi == kResid) FIB_CHECK(resid != nullptr && ldr % 8 == 0, "dlinear (resid): residual pointer / ldr % 8");
  if (epi == kRope)
    FIB_CHECK(cos_sin && cache_row && k_cache && v_cache && head_dim % 32 == 0 && N == (hq + 2 * hkv) * head_dim && ldo % 8 == 0 &&
                  c_sh % 8 == 0,
              "dlinear (rope): cos_sin / cache_row / caches required, N = (hq + 2 hkv) * head_dim, head_dim % 32 == 0");
  if (epi == kResid && world > 1) {
    FIB_CHECK(world <= 8 && recv && epoch && lds % 8 == 0 && slot_elems >= BM * lds && (mc_recv || peer_recv_host),
              "dlinear (all-reduce): world <= 8, symmetric receive buffers (multicast alias or peer table) and the epoch word required");
    if (ar_algo == 0) ar_algo = (world >= 8 && BM % world == 0 && peer_recv_host) ? 2 : 1;
    FIB_CHECK(ar_algo == 1 || ar_algo == 2, "dlinear (all-reduce): ar_algo must be 0 (auto), 1 (one-shot) or 2 (two-shot)");
    if (ar_algo == 2) FIB_CHECK(BM % world == 0 && peer_recv_host, "dlinear (two-shot all-reduce): world must divide 64 and the peer table is required");
  }
  const int sms = num_sms();
  const int kblocks = int((K + BK - 1) / BK);
  // ---- tile plan: one wave, one tile per cluster.  BN = narrowest multiple of 16 (32 with the 2-CTA K split) that covers N with
  //      <= #SM CTAs; the K split halves the activation bytes each SM has to ingest per weight byte ----
  int S = 1;
  int BN = int(((N + sms - 1) / sms + 15) / 16 * 16);
  if (BN < 32) BN = 32;
  if (BN > 256) {
    // several waves of one-tile CTAs (e.g. the LM head): time ~ waves x (weight rows + the 64 activation rows) per k-block
    int best = 256;
    long best_cost = 1L << 60;
    for (int bn = 256; bn >= 192; bn -= 16) {
      const long t = (N + bn - 1) / bn;
      const long cost = ((t + sms - 1) / sms) * long(bn + BM);
      if (cost < best_cost) {
        best_cost = cost;
        best = bn;
      }
    }
    BN = best;
  }
  if (BN <= 64 && kblocks >= 16) {
    // split-K over a cluster: per-SM ingest (activation tile + weight tile per k-block) is the limiter of these single-wave
    // GEMMs, and the 64 activation rows weigh as much as 64 weight rows.  Plans below are the winners of tools/dl_sweep.py on
    // B200 for the Llama-3-8B decode shapes at TP 1 / 2 / 4 / 8 (gpurun_out/r14_sweep.log, profiles/decode_linear_plans.md).
    S = 2;
    BN = int(((N * 2 + sms - 1) / sms + 31) / 32 * 32);
    if (BN < 64) BN = 64;
    const int t128 = int((N + 127) / 128);
    const int cl_sms = (sms * 132) / 148;  // clusters of 4 / 8 CTAs can be placed on 132 of the 148 SMs
    if (kblocks >= 32 && t128 * 4 <= cl_sms && t128 * 4 >= 96) {
      // 128-wide tiles over a 4-CTA cluster when that still fills the SMs: the activation tile is re-read from L2 once per 128
      // weight rows instead of once per 64 (M = 64: o_proj 12.5 -> 11.4 us, down_proj 30.0 -> 24.2 us)
      S = 4;
      BN = 128;
    } else if (kblocks >= 32 && t128 * 4 < 96 && t128 * 8 <= cl_sms) {
      // narrow N (the QKV projection of a TP 4 / 8 shard): only an 8-way K split reaches enough SMs (TP8 qkv 10.6 -> 6.9 us)
      S = 8;
      BN = 128;
    }
  }
  static const int env_bn = env_int("FIB200_DL_BN", 0), env_s = env_int("FIB200_DL_S", 0), env_kb = env_int("FIB200_DL_SMEM_KB", 0);
  if (force_s > 0 || env_s > 0) S = force_s > 0 ? int(force_s) : env_s;
  if (force_bn > 0 || env_bn > 0) BN = force_bn > 0 ? int(force_bn) : env_bn;
  FIB_CHECK(S == 1 || S == 2 || S == 4 || S == 8, "dlinear: cluster split-K factor must be 1, 2, 4 or 8");
  if (w_blockk) FIB_CHECK(K % BK == 0, "dlinear: BlockMajorK weights need K % 64 == 0");
  FIB_CHECK(BN % (16 * S) == 0 && BN >= 16 * S && BN <= 256, "dlinear: BN must be a multiple of 16 * S in [16 S, 256]");
  if (S > 1) FIB_CHECK(kblocks >= S, "dlinear: split-K needs >= S k-blocks");
  if (S >= 4) FIB_CHECK(int((N + BN - 1) / BN) * S <= (sms * 132) / 148, "dlinear: clusters of 4 / 8 fit only 132 of the 148 SMs");
  const int tiles = int((N + BN - 1) / BN);
  FIB_CHECK(int64_t(tiles) * S <= 65535, "dlinear: too many tiles");
  if (epi == kResid && world > 1) FIB_CHECK(tiles * S <= sms, "dlinear (all-reduce): the grid must be a single wave (<= #SM CTAs)");
  int budget_kb = smem_kb > 0 ? int(smem_kb) : (env_kb > 0 ? env_kb : 216);
  const int stage_bytes = dl_stage_bytes(BN);
  const int xbytes = dl_xbuf_bytes(BN, S);
  int stages = (budget_kb * 1024 - xbytes - 1024) / stage_bytes;
  if (stages > 24) stages = 24;
  const int kb_per = (kblocks + S - 1) / S;
  if (stages > kb_per) stages = kb_per;
  FIB_CHECK(stages >= 2 || kb_per < 2, "dlinear: shared-memory budget too small for this tile");
  if (stages < 1) stages = 1;
  const int smem_total = stages * stage_bytes + xbytes + (2 * stages + 2) * 8 + 16 + 1024;
  FIB_CHECK(smem_total <= 227 * 1024, "dlinear: tile does not fit shared memory");

  DLP p;
  memset(&p, 0, sizeof(p));
  p.M = int(M); p.N = int(N); p.K = int(K); p.BN = BN; p.S = S; p.kblocks = kblocks; p.stages = stages; p.epi = int(epi); p.w_blockk = int(w_blockk);
  p.prof = g_prof.buf; p.prof_events = g_prof.max_events;
  g_prof = ProfArm();
  p.out = out; p.ldo = ldo; p.bias = bias;
  p.row_sumsq = reinterpret_cast<const float*>(row_sumsq); p.inv_dim = float(inv_dim); p.eps = float(eps);
  p.resid = resid; p.ldr = ldr; p.sumsq_out = reinterpret_cast<float*>(sumsq_out);
  p.world = int(world); p.rank = int(rank); p.ar_algo = int(ar_algo);
  p.recv = recv; p.lds = lds; p.slot_elems = slot_elems; p.buf_elems = slot_elems * world; p.mc_recv = mc_recv;
  p.epoch = reinterpret_cast<uint32_t*>(epoch);
  if (peer_recv_host) {
    const int64_t* ps = reinterpret_cast<const int64_t*>(peer_recv_host);
    for (int i = 0; i < world && i < kMaxRanks; ++i) p.peer_recv[i] = reinterpret_cast<void*>(ps[i]);
  }
  p.cos_sin = reinterpret_cast<const float*>(cos_sin); p.cache_row = reinterpret_cast<const int64_t*>(cache_row);
  p.k_cache = k_cache; p.v_cache = v_cache; p.c_sh = c_sh;
  p.hq = int(hq); p.hkv = int(hkv); p.head_dim = int(head_dim); p.interleave = int(interleave);

  const bool f16 = dtype == kF16;
  FIB_CHECK(dtype == kF16 || dtype == kBF16, "dlinear: f16 / bf16 only");
  const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {BK, BM};
    if (make_tmap(&tmA, dt, 2, A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  if (w_blockk) {  // [K / 64, N, 64]: ldw = elements between consecutive N rows (64 when dense)
    uint64_t dims[3] = {(uint64_t)BK, (uint64_t)N, (uint64_t)(K / BK)};
    uint64_t str[2] = {(uint64_t)ldw * 2, (uint64_t)N * (uint64_t)ldw * 2};
    uint32_t box[3] = {BK, (uint32_t)BN, 1};
    if (make_tmap(&tmB, dt, 3, W, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  } else {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)BN};
    if (make_tmap(&tmB, dt, 2, W, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  const uint32_t idesc = ptx::make_idesc_f16(f16 ? ptx::kFmtF16 : ptx::kFmtBF16, BM, BN, 0, 0);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LaunchCfg lc(dim3(tiles * S), dim3(256), smem_total, stream, pdl != 0, S);
  if (f16) {
    static bool attr = false;
    if (!attr) {
      FIB_CUDA_CHECK(cudaFuncSetAttribute(dlinear_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr = true;
    }
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, dlinear_kernel<__half>, tmA, tmB, p, idesc));
  } else {
    static bool attr = false;
    if (!attr) {
      FIB_CUDA_CHECK(cudaFuncSetAttribute(dlinear_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr = true;
    }
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, dlinear_kernel<__nv_bfloat16>, tmA, tmB, p, idesc));
  }
  return 0;
}

extern "C" int decode_prep_run(void* tokens, void* embed, void* resid, int64_t lde, int64_t ldr, int64_t hidden, void* sumsq,
                               int64_t n_sumsq, void* positions, void* batch_indices, void* kv_indptr, void* kv_indices,
                               int64_t page_size, int64_t c_sp, int64_t c_sn, void* cos_sin, void* cache_row, int64_t head_dim,
                               int64_t interleave, double rope_scale, double rope_theta, int64_t llama31, double low_freq_factor,
                               double high_freq_factor, double old_context_len, int64_t M, int64_t dtype, int64_t pdl,
                               int64_t stream_) {
  FIB_CHECK(M >= 1 && M <= BM && head_dim % 2 == 0 && head_dim <= 512 && n_sumsq >= 1, "decode_prep: bad shape");
  FIB_CHECK(hidden % 8 == 0 && lde % 8 == 0 && ldr % 8 == 0, "decode_prep: hidden / strides must be multiples of 8");
  PrepParams p;
  memset(&p, 0, sizeof(p));
  p.tokens = reinterpret_cast<const int64_t*>(tokens);
  p.embed = embed; p.resid = resid; p.lde = lde; p.ldr = ldr; p.hidden = int(hidden);
  p.sumsq = reinterpret_cast<float*>(sumsq); p.n_sumsq = int(n_sumsq);
  p.positions = reinterpret_cast<const int32_t*>(positions);
  p.batch_indices = reinterpret_cast<const int32_t*>(batch_indices);
  p.kv_indptr = reinterpret_cast<const int32_t*>(kv_indptr);
  p.kv_indices = reinterpret_cast<const int32_t*>(kv_indices);
  p.page_size = int(page_size); p.c_sp = c_sp; p.c_sn = c_sn;
  p.cos_sin = reinterpret_cast<float*>(cos_sin); p.cache_row = reinterpret_cast<int64_t*>(cache_row);
  p.head_dim = int(head_dim); p.interleave = int(interleave);
  p.rope_rcp_scale = float(1.0 / rope_scale);
  p.rope_theta_log2 = float(log2(rope_theta));
  p.llama31 = int(llama31);
  if (llama31) {
    p.smooth_a = float(old_context_len / (2.0 * M_PI * (high_freq_factor - low_freq_factor)));
    p.smooth_b = float(-1.0 / (high_freq_factor / low_freq_factor - 1.0));
  }
  p.M = int(M);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LaunchCfg lc(dim3(BM), dim3(256), 0, stream, pdl != 0);
  if (dtype == kF16) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, decode_prep_kernel<__half>, p));
  } else if (dtype == kBF16) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, decode_prep_kernel<__nv_bfloat16>, p));
  } else {
    return set_error("decode_prep: f16 / bf16 only");
  }
  return 0;
}

// Arms the intra-kernel profiler for the NEXT dlinear_run of this thread: `buf` from profiler.alloc_profiler_buffer(grid, 3, max_events).
// Only the FIB200_ENABLE_PROFILER build (module decode_linear_sm100_prof) records anything.
extern "C" int dlinear_set_profiler(void* buf, int64_t max_events) {
#ifndef FIB200_ENABLE_PROFILER
  if (buf) return set_error("dlinear_set_profiler: this library was built without FIB200_ENABLE_PROFILER (load decode_linear_sm100_prof)");
#endif
  g_prof.buf = reinterpret_cast<uint64_t*>(buf);
  g_prof.max_events = int(max_events);
  return 0;
}
