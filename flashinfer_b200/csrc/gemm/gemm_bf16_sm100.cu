// Dense bf16/fp16 GEMM for sm_100a:  C[M,N] = A[M,K] * B[N,K]^T (+bias)   ("NT", both K-major)
//
// B200-first design (parity target: reference mm_bf16 / tgv_gemm_sm100,
// flashinfer/gemm/gemm_base.py:485,1446 and include/flashinfer/gemm/tgv_gemm.cuh):
//   * persistent grid (<= #SM CTAs), static tile scheduler
//   * warp-specialised: warp0 = TMA producer, warp1 = single-thread tcgen05.mma issuer,
//     warp2 = TMEM allocator, warps 4-7 = epilogue (TMEM -> registers -> global)
//   * SW128 K-major operand tiles in a multi-stage smem ring fed by TMA, full/empty mbarriers
//   * fp32 accumulators double-buffered in TMEM so the epilogue of tile i overlaps the
//     mainloop of tile i+1
//   * "swap-AB" mode for small M (decode): the weight matrix takes the 128-wide MMA-M side and
//     the token dimension becomes MMA-N (16..128), so no tensor-core work is wasted on padding
//     and the kernel streams weights at HBM speed; split-K spreads long-K/short-N problems over
//     all SMs (fp32 partials + tiny reduce kernel).
//   * PDL: griddepcontrol.wait before the first global read, launch_dependents after the mainloop.
#include <stdlib.h>
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {
thread_local int g_sm_budget = 0;  // gemm_set_sm_budget: persistent-grid size of the next launches of this thread (0 = all SMs)


constexpr int BK = 64;   // 64 x 2B = one 128B swizzle span

// Shared-memory plan for a runtime tile width BN (multiple of 16, <= 256).
struct GemmSmem {
  int stages, stage_bytes, a_bytes, bar_offset, total;
  // budget_kb: ring budget.  220 = one CTA per SM; ~100 = two co-resident CTAs (decode GEMMs: the next kernel's CTA - launched
  // early through PDL - sets up and prefetches its weights while the previous kernel is still in its epilogue on the same SM)
  __host__ __device__ static GemmSmem make(int BMv, int BN, int budget_kb = 220) {
    GemmSmem g;
    g.a_bytes = BMv * BK * 2;
    g.stage_bytes = g.a_bytes + BN * BK * 2;
    int st = (budget_kb * 1024) / g.stage_bytes;
    g.stages = st > 16 ? 16 : st;
    g.bar_offset = g.stages * g.stage_bytes;
    g.total = g.bar_offset + 512 + 1024;
    return g;
  }
};

// ---------------------------------------------------------------------------------------------------
// Scheduler: W = tiles / grid full waves are data-parallel (tile = wave * grid + cta, grouped raster so
// that a wave covers a compact block of the output and its operands stay L2 resident); the R = tiles %
// grid remainder tiles are cut "stream-K" style into equal (tile, k-block) ranges over g_sk CTAs and
// finished by an in-kernel reduce-scatter fix-up: every part publishes its fp32 partial in an L2-resident
// slot, parts rendezvous on a per-tile counter (all CTAs are co-resident: persistent grid <= #SM), then
// each part sums and writes 1/parts of the tile (slot order => bitwise reproducible).
// ---------------------------------------------------------------------------------------------------
struct Sched {
  int tiles_a, tiles_b, kblocks, grid, BN;
  int S;  // cluster split-K factor (1 = off); when > 1 the grid is exactly tiles * S (one tile per cluster)
  int gated;    // epilogue computes out[:, j] = silu(acc[:, 2j]) * acc[:, 2j + 1] (SwiGLU with row-interleaved gate / up weights)
  int smem_kb;  // ring budget handed to GemmSmem::make (host and device must agree)
  int trace;  // debug (FIB200_GEMM_TRACE=1): the cluster split-K path writes clock64 stamps into the partial workspace
  int G;  // with S > 1: N tiles per cluster that share the A operand through TMA multicast (1 or 2); cluster = S * G CTAs
  int W, R, g_sk, max_parts, group_a;
  int64_t u_r;  // R * kblocks
  __device__ __forceinline__ int64_t sk_begin(int c) const { return (int64_t(c) * u_r) / g_sk; }
  __device__ __forceinline__ int sk_cta_of(int64_t u) const { return int(((u + 1) * g_sk + u_r - 1) / u_r) - 1; }
  __device__ __forceinline__ void coords(int t, int& ta, int& tb) const {
    // grouped rasterisation: walk `group_a` A-tiles for every B-tile before moving on
    const int per_group = group_a * tiles_b;
    const int g = t / per_group;
    const int first = g * group_a;
    const int rows = min(group_a, tiles_a - first);
    const int in = t - g * per_group;
    ta = first + in % rows;
    tb = in / rows;
  }
};

// Iterates the segments (tile, kb0, kb1) of one CTA: stream-K remainder first, then its DP tiles.
struct SegIter {
  const Sched& s;
  int cta;
  int64_t u, u_end;  // stream-K cursor
  int wave;
  __device__ SegIter(const Sched& s_, int cta_) : s(s_), cta(cta_), wave(0) {
    if (s.g_sk > 0 && cta < s.g_sk) {
      u = s.sk_begin(cta);
      u_end = s.sk_begin(cta + 1);
    } else {
      u = u_end = 0;
    }
  }
  // returns false when done. partial=true => needs the fix-up (r_idx = remainder tile index)
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1, bool& partial, int& r_idx) {
    if (u < u_end) {
      r_idx = int(u / s.kblocks);
      kb0 = int(u % s.kblocks);
      const int64_t lim = kb0 + (u_end - u);
      kb1 = lim < int64_t(s.kblocks) ? int(lim) : s.kblocks;
      u += kb1 - kb0;
      tile = s.W * s.grid + r_idx;
      partial = !(kb0 == 0 && kb1 == s.kblocks);
      return true;
    }
    partial = false;
    r_idx = 0;
    kb0 = 0;
    kb1 = s.kblocks;
    if (wave < s.W) {
      tile = wave * s.grid + cta;
      ++wave;
      return true;
    }
    if (wave == s.W && s.g_sk == 0 && cta < s.R) {  // remainder handled data-parallel
      tile = s.W * s.grid + cta;
      ++wave;
      return true;
    }
    return false;
  }
};

template <int BM, bool kSwap, typename OutT>
__global__ void __launch_bounds__(256, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, OutT* __restrict__ C,
               float* __restrict__ partial, int* __restrict__ counters, const OutT* __restrict__ bias, int rowsA,
               int rowsB, int K, int64_t ldc, const Sched sk, uint32_t idesc) {
  const long long t_entry = clock64();
  const int BN = sk.BN;
  const GemmSmem S = GemmSmem::make(BM, BN, sk.smem_kb);
  const int kStages = S.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S.bar_offset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* go_bar = tmem_empty + 2;        // cluster split-K: "leader smem is free, send your partial"
  uint64_t* partials_bar = go_bar + 1;      // leader only: all peers delivered their partials
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(partials_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S_split = sk.S;
  const bool clustered = S_split > 1 || sk.G > 1;  // one output tile per CTA, cluster = S (K split) x G (A multicast)
  const int crank = clustered ? int(ptx::cluster_ctarank()) : 0;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], clustered ? sk.G : 1);  // multicast: every CTA sharing A must have freed the slot
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 4);
    }
    ptx::mbar_init(go_bar, 1);
    ptx::mbar_init(partials_bar, S_split > 1 ? S_split - 1 : 1);
    ptx::fence_mbar_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < uint32_t(2 * BN)) tmem_cols <<= 1;
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_ptr, tmem_cols);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (clustered) ptx::cluster_sync();  // peers' mbarriers must exist before any remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // PDL: the weight operand does not depend on the previous kernel, so the TMA producer prefetches the first
  // pipeline stages of W before griddepcontrol.wait and only the activation loads wait (hides the pipeline ramp).
  if (warp != 0) ptx::grid_dep_wait();
  ptx::grid_dep_launch();

  long long* trace = (sk.trace && partial) ? reinterpret_cast<long long*>(partial) + int64_t(blockIdx.x) * 8 : nullptr;
  if (trace && threadIdx.x == 0) {
    trace[7] = t_entry;
    trace[0] = clock64();  // after setup + cluster_sync + griddepcontrol
  }
  if (clustered) {
    // =====================================================================================
    // Cluster split-K (single wave): cluster = one output tile, rank r owns k-blocks
    // [r*KB/S, (r+1)*KB/S).  Ranks > 0 hand their fp32 partial to the leader through DSMEM.
    // =====================================================================================
    // Rank layout inside the cluster: ksplit = rank % S (K range), g = rank / S (which of the G N-tiles).  The G CTAs with the
    // same ksplit read the SAME activation k-blocks: each loads 1/G of the A rows and multicasts them to the others, so the
    // activations leave L2 once per pair instead of once per CTA (at M = 64 they were half of all L2 -> SM bytes).
    const int G = sk.G;
    const int ksplit = crank % S_split, gidx = crank / S_split;
    const int tile = (blockIdx.x / (S_split * G)) * G + gidx;
    int ta, tb;
    sk.coords(tile, ta, tb);
    const int kb0 = (ksplit * sk.kblocks) / S_split, kb1 = ((ksplit + 1) * sk.kblocks) / S_split;
    uint16_t amask = 0;
    for (int g2 = 0; g2 < G; ++g2) amask |= uint16_t(1u << (ksplit + S_split * g2));
    const int a_rows = BM / G;  // rows of A this CTA loads (and multicasts)
    const int leader_rank = gidx * S_split;
    if (warp == 0) {
      if (ptx::elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        const int npre = (kb1 - kb0) < kStages ? (kb1 - kb0) : kStages;
        for (int i = 0; i < npre; ++i) {  // weights (B operand) first: independent of the previous kernel
          uint8_t* sb = smem + i * S.stage_bytes + S.a_bytes;
          ptx::mbar_arrive_expect_tx(&full_bar[i], S.stage_bytes);
          ptx::tma_load_2d(sb, &tmB, &full_bar[i], (kb0 + i) * BK, tb * BN, ptx::kEvictFirst);
        }
        ptx::grid_dep_wait();
        auto load_a = [&](int st, int kb) {
          uint8_t* sa = smem + st * S.stage_bytes;
          if (G > 1)
            ptx::tma_load_2d_mcast(sa + gidx * a_rows * 128, &tmA, &full_bar[st], kb * BK, ta * BM + gidx * a_rows, amask);
          else
            ptx::tma_load_2d(sa, &tmA, &full_bar[st], kb * BK, ta * BM, ptx::kEvictLast);
        };
        for (int i = 0; i < npre; ++i) load_a(i, kb0 + i);
        stage = npre == kStages ? 0 : npre;
        phase = npre == kStages ? 1 : 0;
        for (int kb = kb0 + npre; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S.stage_bytes;
          uint8_t* sb = sa + S.a_bytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], S.stage_bytes);
          load_a(stage, kb);
          ptx::tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, tb * BN, ptx::kEvictFirst);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (trace && lane == 0 && (kb == kb0 || kb == kb1 - 1 || kb == kb0 + 8)) trace[kb == kb0 ? 1 : (kb == kb1 - 1 ? 3 : 2)] = clock64();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + stage * S.stage_bytes);
          const uint32_t sb = sa + S.a_bytes;
          const uint64_t da = ptx::make_smem_desc(sa, 16, 1024, ptx::kSwz128);
          const uint64_t db = ptx::make_smem_desc(sb, 16, 1024, ptx::kSwz128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            ptx::mma_f16_ss<1>(tmem_base, ptx::desc_advance(da, k * 32), ptx::desc_advance(db, k * 32), idesc,
                               (kb > kb0 || k > 0) ? 1u : 0u);
          if (G > 1) ptx::mma_commit_mcast(&empty_bar[stage], amask);
          else ptx::mma_commit(&empty_bar[stage]);
          if (kb == kb1 - 1) ptx::mma_commit(&tmem_full[0]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    } else if (warp >= 4) {
      const int q = warp - 4;
      const int etid = threadIdx.x - 128;
      const bool row_ok = (BM == 128) || (lane < 16);
      const int r_in_tile = (BM == 128) ? q * 32 + lane : q * 16 + (lane & 15);
      const int a_row = row_ok ? ta * BM + r_in_tile : (1 << 30);
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
      // partial exchange buffer in the LEADER's smem (its pipeline stages are idle by then):
      // [peer-1][16-column chunk][row][16 floats]  (64 B per thread per chunk -> conflict-free)
      const uint32_t xbuf = ptx::smem_u32(smem);
      ptx::mbar_wait(&tmem_full[0], 0);  // my accumulator is complete => all my MMAs (smem reads) retired
      ptx::tc_fence_after();
      if (trace && etid == 0) trace[4] = clock64();
      if (ksplit == 0) {
        if (etid == 0) {
          for (int r = 1; r < S_split; ++r) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(go_bar), leader_rank + r));
        }
        if (S_split > 1) ptx::mbar_wait_cluster(partials_bar, 0);
        for (int c0 = 0; c0 < BN; c0 += 16) {
          uint32_t r[16];
          ptx::tmem_ld_x16(taddr + c0, r);
          ptx::tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          if (row_ok) {
            for (int pr = 0; pr < S_split - 1; ++pr) {
              const float4* src = reinterpret_cast<const float4*>(
                  smem + ((int64_t(pr) * (BN / 16) + c0 / 16) * BM + r_in_tile) * 64);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 x = src[j];
                v[4 * j] += x.x;
                v[4 * j + 1] += x.y;
                v[4 * j + 2] += x.z;
                v[4 * j + 3] += x.w;
              }
            }
          }
          const int b_row0 = tb * BN + c0;
          if (a_row < rowsA) {
            OutT* dst = C + int64_t(a_row) * ldc + b_row0;
            if (b_row0 + 16 <= rowsB && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
              constexpr int VN = 16 / sizeof(OutT);
#pragma unroll
              for (int j = 0; j < 16; j += VN) {
                Vec16<OutT> o;
#pragma unroll
                for (int e = 0; e < VN; ++e) {
                  float x = v[j + e];
                  if (bias) x += to_f32(bias[b_row0 + j + e]);
                  o.v[e] = from_f32<OutT>(x);
                }
                st16(dst + j, o);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (b_row0 + j < rowsB) {
                  float x = v[j];
                  if (bias) x += to_f32(bias[b_row0 + j]);
                  dst[j] = from_f32<OutT>(x);
                }
            }
          }
        }
      } else {
        ptx::mbar_wait_cluster(go_bar, 0);
        const uint32_t remote = ptx::mapa(xbuf, leader_rank);
        for (int c0 = 0; c0 < BN; c0 += 16) {
          uint32_t r[16];
          ptx::tmem_ld_x16(taddr + c0, r);
          ptx::tmem_ld_wait();
          if (row_ok) {
            const uint32_t dst = remote + uint32_t(((int64_t(ksplit - 1) * (BN / 16) + c0 / 16) * BM + r_in_tile) * 64);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ptx::st_dsmem_v4(dst + j * 16, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                         __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
          }
        }
        ptx::named_bar_sync(1, 128);
        if (etid == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(partials_bar), leader_rank));
      }
    }
    if (trace && threadIdx.x == 128) trace[5] = clock64();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();  // nobody leaves while a peer may still touch its shared memory
    if (trace && threadIdx.x == 0) trace[6] = clock64();
    if (warp == 2) {
      ptx::tc_fence_after();
      ptx::tmem_dealloc<1>(tmem_base, tmem_cols);
    }
    return;
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      SegIter it(sk, blockIdx.x);
      int tile, kb0, kb1, r_idx;
      bool part;
      bool first = true;
      while (it.next(tile, kb0, kb1, part, r_idx)) {
        int ta, tb;
        sk.coords(tile, ta, tb);
        int kb_start = kb0;
        if (first) {
          // first segment: weight stages before griddepcontrol.wait (kSwap: weights are the A operand)
          first = false;
          const int npre = (kb1 - kb0) < kStages ? (kb1 - kb0) : kStages;
          for (int i = 0; i < npre; ++i) {
            uint8_t* sa = smem + i * S.stage_bytes;
            ptx::mbar_arrive_expect_tx(&full_bar[i], S.stage_bytes);
            if constexpr (kSwap) ptx::tma_load_2d(sa, &tmA, &full_bar[i], (kb0 + i) * BK, ta * BM, ptx::kEvictFirst);
            else ptx::tma_load_2d(sa + S.a_bytes, &tmB, &full_bar[i], (kb0 + i) * BK, tb * BN, ptx::kEvictNormal);
          }
          ptx::grid_dep_wait();
          for (int i = 0; i < npre; ++i) {
            uint8_t* sa = smem + i * S.stage_bytes;
            if constexpr (kSwap) ptx::tma_load_2d(sa + S.a_bytes, &tmB, &full_bar[i], (kb0 + i) * BK, tb * BN, ptx::kEvictLast);
            else ptx::tma_load_2d(sa, &tmA, &full_bar[i], (kb0 + i) * BK, ta * BM, ptx::kEvictNormal);
          }
          stage = npre == kStages ? 0 : npre;
          phase = npre == kStages ? 1 : 0;
          kb_start = kb0 + npre;
        }
        for (int kb = kb_start; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S.stage_bytes;
          uint8_t* sb = sa + S.a_bytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], S.stage_bytes);
          ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, ta * BM, kSwap ? ptx::kEvictFirst : ptx::kEvictNormal);
          ptx::tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, tb * BN, kSwap ? ptx::kEvictLast : ptx::kEvictNormal);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    SegIter it(sk, blockIdx.x);
    int tile, kb0, kb1, r_idx;
    bool part;
    while (it.next(tile, kb0, kb1, part, r_idx)) {
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + stage * S.stage_bytes);
          const uint32_t sb = sa + S.a_bytes;
          const uint64_t da = ptx::make_smem_desc(sa, 16, 1024, ptx::kSwz128);
          const uint64_t db = ptx::make_smem_desc(sb, 16, 1024, ptx::kSwz128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            ptx::mma_f16_ss<1>(d_tmem, ptx::desc_advance(da, k * 32), ptx::desc_advance(db, k * 32), idesc,
                               (kb > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::mma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (kb == kb1 - 1) ptx::mma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    ptx::grid_dep_launch();
  } else if (warp >= 4) {
    // ===================== epilogue (+ stream-K reduce-scatter fix-up) =====================
    const int q = warp - 4;  // TMEM lane quadrant == warp % 4
    const int etid = threadIdx.x - 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr int CH = 16;
    SegIter it(sk, blockIdx.x);
    int tile, kb0, kb1, r_idx;
    bool part;
    int npend = 0;
    int pend_r[4], pend_parts[4], pend_me[4], pend_ta[4], pend_tb[4];
    while (it.next(tile, kb0, kb1, part, r_idx)) {
      int ta, tb;
      sk.coords(tile, ta, tb);
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      // M=128: row i <-> TMEM lane i.  M=64: rows 16q..16q+15 live in the lower 16 lanes of quadrant q.
      const bool row_ok = (BM == 128) || (lane < 16);
      const int r_in_tile = (BM == 128) ? q * 32 + lane : q * 16 + (lane & 15);
      const int a_row = row_ok ? ta * BM + r_in_tile : (1 << 30);  // row of the A-side operand owned by this thread
      const uint32_t taddr = tmem_base + acc * BN + (uint32_t(q * 32) << 16);

      if (!part) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CH) {
          uint32_t r[CH];
          ptx::tmem_ld_x16(taddr + c0, r);
          ptx::tmem_ld_wait();
          const int b_row0 = tb * BN + c0;
          if constexpr (kSwap) {
            // C[b_row][a_row]: consecutive lanes -> consecutive a_row -> coalesced 2B stores
            if (a_row < rowsA) {
              const float bv = bias ? to_f32(bias[a_row]) : 0.f;
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (b_row0 + j < rowsB) C[int64_t(b_row0 + j) * ldc + a_row] = from_f32<OutT>(__uint_as_float(r[j]) + bv);
            }
          } else if (sk.gated) {
            // fused SwiGLU: weight rows are interleaved (gate_0, up_0, gate_1, up_1, ...), so a 16-column chunk of the
            // accumulator holds 8 (gate, up) pairs -> 8 outputs = one 16-byte store; the [M, 2I] intermediate never exists
            if (a_row < rowsA && b_row0 < rowsB) {
              OutT* dst = C + int64_t(a_row) * ldc + b_row0 / 2;
              OutT o8[CH / 2];
#pragma unroll
              for (int j = 0; j < CH; j += 2) {
                const float g = __uint_as_float(r[j]), u = __uint_as_float(r[j + 1]);
                o8[j / 2] = from_f32<OutT>(g / (1.f + __expf(-g)) * u);
              }
              if (b_row0 + CH <= rowsB && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                *reinterpret_cast<int4*>(dst) = *reinterpret_cast<const int4*>(o8);
              } else {
#pragma unroll
                for (int j = 0; j < CH / 2; ++j)
                  if (b_row0 + 2 * j + 1 < rowsB) dst[j] = o8[j];
              }
            }
          } else {
            if (a_row < rowsA) {
              OutT* dst = C + int64_t(a_row) * ldc + b_row0;
              if (b_row0 + CH <= rowsB && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                constexpr int VN = 16 / sizeof(OutT);
#pragma unroll
                for (int j = 0; j < CH; j += VN) {
                  Vec16<OutT> o;
#pragma unroll
                  for (int e = 0; e < VN; ++e) {
                    float x = __uint_as_float(r[j + e]);
                    if (bias) x += to_f32(bias[b_row0 + j + e]);
                    o.v[e] = from_f32<OutT>(x);
                  }
                  st16(dst + j, o);
                }
              } else {
#pragma unroll
                for (int j = 0; j < CH; ++j)
                  if (b_row0 + j < rowsB) {
                    float x = __uint_as_float(r[j]);
                    if (bias) x += to_f32(bias[b_row0 + j]);
                    dst[j] = from_f32<OutT>(x);
                  }
              }
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      } else {
        // ---- stream-K part: publish the fp32 partial now, fix up later (after ALL my partials are out,
        //      otherwise neighbouring CTAs would wait on each other in a chain) ----
        const int64_t t0u = int64_t(r_idx) * sk.kblocks;
        const int c_first = sk.sk_cta_of(t0u);
        const int c_last = sk.sk_cta_of(t0u + sk.kblocks - 1);
        const int parts = c_last - c_first + 1;
        const int my_part = blockIdx.x - c_first;
        float* tile_ws = partial + int64_t(r_idx) * sk.max_parts * (BM * BN);
        float* my_slot = tile_ws + int64_t(my_part) * (BM * BN);
        // publish in output-major order (swap: [col][row], else [row][col])
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CH) {
          uint32_t r[CH];
          ptx::tmem_ld_x16(taddr + c0, r);
          ptx::tmem_ld_wait();
          if (!row_ok) continue;
          if constexpr (kSwap) {
#pragma unroll
            for (int j = 0; j < CH; ++j) __stcg(my_slot + (c0 + j) * BM + r_in_tile, __uint_as_float(r[j]));
          } else {
            float4* dst = reinterpret_cast<float4*>(my_slot + r_in_tile * BN + c0);
#pragma unroll
            for (int j = 0; j < CH; j += 4)
              __stcg(dst + j / 4, make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                              __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
        __threadfence();
        ptx::named_bar_sync(1, 128);
        if (etid == 0) atomicAdd(counters + 2 * r_idx, 1);
        pend_r[npend] = r_idx;
        pend_parts[npend] = parts;
        pend_me[npend] = my_part;
        pend_ta[npend] = ta;
        pend_tb[npend] = tb;
        ++npend;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;

      // ---- deferred fix-ups: once this CTA has published every stream-K partial it owns ----
      if (npend > 0 && it.u >= it.u_end) {
        for (int pi = 0; pi < npend; ++pi) {
          const int fr = pend_r[pi], parts = pend_parts[pi], my_part = pend_me[pi];
          const int fta = pend_ta[pi], ftb = pend_tb[pi];
          int* arrive = counters + 2 * fr;
          int* done = arrive + 1;
          float* tile_ws = partial + int64_t(fr) * sk.max_parts * (BM * BN);
          if (etid == 0) {
            while (*reinterpret_cast<volatile int*>(arrive) < parts) {
            }
            __threadfence();
          }
          ptx::named_bar_sync(1, 128);
          // reduce my 1/parts share of the tile (slot order => deterministic) and write it out
          const int NV = BM * BN / 4;
          const int v_begin = (my_part * NV) / parts, v_end = ((my_part + 1) * NV) / parts;
          for (int v = v_begin + etid; v < v_end; v += 128) {
            float4 accv = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int pth = 0; pth < parts; ++pth) {
              const float4 x = __ldcg(reinterpret_cast<const float4*>(tile_ws + int64_t(pth) * (BM * BN)) + v);
              accv.x += x.x;
              accv.y += x.y;
              accv.z += x.z;
              accv.w += x.w;
            }
            const float vals[4] = {accv.x, accv.y, accv.z, accv.w};
            if constexpr (kSwap) {
              const int col = (v * 4) / BM, row = (v * 4) % BM;  // 4 consecutive a-rows of one token column
              const int b_row = ftb * BN + col;
              const int a0 = fta * BM + row;
              if (b_row < rowsB) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (a0 + e < rowsA)
                    C[int64_t(b_row) * ldc + a0 + e] = from_f32<OutT>(vals[e] + (bias ? to_f32(bias[a0 + e]) : 0.f));
              }
            } else {
              const int row = (v * 4) / BN, col = (v * 4) % BN;
              const int ar = fta * BM + row;
              const int b0 = ftb * BN + col;
              if (ar < rowsA) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (b0 + e < rowsB)
                    C[int64_t(ar) * ldc + b0 + e] = from_f32<OutT>(vals[e] + (bias ? to_f32(bias[b0 + e]) : 0.f));
              }
            }
          }
          // the last part to finish resets the counters (graph-replay safe)
          ptx::named_bar_sync(1, 128);
          if (etid == 0) {
            const int old = atomicAdd(done, 1);
            if (old == parts - 1) {
              *reinterpret_cast<volatile int*>(done) = 0;
              __threadfence();
              *reinterpret_cast<volatile int*>(arrive) = 0;
            }
          }
        }
        npend = 0;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, tmem_cols);
  }
}

inline int env_int(const char* name, int dflt);

template <int BM, bool kSwap, typename OutT>
int launch_gemm(int BN, int cluster_split, int mcast_g, int gated, const CUtensorMap& tmA, const CUtensorMap& tmB, OutT* C, float* workspace, int64_t workspace_bytes,
                const OutT* bias, int rowsA, int rowsB, int K, int64_t ldc, bool f16, bool pdl, cudaStream_t stream) {
  // small-M cluster split-K launches (decode GEMMs, one tile per cluster, a few us long): half-size ring so that two CTAs fit
  // an SM and consecutive PDL-chained GEMMs overlap their prologue / weight prefetch with the predecessor's epilogue
  static const int small_kb = env_int("FIB200_GEMM_SMALL_SMEM_KB", 0);  // measured neutral on B200 (o_proj 16.0 -> 15.2 us, qkv 19.5 -> 20.0): opt-in
  const bool co_resident = !kSwap && cluster_split > 1 && small_kb > 0 && small_kb < 220;
  GemmSmem S = GemmSmem::make(BM, BN, co_resident ? small_kb : 220);
  auto kern = gemm_nt_kernel<BM, kSwap, OutT>;
  static bool attr_set = false;
  if (!attr_set) {
    FIB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  Sched sk;
  sk.gated = gated;
  sk.smem_kb = co_resident ? small_kb : 220;
  sk.BN = BN;
  sk.tiles_a = (rowsA + BM - 1) / BM;
  sk.tiles_b = (rowsB + BN - 1) / BN;
  sk.kblocks = (K + BK - 1) / BK;
  const int tiles = sk.tiles_a * sk.tiles_b;
  int grid = num_sms();
  if (g_sm_budget > 0 && g_sm_budget < grid) grid = g_sm_budget;  // SM-constrained GEMM: the persistent grid is the budget
  if (tiles < grid && (!kSwap || workspace == nullptr || int64_t(tiles) * sk.kblocks < 2 * grid)) grid = tiles;
  sk.grid = grid;
  sk.W = tiles / grid;
  sk.R = tiles % grid;
  sk.group_a = sk.tiles_a < 8 ? sk.tiles_a : 8;
  sk.g_sk = 0;
  sk.max_parts = 1;
  sk.u_r = int64_t(sk.R) * sk.kblocks;
  // workspace layout: [counters: 2 ints per remainder tile, 4 KB][fp32 partial slots]
  const int64_t slot_bytes = int64_t(BM) * BN * 4;
  if (workspace != nullptr && sk.R > 0 && sk.R <= 512 && (kSwap || sk.W >= 1)) {
    int64_t mp = (workspace_bytes - 4096) / (int64_t(sk.R) * slot_bytes);
    if (mp > 8) mp = 8;
    if (mp >= 2) {
      int64_t g = int64_t(sk.R) * (mp - 1);       // every tile spans at most mp CTAs
      if (g > grid) g = grid;
      if (g > sk.u_r / 2) g = sk.u_r / 2;          // at least 2 k-blocks per CTA
      if (g > sk.R) {                              // otherwise plain DP is just as good
        sk.g_sk = (int)g;
        const int64_t per = sk.u_r / g;
        sk.max_parts = int((sk.kblocks + per - 1) / per) + 1;
        if (sk.max_parts > mp) sk.max_parts = (int)mp;
      }
    }
  }
  static const int trace_env = env_int("FIB200_GEMM_TRACE", 0);
  sk.trace = trace_env;
  sk.S = 1;
  sk.G = 1;
  if (!kSwap && (cluster_split > 1 || mcast_g > 1) &&
      tiles * cluster_split <= (cluster_split * mcast_g >= 4 ? (num_sms() * 132) / 148 : num_sms()) && tiles % mcast_g == 0 &&
      sk.kblocks >= 4 * cluster_split &&
      int64_t(cluster_split - 1) * BM * BN * 4 <= int64_t(S.stages) * S.stage_bytes) {
    sk.S = cluster_split;
    sk.G = mcast_g;
    sk.grid = grid = tiles * cluster_split;
    sk.W = 1;
    sk.R = 0;
    sk.g_sk = 0;
  }
  if (sk.S == 1 && co_resident) {  // the cluster split was rejected: plain persistent kernel with the full ring
    sk.smem_kb = 220;
    S = GemmSmem::make(BM, BN, 220);
  }
  const uint32_t idesc = ptx::make_idesc_f16(f16 ? ptx::kFmtF16 : ptx::kFmtBF16, BM, BN, 0, 0);
  if (sk.G != mcast_g) return set_error("gemm: A-multicast was planned but the cluster launch was rejected");
  LaunchCfg lc(dim3(grid), dim3(256), S.total, stream, pdl, sk.S * sk.G);
  int* counters = reinterpret_cast<int*>(workspace);
  float* partial = workspace ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 4096) : nullptr;
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, tmA, tmB, C, partial, counters, bias, rowsA, rowsB, K, ldc, sk, idesc));
  return 0;
}

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <typename OutT>
int gemm_dispatch(const void* A, const void* B, OutT* C, const OutT* bias, int M, int N, int K, int64_t lda, int64_t ldb,
                  int64_t ldc, bool f16, float* workspace, int64_t workspace_bytes, bool pdl, cudaStream_t stream,
                  bool gated = false) {
  if (gated) {
    // the gated epilogue lives in the data-parallel path only: no stream-K fix-up (workspace off), no cluster split, no swap
    FIB_CHECK(bias == nullptr && N % 16 == 0, "gemm (gated): no bias, N must be a multiple of 16");
    workspace = nullptr;
    workspace_bytes = 0;
  }
  const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  // Small M (decode): activations take the 128-row MMA-M side (rows past M are TMA zero-fill, no traffic) and
  // the weight matrix is cut into narrow N tiles so that ONE wave covers the machine with no split-K:
  // BN = smallest multiple of 16 with ceil(N / BN) <= #SM (>= 64 to keep the L2 re-read of the activations
  // below the HBM traffic).  Large M: 128x256 / 128x128 tiles, round-robin waves + stream-K remainder.
  static const int force_swap = env_int("FIB200_GEMM_SWAP", -1);
  static const int force_bn = env_int("FIB200_GEMM_BN", 0);
  static const int min_bn_small = env_int("FIB200_GEMM_MIN_BN", 64);
  static const int force_bm = env_int("FIB200_GEMM_BM", 0);
  bool swap = false;
  int BN;
  int BMsel = 128;
  static const int force_s = env_int("FIB200_GEMM_CLUSTER", 0);
  int Ssel = 1;
  if (M <= 128) {
    const int sms = num_sms();
    if (M <= 64) BMsel = 64;  // half-height MMA tile: the activation tile costs 8 KB/stage -> deeper ring
    BN = ((N + sms - 1) / sms + 15) / 16 * 16;
    const int min_bn = (BMsel == 64 && min_bn_small == 64) ? 32 : min_bn_small;
    if (BN < min_bn) BN = min_bn;
    // Few, wide N tiles + K split over a 4-CTA cluster (DSMEM reduction): the activation tile is re-read
    // from L2 once per N tile, so wide tiles cut L2 traffic 4x while the cluster keeps every SM streaming.
    if (BN <= 64 && K >= 2048 && !gated) {  // (the gated epilogue has no split-K path: it keeps the narrow one-wave tiles)
      const int s_try = force_s > 0 ? force_s : 2;
      // clusters of 4 can only be placed on 132 of the 148 SMs (GPC granularity), clusters of 2 on all
      const int eff = s_try >= 4 ? (sms * 132) / 148 : sms;
      int bn4 = ((N * s_try + eff - 1) / eff + 15) / 16 * 16;
      if (bn4 < 64) bn4 = 64;
      if (bn4 <= 256 && s_try > 1) {
        BN = bn4;
        Ssel = s_try;
      }
    }
    if (BN > 256) {
      // many waves: pick BN in [192, 256] minimising the tail of the last wave
      int best = 256;
      double best_eff = 0.0;
      for (int bn = 256; bn >= 192; bn -= 16) {
        const int t = (N + bn - 1) / bn;
        const double eff = double(t) / (double((t + sms - 1) / sms) * sms);
        if (eff > best_eff + 1e-9) {
          best_eff = eff;
          best = bn;
        }
      }
      BN = best;
    }
  } else {
    BN = (N >= 256 && (int64_t(M) * N >= int64_t(256) * 256 * 64)) ? 256 : 128;
  }
  if (gated) Ssel = 1;
  if (g_sm_budget > 0) Ssel = 1;  // SM-constrained: plain persistent tiles (cluster split-K assumes one tile per cluster on the full machine)
  if (force_swap == 1 && M <= 128 && !gated) {
    swap = true;
    BMsel = 128;
    BN = M <= 16 ? 16 : (M <= 32 ? 32 : (M <= 64 ? 64 : 128));
  }
  if (force_bn > 0) BN = force_bn;
  if (force_bm > 0 && !swap) BMsel = force_bm;
  const int BM = BMsel;
  // Cluster split-K shapes: pair up neighbouring N tiles so that the activation k-blocks are multicast (cluster = S x 2).
  // Clusters of 4 only fit 132 of the 148 SMs, and the pairing needs an even number of N tiles and a single M tile.
  // measured on B200 (o_proj shape, clock64 trace): the main loop is DRAM-bound at 5.6 TB/s with or without the multicast,
  // so it stays opt-in (FIB200_GEMM_MCAST=2)
  static const int mcast_env = env_int("FIB200_GEMM_MCAST", 1);
  int Gsel = 1;
  if (!swap && Ssel == 2 && mcast_env == 2 && M <= BM) {
    const int tb = (N + BN - 1) / BN;
    if (tb % 2 == 0 && tb * Ssel <= 132) Gsel = 2;
  }
  // FIB200_GEMM_MCAST=4: no K split at all - narrow N tiles with full K, the activation k-blocks multicast over clusters of
  // four neighbouring tiles (each CTA loads a quarter of the rows): no partial exchange in the epilogue
  if (!swap && mcast_env == 4 && M <= 64 && Ssel == 2 && force_bn == 0) {
    int bn = ((N + 131) / 132 + 15) / 16 * 16;
    if (bn < 32) bn = 32;
    const int tb = (N + bn - 1) / bn;
    if (tb % 4 == 0 && tb <= 132 && bn <= 128) {
      BN = bn;
      Ssel = 1;
      Gsel = 4;
    }
  }
  // A-side = 128-row operand.  normal: activations; swap: weights.
  const void* pa = swap ? B : A;
  const void* pb = swap ? A : B;
  const int rowsA = swap ? N : M, rowsB = swap ? M : N;
  const int64_t ldA = swap ? ldb : lda, ldB = swap ? lda : ldb;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)rowsA};
    uint64_t str[1] = {(uint64_t)ldA * 2};
    uint32_t box[2] = {BK, (uint32_t)(BM / Gsel)};  // multicast: every CTA of a pair loads half of the rows
    if (make_tmap(&tmA, dt, 2, pa, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)rowsB};
    uint64_t str[1] = {(uint64_t)ldB * 2};
    uint32_t box[2] = {BK, (uint32_t)BN};
    if (make_tmap(&tmB, dt, 2, pb, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  if (swap)
    return launch_gemm<128, true, OutT>(BN, 1, 1, 0, tmA, tmB, C, workspace, workspace_bytes, bias, rowsA, rowsB, K, ldc, f16, pdl,
                                        stream);
  if (BM == 64)
    return launch_gemm<64, false, OutT>(BN, Ssel, Gsel, gated ? 1 : 0, tmA, tmB, C, workspace, workspace_bytes, bias, rowsA, rowsB, K, ldc, f16, pdl,
                                        stream);
  return launch_gemm<128, false, OutT>(BN, Ssel, Gsel, gated ? 1 : 0, tmA, tmB, C, workspace, workspace_bytes, bias, rowsA, rowsB, K, ldc, f16, pdl,
                                       stream);
}

}  // namespace

// C[M,N] (row-major, ldc) = A[M,K] (lda) * B[N,K]^T (ldb) + bias[N]; dtype: 0 = f16, 1 = bf16.
extern "C" int gemm_nt(void* A, void* B, void* C, void* bias, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                       int64_t ldc, int64_t dtype, void* workspace, int64_t workspace_bytes, int64_t pdl,
                       int64_t stream) {
  FIB_CHECK(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "K/lda/ldb must be multiples of 8 (16B TMA alignment)");
  FIB_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
            "A/B must be 16B aligned");
  if (M == 0 || N == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == kF16) {
    return gemm_dispatch<__half>(A, B, (__half*)C, (const __half*)bias, (int)M, (int)N, (int)K, lda, ldb, ldc, true,
                                 (float*)workspace, workspace_bytes, pdl != 0, s);
  } else if (dtype == kBF16) {
    return gemm_dispatch<__nv_bfloat16>(A, B, (__nv_bfloat16*)C, (const __nv_bfloat16*)bias, (int)M, (int)N, (int)K, lda,
                                        ldb, ldc, false, (float*)workspace, workspace_bytes, pdl != 0, s);
  }
  return set_error("gemm_nt: unsupported dtype");
}

// C[M, N/2] = silu(A W_even^T) * (A W_odd^T): W rows interleaved (gate_0, up_0, gate_1, up_1, ...); the SwiGLU runs in the
// GEMM epilogue on the fp32 accumulators (reference: gated-activation GEMM epilogues of the trtllm-gen MoE / MLP kernels).
extern "C" int gemm_nt_gated_silu(void* A, void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                                  int64_t dtype, int64_t pdl, int64_t stream) {
  FIB_CHECK(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0, "K/lda/ldb/ldc must be multiples of 8");
  if (M == 0 || N == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == kF16)
    return gemm_dispatch<__half>(A, B, (__half*)C, nullptr, (int)M, (int)N, (int)K, lda, ldb, ldc, true, nullptr, 0, pdl != 0, s, true);
  if (dtype == kBF16)
    return gemm_dispatch<__nv_bfloat16>(A, B, (__nv_bfloat16*)C, nullptr, (int)M, (int)N, (int)K, lda, ldb, ldc, false, nullptr, 0,
                                        pdl != 0, s, true);
  return set_error("gemm_nt_gated_silu: unsupported dtype");
}

// SM-constrained GEMM (reference flashinfer/triton/sm_constraint_gemm.py gemm_persistent(num_sms=...)): the persistent tcgen05
// kernel runs with `n` CTAs (one per SM) until the budget is reset to 0, leaving the other SMs to a concurrent kernel
// (communication, another stream's GEMM).  Thread-local, like the launch itself.
extern "C" int gemm_set_sm_budget(int64_t n) {
  FIB_CHECK(n >= 0, "gemm_set_sm_budget: n >= 0 (0 = no constraint)");
  g_sm_budget = int(n);
  return 0;
}
