// Multi-head Latent Attention (DeepSeek MLA) decode over a paged latent cache, sm_100a tcgen05/TMEM/TMA.
//
// Parity: reference BatchMLAPagedAttentionWrapper (flashinfer/mla/_core.py:219-629), kernels
// include/flashinfer/attention/mla.cuh (fa2, mma.sync), the CUTLASS sm100 MLA kernel
// (include/flashinfer/attention/blackwell/kernel/sm100_fmha_mla_tma_warpspecialized.hpp) and
// trtllm_batch_decode_with_kv_cache_mla.
//
// Shapes: q_nope [n, H<=128, 512], q_pe [n, H, 64]; ckv cache [pages, page, 512]; kpe cache [pages, page, 64];
// S = q_nope.ckv^T + q_pe.kpe^T ; O = softmax(S).ckv  -> [n, H, 512].
//
// B200-first design:
//  * all (<=128) heads of one query token form the MMA-M dimension (MQA: every head shares the latent KV).
//  * O (128 x 512 fp32) would fill all 512 TMEM columns, so a work item runs on a CLUSTER OF TWO CTAs that split the
//    576-wide latent dimension: CTA 0 owns ckv[0:256), CTA 1 owns ckv[256:512) + kpe.  Each CTA keeps only ITS slice of Q
//    resident (64 / 80 KB instead of 144 KB), streams only ITS slice of the latent cache (half the bytes, which leaves room
//    for a 5-deep TMA ring of 32-token tiles), computes a PARTIAL S over its slice of the reduction dimension, and the two
//    partials are exchanged through distributed shared memory (st.shared::cluster + a remote mbarrier arrive per thread)
//    and summed; both CTAs then run the same softmax and each accumulates O for its own 256 d_v columns - which are exactly
//    the ckv columns it already holds (the same smem tile is the K-major B operand of QK^T and the MN-major B of P.V).
//    No FLOP and no cache byte is duplicated across the pair.
//  * TMEM map: S0 | S1 (32 cols each, P is written over S as the TMEM A operand of P.V) | O (256 cols); O rescaled lazily.
//  * split-KV: grid = 2 x sum of per-request chunk counts; partial (o, lse) are folded by mla_merge_kernel (PDL-chained in
//    the same host call, all SMs).
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>
#include <type_traits>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

constexpr int kHeads = 128;   // MMA M
constexpr int kCkv = 512, kKpe = 64, kDqk = kCkv + kKpe;
constexpr int kTile = 32;     // kv tokens per tile
constexpr int kDvHalf = 256;
constexpr int kChunksQK = kDqk / 64;  // 9
constexpr int kWorkInts = 8;

struct MlaParams {
  const int32_t* kv_indices;
  const int32_t* work;  // [nwork][8] {q_row, kv_page_start, kv_begin(token), kv_end(token), kv_len, out_slot, num_pages, -}
  void* out;            // final bf16/f16 [n, H, 512]            (when partial == nullptr)
  float* partial_o;     // [slots][H][512] fp32                  (split-KV)
  float* partial_lse;   // [slots][H]
  float* lse;           // optional final lse [n, H]
  int num_heads, page_size;
  int64_t o_stride_n, o_stride_h;
  float sm_scale_log2;
};

struct Smem {
  static constexpr int kStages = 5;
  static constexpr int kSBufs = 4;   // S / P TMEM buffers: QK^T runs two tiles ahead of the softmax
  static constexpr int kXBufs = 3;   // partial-S exchange buffers (see the protocol note in the softmax loop)
  static constexpr int kMaxChunks = 5;                         // CTA 1: 4 ckv chunks + kpe
  static constexpr int kQBytes = kMaxChunks * kHeads * 128;    // 81920
  static constexpr int kChunkBytes = kTile * 128;              // 4096
  static constexpr int kTileBytes = kMaxChunks * kChunkBytes;  // 20480
  static constexpr int kXchgBytes = kHeads * kTile * 2;        // 8192: the peer's partial S (pre-scaled, fp16)
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = kQBytes;
  static constexpr int kOffX = kOffK + kStages * kTileBytes;
  static constexpr int kOffBar = kOffX + kXBufs * kXchgBytes;
  static constexpr int kNumBars = 2 * kStages + 1 /*q_full*/ + kSBufs /*s_full*/ + kSBufs /*p_ready*/ + 1 /*o_done*/ + kXBufs;
  static constexpr int kTotal = kOffBar + kNumBars * 8 + 16 + 1024;
};
static_assert(Smem::kTotal <= 227 * 1024, "MLA smem budget");

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack2_f16_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack2_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
mla_decode_kernel(const __grid_constant__ CUtensorMap tmQn, const __grid_constant__ CUtensorMap tmQp,
                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP, const MlaParams p,
                  uint32_t idesc_qk, uint32_t idesc_pv) {
  using S = Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* k_full = bars;
  uint64_t* k_empty = k_full + S::kStages;
  uint64_t* q_full = k_empty + S::kStages;
  uint64_t* s_full = q_full + 1;            // [kSBufs]
  uint64_t* p_ready = s_full + S::kSBufs;   // [kSBufs]
  uint64_t* o_done = p_ready + S::kSBufs;   // [1]
  uint64_t* x_full = o_done + 1;            // [kXBufs] peer's partial S landed in my exchange buffer (tx-count: 16 KB of st.async)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(x_full + S::kXBufs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = int(ptx::cluster_ctarank());  // 0: ckv[0:256)   1: ckv[256:512) + kpe   (also: which d_v half)
  const int nch = half == 0 ? 4 : 5;             // 64-column chunks of the latent dimension this CTA owns
  const int32_t* wk = p.work + (blockIdx.x >> 1) * kWorkInts;
  const int q_row = wk[0], page_start = wk[1], kv_begin = wk[2], kv_end = wk[3], out_slot = wk[5], num_pages = wk[6];
  const int ntiles = (kv_end - kv_begin + kTile - 1) / kTile;
  const int ps = p.page_size;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmQn);
    ptx::prefetch_tmap(&tmQp);
    ptx::prefetch_tmap(&tmC);
    ptx::prefetch_tmap(&tmP);
    for (int i = 0; i < S::kStages; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&k_empty[i], 1);
    }
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < S::kSBufs; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_ready[i], 128);
    }
    for (int i = 0; i < S::kXBufs; ++i) {
      ptx::mbar_init(&x_full[i], 1);
      if (i < ntiles) ptx::mbar_arrive_expect_tx(&x_full[i], S::kXchgBytes);  // first use of every exchange buffer
    }
    ptx::mbar_init(o_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_ptr, 512);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();  // barriers of both CTAs are initialised before any remote arrive (also a CTA-wide barrier)
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tm_s = tmem_base;                        // S buffer i at +32 * i
  const uint32_t tm_o = tmem_base + 32 * S::kSBufs;       // 256 columns

  ptx::grid_dep_wait();

  if (warp == 2) {
    // ============================ Q loader (once) ============================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, nch * kHeads * 128);
      for (int lc = 0; lc < nch; ++lc) {
        const int gc = half * 4 + lc;  // chunk of the 576-wide latent dimension
        if (gc < 8)
          ptx::tma_load_3d(smem + S::kOffQ + lc * (kHeads * 128), &tmQn, q_full, gc * 64, 0, q_row, ptx::kEvictFirst);
        else
          ptx::tma_load_3d(smem + S::kOffQ + lc * (kHeads * 128), &tmQp, q_full, 0, 0, q_row, ptx::kEvictFirst);
      }
    }
  } else if (warp == 0 || warp == 3) {
    // ============================ latent-cache producers ============================
    // warp 0: local chunks 0..1, warp 3: the rest.  One TMA box = (rows of one page) x 64 columns.
    const int c_lo = (warp == 0) ? 0 : 2, c_hi = (warp == 0) ? 2 : nch;
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < ntiles; ++j) {
      const int tok0 = kv_begin + j * kTile;
      // boxes of this tile: page-granular pieces (page_size >= 32: one piece; smaller pages: several)
      const int rows_per_box = ps < kTile ? ps : kTile;
      const int nbox = kTile / rows_per_box;
      if (lane == 0) {
        ptx::mbar_wait(&k_empty[st], ph ^ 1);
        if (warp == 0) ptx::mbar_arrive_expect_tx(&k_full[st], nch * S::kChunkBytes);
      }
      __syncwarp();
      if (lane < nbox) {
        const int tok = tok0 + lane * rows_per_box;
        const int pidx = tok / ps;
        // pages past the request's list: clamp (rows are masked in softmax, zero-filled V not required
        // because P is exactly 0 there and the cache memory is finite by contract of the clamp)
        const int page = __ldg(p.kv_indices + page_start + min(pidx, num_pages - 1));
        const int off = tok % ps;
        uint8_t* dst = smem + S::kOffK + st * S::kTileBytes + lane * rows_per_box * 128;
        for (int c = c_lo; c < c_hi; ++c) {
          const int gc = half * 4 + c;
          if (gc < 8)
            ptx::tma_load_3d(dst + c * S::kChunkBytes, &tmC, &k_full[st], gc * 64, off, page, ptx::kEvictFirst);
          else
            ptx::tma_load_3d(dst + c * S::kChunkBytes, &tmP, &k_full[st], 0, off, page, ptx::kEvictFirst);
        }
      }
      if (++st == S::kStages) {
        st = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    int st = 0;
    uint32_t ph = 0;
    const uint32_t q_addr = ptx::smem_u32(smem + S::kOffQ);
    ptx::mbar_wait(q_full, 0);
    auto issue_qk = [&](int j) {
      const uint32_t b = j & (S::kSBufs - 1);
      ptx::mbar_wait(&k_full[st], ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t k_addr = ptx::smem_u32(smem + S::kOffK + st * S::kTileBytes);
        for (int k = 0; k < nch * 4; ++k) {
          const int c = k / 4, o = (k % 4) * 32;
          const uint64_t da = ptx::make_smem_desc(q_addr + c * (kHeads * 128) + o, 16, 1024, ptx::kSwz128);
          const uint64_t db = ptx::make_smem_desc(k_addr + c * S::kChunkBytes + o, 16, 1024, ptx::kSwz128);
          ptx::mma_f16_ss<1>(tm_s + b * 32, da, db, idesc_qk, k > 0 ? 1u : 0u);
        }
        ptx::mma_commit(&s_full[b]);
      }
      __syncwarp();
      if (++st == S::kStages) {
        st = 0;
        ph ^= 1;
      }
    };
    auto issue_pv = [&](int j) {
      const uint32_t b = j & (S::kSBufs - 1);
      const int stage = j % S::kStages;
      ptx::mbar_wait(&p_ready[b], (j / S::kSBufs) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        // B = ckv[kv rows, my 256 dv columns]: MN-major SW128; chunk stride (LBO) = 4096, 8-row group (SBO) = 1024
        const uint32_t v_addr = ptx::smem_u32(smem + S::kOffK + stage * S::kTileBytes);  // local chunks 0..3 = my d_v half
        const uint64_t db = ptx::make_smem_desc(v_addr, S::kChunkBytes, 1024, ptx::kSwz128);
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k)
          ptx::mma_f16_ts<1>(tm_o, tm_s + b * 32 + k * 8, ptx::desc_advance(db, k * 16 * 128), idesc_pv,
                             (j == 0 && k == 0) ? 0u : 1u);
        ptx::mma_commit(&k_empty[stage]);
        ptx::mma_commit(o_done);
      }
      __syncwarp();
    };
    // QK^T runs two tiles ahead of P.V: S(j + 1) is complete (and its partial already on its way to the peer) while the
    // softmax of tile j runs.  S buffer (j + 2) & 3 was last read by P.V(j - 2), issued earlier on the same in-order pipe.
    if (ntiles > 0) issue_qk(0);
    if (ntiles > 1) issue_qk(1);
    for (int j = 0; j < ntiles; ++j) {
      if (j + 2 < ntiles) issue_qk(j + 2);
      issue_pv(j);
    }
  } else if (warp >= 4) {
    // ============================ softmax + epilogue ============================
    const int q4 = warp - 4;
    const int row = q4 * 32 + lane;  // head index
    const uint32_t lane_addr = uint32_t(q4 * 32) << 16;
    constexpr bool kIsBf16 = std::is_same<T, __nv_bfloat16>::value;
    float m_used = -INFINITY, l = 0.f;
    uint32_t od_cnt = 0;
    // ---- partial-S exchange with the peer CTA (it reduced over the other slice of the latent dimension) ----
    // DSMEM moves ~17 B/clk per SM, so the exchange is the scarce resource of this design: partials are pre-scaled by
    // sm_scale * log2(e) and shipped as fp16 (64 B per row and tile; my own partial stays fp32, so a logit carries one fp16
    // rounding of HALF of its value - below the bf16 rounding P gets anyway).  The partial of tile j + 1 is pushed with
    // st.async (16-byte stores that credit the PEER's x_full barrier: no release fence, no per-thread arrive) right after
    // the peer's partial of tile j has landed, so it travels while the softmax of tile j runs.
    // Layout: row-major [128][64 B], the four 16-byte chunks of a row XOR-swizzled by (row >> 1) & 3 (conflict-free LDS.128).
    // Buffer reuse (3 buffers, no extra handshake): when I push tile j + 1 I have received the peer's tile j, which it
    // pushed after finishing its iteration j - 2 -> it is done reading buffer (j - 2) % 3 == (j + 1) % 3.
    const uint32_t peer = uint32_t(half ^ 1);
    const int xsw = (row >> 1) & 3;
    auto push = [&](int jn, const uint32_t(&v)[32]) {
      const int xb = jn % S::kXBufs;
      const uint32_t x_remote = ptx::mapa(ptx::smem_u32(smem + S::kOffX + xb * S::kXchgBytes + row * 64), peer);
      const uint32_t bar_remote = ptx::mapa(ptx::smem_u32(&x_full[xb]), peer);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          h[e] = pack2_f16_sat(__uint_as_float(v[8 * q + 2 * e]), __uint_as_float(v[8 * q + 2 * e + 1]));
        ptx::st_async_v4(x_remote + ((q ^ xsw) << 4),
                         make_float4(__uint_as_float(h[0]), __uint_as_float(h[1]), __uint_as_float(h[2]), __uint_as_float(h[3])),
                         bar_remote);
      }
    };
    auto scale32 = [&](uint32_t(&v)[32]) {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * p.sm_scale_log2);
    };
    uint32_t r[32];
    if (ntiles > 0) {
      ptx::mbar_wait(&s_full[0], 0);
      ptx::tc_fence_after();
      ptx::tmem_ld_x32(tm_s + lane_addr, r);
      ptx::tmem_ld_wait();
      scale32(r);
      push(0, r);
    }
    for (int j = 0; j < ntiles; ++j) {
      const uint32_t b = j & (S::kSBufs - 1);
      const int tok0 = kv_begin + j * kTile;
      const int xb = j % S::kXBufs;
      ptx::mbar_wait(&x_full[xb], (j / S::kXBufs) & 1);  // tx-count completion (like a TMA write): plain acquire.cta
      {
        const int4* xl = reinterpret_cast<const int4*>(smem + S::kOffX + xb * S::kXchgBytes + row * 64);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int4 t = xl[q ^ xsw];
          const uint32_t w[4] = {uint32_t(t.x), uint32_t(t.y), uint32_t(t.z), uint32_t(t.w)};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
            r[8 * q + 2 * e] = __float_as_uint(__uint_as_float(r[8 * q + 2 * e]) + f.x);
            r[8 * q + 2 * e + 1] = __float_as_uint(__uint_as_float(r[8 * q + 2 * e + 1]) + f.y);
          }
        }
      }
      uint32_t rn[32];
      if (j + 1 < ntiles) {
        const uint32_t bn = (j + 1) & (S::kSBufs - 1);
        ptx::mbar_wait(&s_full[bn], ((j + 1) / S::kSBufs) & 1);
        ptx::tc_fence_after();
        ptx::tmem_ld_x32(tm_s + lane_addr + bn * 32, rn);
        ptx::tmem_ld_wait();
        scale32(rn);
        push(j + 1, rn);
      }
      // re-arm this exchange buffer for tile j + 3 (one thread; the 128 waits above have not necessarily all passed, but
      // the next phase cannot complete before the peer's tile-(j + 3) bytes arrive, three tiles from now)
      if (threadIdx.x == 128 && j + S::kXBufs < ntiles) ptx::mbar_arrive_expect_tx(&x_full[xb], S::kXchgBytes);
      const int valid = kv_end - tok0;  // columns >= valid are past the chunk
      float tmax = -INFINITY;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float x = (e < valid) ? __uint_as_float(r[e]) : -INFINITY;
        tmax = fmaxf(tmax, x);
      }
      const float m_tile = tmax;  // logits are already in the scaled log2 domain
      const bool grow = (m_tile > m_used + 8.f) || (m_used == -INFINITY && m_tile > -INFINITY);
      if (j > 0 && __any_sync(0xffffffffu, grow && l > 0.f)) {
        ptx::mbar_wait(o_done, (od_cnt - 1) & 1);
        ptx::tc_fence_after();
        const float alpha = (grow && m_used > -INFINITY) ? ptx::ex2(m_used - m_tile) : 1.f;
#pragma unroll 1
        for (int c = 0; c < kDvHalf / 32; ++c) {
          uint32_t o[32];
          ptx::tmem_ld_x32(tm_o + lane_addr + c * 32, o);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          ptx::tmem_st_x32(tm_o + lane_addr + c * 32, o);
        }
        ptx::tmem_st_wait();
        l *= alpha;
      } else if (grow && m_used > -INFINITY) {
        l *= ptx::ex2(m_used - m_tile);
      }
      if (grow) m_used = m_tile;
      const float m_ref = (m_used == -INFINITY) ? 0.f : m_used;
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        float p0 = ptx::ex2(__uint_as_float(r[e]) - m_ref);
        float p1 = ptx::ex2(__uint_as_float(r[e + 1]) - m_ref);
        if (e >= valid) p0 = 0.f;
        if (e + 1 >= valid) p1 = 0.f;
        l += p0 + p1;
        pk[e / 2] = kIsBf16 ? pack2_bf16(p0, p1) : pack2_f16(p0, p1);
      }
      ptx::tmem_st_x16(tm_s + lane_addr + b * 32, pk);
      ptx::tmem_st_wait();
      if (valid < kTile) {
        // rows past the end of the sequence may hold uninitialised cache data: zero them in the ckv
        // chunks (the V operand) so that 0 * garbage can never produce NaN.  Stage of tile j = j % 2.
        const int stg = j % S::kStages;
        ptx::mbar_wait(&k_full[stg], (j / S::kStages) & 1);
        uint8_t* tile = smem + S::kOffK + stg * S::kTileBytes;
        const int nz = (kTile - valid) * 4 * 8;  // rows x my 4 ckv chunks x 8 int4 per 128 B row
        for (int i = threadIdx.x - 128; i < nz; i += 128) {
          const int rrow = valid + i / 32, c = (i / 8) % 4, e = i % 8;
          reinterpret_cast<int4*>(tile + c * S::kChunkBytes + rrow * 128)[e] = make_int4(0, 0, 0, 0);
        }
        ptx::fence_proxy_async_smem();
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_ready[b]);
      ++od_cnt;
      if (j + 1 < ntiles) {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = rn[e];
      }
    }
    // ---- epilogue ----
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const float lse_v = l > 0.f ? m_used + ptx::lg2(l) : -INFINITY;
    if (ntiles > 0) {
      ptx::mbar_wait(o_done, (od_cnt - 1) & 1);
      ptx::tc_fence_after();
    }
    const bool row_ok = row < p.num_heads;
#pragma unroll 1
    for (int c = 0; c < kDvHalf / 32; ++c) {
      uint32_t o[32];
      if (ntiles > 0) {
        ptx::tmem_ld_x32(tm_o + lane_addr + c * 32, o);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = 0u;
      }
      if (!row_ok) continue;
      const int col0 = half * kDvHalf + c * 32;
      if (p.partial_o) {
        float* dst = p.partial_o + (int64_t(out_slot) * p.num_heads + row) * kCkv + col0;
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<float4*>(dst + e) = make_float4(__uint_as_float(o[e]) * inv, __uint_as_float(o[e + 1]) * inv,
                                                            __uint_as_float(o[e + 2]) * inv, __uint_as_float(o[e + 3]) * inv);
      } else {
        T* dst = reinterpret_cast<T*>(p.out) + int64_t(q_row) * p.o_stride_n + int64_t(row) * p.o_stride_h + col0;
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          Vec16<T> v;
#pragma unroll
          for (int u = 0; u < 8; ++u) v.v[u] = from_f32<T>(__uint_as_float(o[e + u]) * inv);
          st16(dst + e, v);
        }
      }
    }
    if (p.partial_o) {
      if (row_ok && half == 0) p.partial_lse[int64_t(out_slot) * p.num_heads + row] = lse_v;
    } else if (row_ok && half == 0 && p.lse) {
      p.lse[int64_t(q_row) * p.num_heads + row] = lse_v;
    }
  }

  ptx::grid_dep_launch();
  ptx::tc_fence_before();
  ptx::cluster_sync();  // neither CTA may retire while its peer can still write into its exchange buffers
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, 512);
  }
}

// Fold the split-KV partials of one query row: out[h] = sum_s w_s * partial_o[s][h], w_s = 2^(lse_s - max) / sum.
// grid (n_q, ceil(H / 8)), one warp per head, each lane 4 x float4 -> 512 B coalesced per load instruction.
template <typename T>
__global__ void __launch_bounds__(256)
mla_merge_kernel(const float* __restrict__ partial_o, const float* __restrict__ partial_lse,
                 const int32_t* __restrict__ row_parts, T* __restrict__ out, float* __restrict__ lse, int num_heads, int kmax,
                 int64_t o_stride_n, int64_t o_stride_h) {
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();
  const int q_row = blockIdx.x, h = blockIdx.y * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (h >= num_heads) return;
  const int nparts = row_parts[q_row];
  const int64_t slot0 = int64_t(q_row) * kmax;
  float mx = -INFINITY;
  for (int s = 0; s < nparts; ++s) mx = fmaxf(mx, partial_lse[(slot0 + s) * num_heads + h]);
  float den = 0.f;
  float4 acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < nparts; ++s) {
    const float ls = partial_lse[(slot0 + s) * num_heads + h];
    const float w = (mx == -INFINITY) ? 0.f : ptx::ex2(ls - mx);
    den += w;
    const float4* src = reinterpret_cast<const float4*>(partial_o + ((slot0 + s) * num_heads + h) * kCkv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = __ldcs(src + lane + 32 * i);
      acc[i].x += w * v.x;
      acc[i].y += w * v.y;
      acc[i].z += w * v.z;
      acc[i].w += w * v.w;
    }
  }
  const float inv = den > 0.f ? 1.f / den : 0.f;
  T* dst = out + int64_t(q_row) * o_stride_n + int64_t(h) * o_stride_h;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    struct alignas(8) Quad { T a, b, c, d; } q{from_f32<T>(acc[i].x * inv), from_f32<T>(acc[i].y * inv),
                                               from_f32<T>(acc[i].z * inv), from_f32<T>(acc[i].w * inv)};
    *reinterpret_cast<Quad*>(dst + (lane + 32 * i) * 4) = q;
  }
  if (lse && lane == 0) lse[int64_t(q_row) * num_heads + h] = den > 0.f ? mx + ptx::lg2(den) : -INFINITY;
}

}  // namespace

// q_nope [n, H, 512], q_pe [n, H, 64] (strides in elements), ckv_cache [pages, page, 512], kpe_cache [pages, page, 64]
extern "C" int mla_decode_run(void* q_nope, void* q_pe, void* ckv_cache, void* kpe_cache, void* kv_indices, void* work,
                              int64_t num_work, void* out, void* partial_o, void* partial_lse, void* row_parts,
                              int64_t kmax, void* lse, int64_t n_q, int64_t num_heads, int64_t page_size, int64_t num_pages_total, int64_t qn_sn, int64_t qn_sh,
                              int64_t qp_sn, int64_t qp_sh, int64_t ckv_sp, int64_t ckv_sn, int64_t kpe_sp, int64_t kpe_sn,
                              int64_t o_sn, int64_t o_sh, double sm_scale, int64_t dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(num_heads >= 1 && num_heads <= kHeads, "mla_sm100: num_heads must be <= 128");
  FIB_CHECK(dtype == kF16 || dtype == kBF16, "mla_sm100: dtype must be f16/bf16");
  FIB_CHECK(page_size >= 8 && (page_size % 8) == 0 && (page_size >= kTile ? page_size % kTile == 0 : kTile % page_size == 0),
            "mla_sm100: page_size must be 8, 16 or a multiple of 32");
  if (num_work == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const CUtensorMapDataType dt = dtype == kF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmQn, tmQp, tmC, tmP;
  {
    uint64_t dims[3] = {(uint64_t)kCkv, (uint64_t)num_heads, (uint64_t)n_q};
    uint64_t str[2] = {(uint64_t)qn_sh * 2, (uint64_t)qn_sn * 2};
    uint32_t box[3] = {64, (uint32_t)kHeads, 1};
    if (make_tmap(&tmQn, dt, 3, q_nope, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)kKpe, (uint64_t)num_heads, (uint64_t)n_q};
    uint64_t str[2] = {(uint64_t)qp_sh * 2, (uint64_t)qp_sn * 2};
    uint32_t box[3] = {64, (uint32_t)kHeads, 1};
    if (make_tmap(&tmQp, dt, 3, q_pe, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  const uint32_t box_rows = page_size < kTile ? (uint32_t)page_size : (uint32_t)kTile;
  {
    uint64_t dims[3] = {(uint64_t)kCkv, (uint64_t)page_size, (uint64_t)num_pages_total};
    uint64_t str[2] = {(uint64_t)ckv_sn * 2, (uint64_t)ckv_sp * 2};
    uint32_t box[3] = {64, box_rows, 1};
    if (make_tmap(&tmC, dt, 3, ckv_cache, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)kKpe, (uint64_t)page_size, (uint64_t)num_pages_total};
    uint64_t str[2] = {(uint64_t)kpe_sn * 2, (uint64_t)kpe_sp * 2};
    uint32_t box[3] = {64, box_rows, 1};
    if (make_tmap(&tmP, dt, 3, kpe_cache, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  MlaParams p;
  p.kv_indices = (const int32_t*)kv_indices;
  p.work = (const int32_t*)work;
  p.out = out;
  p.partial_o = (float*)partial_o;
  p.partial_lse = (float*)partial_lse;
  FIB_CHECK(!partial_o || (row_parts && partial_lse && kmax >= 1), "mla_sm100: split-KV needs the per-row part counts");
  p.lse = (float*)lse;
  p.num_heads = (int)num_heads;
  p.page_size = (int)page_size;
  p.o_stride_n = o_sn;
  p.o_stride_h = o_sh;
  p.sm_scale_log2 = (float)(sm_scale * 1.4426950408889634);
  const bool f16 = dtype == kF16;
  const uint32_t fmt = f16 ? ptx::kFmtF16 : ptx::kFmtBF16;
  const uint32_t idesc_qk = ptx::make_idesc_f16(fmt, kHeads, kTile, 0, 0);
  const uint32_t idesc_pv = ptx::make_idesc_f16(fmt, kHeads, kDvHalf, 0, 1);
  LaunchCfg lc(dim3((unsigned)num_work * 2), dim3(256), Smem::kTotal, stream, pdl != 0);  // clusters of 2 (kernel attribute)
  LaunchCfg lm(dim3((unsigned)n_q, (unsigned)((num_heads + 7) / 8)), dim3(256), 0, stream, pdl != 0);
  if (f16) {
    static bool set = false;
    if (!set) {
      FIB_CUDA_CHECK(cudaFuncSetAttribute(mla_decode_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::kTotal));
      set = true;
    }
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, mla_decode_kernel<__half>, tmQn, tmQp, tmC, tmP, p, idesc_qk, idesc_pv));
    if (partial_o)
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lm.cfg, mla_merge_kernel<__half>, (const float*)partial_o, (const float*)partial_lse,
                                        (const int32_t*)row_parts, (__half*)out, (float*)lse, (int)num_heads, (int)kmax, o_sn, o_sh));
  } else {
    static bool set = false;
    if (!set) {
      FIB_CUDA_CHECK(cudaFuncSetAttribute(mla_decode_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Smem::kTotal));
      set = true;
    }
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, mla_decode_kernel<__nv_bfloat16>, tmQn, tmQp, tmC, tmP, p, idesc_qk, idesc_pv));
    if (partial_o)
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lm.cfg, mla_merge_kernel<__nv_bfloat16>, (const float*)partial_o,
                                        (const float*)partial_lse, (const int32_t*)row_parts, (__nv_bfloat16*)out, (float*)lse,
                                        (int)num_heads, (int)kmax, o_sn, o_sh));
  }
  return 0;
}
