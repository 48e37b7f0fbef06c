// Paged-KV batch decode / small-q append attention for sm_100a (tcgen05 + TMEM + TMA).
//
// Capability parity target: reference BatchDecodeWithPagedKVCacheWrapper.run
// (flashinfer/decode.py:1229-1566; kernels include/flashinfer/attention/decode.cuh:396-656 and the
// closed trtllm-gen "SwapsMmaAbForGeneration" family, include/flashinfer/trtllm/fmha/fmhaKernels.cuh:799-838).
//
// B200-first design (NOT a port):
//  * swap-AB: S^T[kv,q] = K[kv,D] * Q^T  so the 128-token KV tile is the MMA-M side and the packed
//    (q_len x GQA-group) query rows are MMA-N (16..64): no tensor-core work is wasted on padding.
//    O^T[D,q] = V^T[D,kv] * P^T uses V straight from its TMA landing zone as an MN-major A operand.
//  * persistent grid of #SM CTAs; a host planner (runtime/planner.cpp) cuts the flattened
//    (request, kv_head, kv_tile) space into equal per-CTA quotas ("stream-K" over KV), so the
//    machine is balanced for any batch / length mix; partial (o,lse) go to a workspace and a
//    small merge kernel folds them.
//  * warp roles: warp0 = TMA producer (one lane per KV page -> paged gather is just a TMA
//    coordinate), warp1 = tcgen05.mma issuer, warp2 = TMEM allocator, warps4-7 = softmax +
//    output accumulation. K and V live in separate 3-stage 32 KB rings (192 KB in flight per SM).
//  * S and O_tile are double-buffered in TMEM so QK(i+1) overlaps softmax(i) and PV(i) overlaps
//    softmax(i+1); running O is kept in registers (fp32, one head-dim row per thread).
//  * LSE is returned in base-2 units like the reference (include/flashinfer/attention/state.cuh:46).
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>
#include <cuda_fp8.h>

using namespace fib200;

#ifndef FIB_POD_TU
FIB_EXPORT_LAST_ERROR()
#endif

namespace {

constexpr int kTileKV = 128;
constexpr int kSegInts = 12;

struct DecodeParams {
  const void* q;
  void* out;
  float* lse;  // optional [total_q, Hq]
  const int32_t* kv_indices;
  const int32_t* seg_info;        // [nseg][12]
  const int32_t* cta_seg_indptr;  // [grid+1]
  float* partial_o;               // [slots][rows_per_slot][D]
  float* partial_lse;             // [slots][rows_per_slot]
  int* merge_counters;            // [slots] zero-initialised, self-resetting (in-kernel merge); may be null
  int64_t q_stride_n, q_stride_h, o_stride_n, o_stride_h;
  int num_qo_heads, num_kv_heads, group, page_size, layout_hnd, rows_per_slot;
  int window_left, causal;
  int kv_early;          // PDL: the preceding kernel only APPENDS the q_len newest tokens of every request to the cache, so the TMA
                         // producers may fetch every older KV tile before griddepcontrol.wait (pipeline fill under the tail of the
                         // previous kernel); they wait right before the first tile that can hold an appended token
  float sm_scale_log2;   // sm_scale * log2(e)
  float soft_cap;        // 0 = off ; else logits = cap * tanh(x * sm_scale / cap)
  float sm_scale;
  const float* sinks;    // [num_qo_heads] attention-sink logits (natural log units) or null: exp(sink) joins the softmax denominator
};

// KVB = bytes per KV element: 2 (f16 / bf16 cache) or 1 (fp8 cache: Q and P are converted to e4m3 and both MMAs run
// as tcgen05 kind::f8f6f4 -- half the shared memory per tile, twice the MMA rate, no de-quantisation pass).
template <int NQ, int D, int KVB = 2>
struct DecodeSmem {
  static constexpr int kStagesK = 3, kStagesV = 3;
  static constexpr int kChunks = D * KVB / 128;              // 128-byte (SWIZZLE_128B) column chunks per row
  static constexpr int kCore = 16 / KVB;                      // elements per 16-byte core-matrix row
  static constexpr int kTileBytes = kTileKV * D * KVB;        // 32 KB (16-bit) / 16 KB (fp8) for D=128
  static constexpr int kChunkBytes = kTileKV * 128;           // one 64-col chunk of 128 rows
  static constexpr int kLBO = NQ * 16 + 16;                   // padded core-matrix stride (bank-conflict free)
  static constexpr int kQBytes = ((D / kCore) * kLBO + 1023) / 1024 * 1024;
  static constexpr int kPBytes = ((kTileKV / kCore) * kLBO + 1023) / 1024 * 1024;
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kOffK + kStagesK * kTileBytes;
  static constexpr int kOffQ = kOffV + kStagesV * kTileBytes;
  static constexpr int kOffP = kOffQ + 2 * kQBytes;          // Q^T is double-buffered: the next segment's Q is staged early
  static constexpr int kOffRed = kOffP + 2 * kPBytes;         // 2 x NQ x 4 floats (max) + NQ x 4 floats (sum)
  static constexpr int kRedBytes = 3 * NQ * 4 * 4;
  static constexpr int kOffBar = kOffRed + kRedBytes;
  static constexpr int kNumBars = 2 * kStagesK + 2 * kStagesV + 2 + 2 + 2 + 2;
  static constexpr int kTotal = kOffBar + kNumBars * 8 + 16 + 1024;
};

// Warp-level "halving butterfly": every lane holds NV per-column values for its own row; after
// the call lane-group `col` holds the warp-wide reduction of column `col` in v[0].
template <int NV, bool kMax>
__device__ __forceinline__ int butterfly_reduce(float (&v)[NV], int lane) {
  int col = 0;
  int held = NV;
#pragma unroll
  for (int mask = 16; mask >= 1; mask >>= 1) {
    if (held > 1) {
      const int half = held / 2;
      const bool upper = (lane & mask) != 0;
#pragma unroll
      for (int j = 0; j < NV / 2; ++j) {
        if (j < half) {
          const float lo = v[j], hi = v[j + half];
          const float send = upper ? lo : hi;
          const float keep = upper ? hi : lo;
          const float recv = __shfl_xor_sync(0xffffffffu, send, mask);
          v[j] = kMax ? fmaxf(keep, recv) : (keep + recv);
        }
      }
      held = half;
      col = col * 2 + (upper ? 1 : 0);
    } else {
      const float recv = __shfl_xor_sync(0xffffffffu, v[0], mask);
      v[0] = kMax ? fmaxf(v[0], recv) : (v[0] + recv);
    }
  }
  return col;
}

template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t* r) {
  if constexpr (N == 1) ptx::tmem_ld_x1(taddr, r);
  else if constexpr (N == 2) ptx::tmem_ld_x2(taddr, r);
  else if constexpr (N == 4) ptx::tmem_ld_x4(taddr, r);
  else if constexpr (N == 8) ptx::tmem_ld_x8(taddr, r);
  else if constexpr (N == 16) ptx::tmem_ld_x16(taddr, r);
  else ptx::tmem_ld_x32(taddr, r);
}

struct TileGeom {
  int token_start;  // first kv token covered by row 0 of the tile
  int rows;         // rows of the tile that map to cache slots (<=128)
  int first_page;   // index into the request's page list
  int page_off;     // row offset inside the page (page_size > 128 only)
};

__device__ __forceinline__ TileGeom tile_geom(int ti, int ps) {
  TileGeom g;
  if (ps <= kTileKV) {
    const int ppt = kTileKV / ps;
    g.first_page = ti * ppt;
    g.token_start = g.first_page * ps;
    g.rows = ppt * ps;
    g.page_off = 0;
  } else {
    const int tpp = (ps + kTileKV - 1) / kTileKV;
    g.first_page = ti / tpp;
    g.page_off = (ti % tpp) * kTileKV;
    g.token_start = g.first_page * ps + g.page_off;
    g.rows = min(kTileKV, ps - g.page_off);
  }
  return g;
}

// NQ: MMA N (padded q rows, multiple of 16); NV: power-of-two number of columns actually processed.
// Body as a device function of the CTA index (shared with the fused POD kernel, pod_sm100.cu, where it runs with a
// 384-thread block: warps 9..11 have no role and only take part in the CTA-wide barriers).
template <int NQ, int NV, int D, typename T, int KVB>
__device__ __forceinline__ void decode_body(const CUtensorMap& tmK, const CUtensorMap& tmV, const DecodeParams& p,
                                            uint32_t idesc_qk, uint32_t idesc_pv, int cta) {
  using S = DecodeSmem<NQ, D, KVB>;
  static_assert(D == 128, "head_dim 128 specialisation");
  constexpr bool kKV8 = KVB == 1;
  constexpr float kPScale = kKV8 ? 256.f : 1.f;  // fp8 P is stored as p * 256 (<= 256 < 448): keeps small probabilities
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* k_full = bars;
  uint64_t* k_empty = k_full + S::kStagesK;
  uint64_t* v_full = k_empty + S::kStagesK;
  uint64_t* v_empty = v_full + S::kStagesV;
  uint64_t* s_full = v_empty + S::kStagesV;  // [2]
  uint64_t* o_full = s_full + 2;             // [2]
  uint64_t* p_ready = o_full + 2;            // [2]
  uint64_t* q_full = p_ready + 2;            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(q_full + 2);
  float* red_max = reinterpret_cast<float*>(smem + S::kOffRed);  // [2][NV][4]
  float* red_sum = red_max + 2 * NQ * 4;                         // [NV][4]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int seg_begin = p.cta_seg_indptr[cta];
  const int seg_end = p.cta_seg_indptr[cta + 1];
  if (seg_begin >= seg_end) return;

  constexpr uint32_t kTmemCols = (4 * NQ < 32) ? 32 : 4 * NQ;
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    for (int i = 0; i < S::kStagesK; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < S::kStagesV; ++i) {
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&o_full[i], 1);
      ptx::mbar_init(&p_ready[i], 128);
    }
    ptx::mbar_init(&q_full[0], 128);
    ptx::mbar_init(&q_full[1], 128);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish<1>();
  }
  // zero Q / P staging buffers once (padding columns stay zero forever)
  for (int i = threadIdx.x; i < (2 * S::kQBytes + 2 * S::kPBytes) / 16; i += blockDim.x)
    reinterpret_cast<int4*>(smem + S::kOffQ)[i] = make_int4(0, 0, 0, 0);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const bool is_producer = warp == 0 || warp == 2 || warp == 3 || warp == 8;
  bool dep_waited = !(p.kv_early && is_producer);
  if (dep_waited) ptx::grid_dep_wait();
  ptx::grid_dep_launch();  // early trigger: dependents overlap their prologue, they still wait for our completion

  const int ps = p.page_size;

  if (is_producer) {
    // ============================ TMA producers (4 warps) ============================
    // ncu showed one producer warp saturating on UTMALDG issue (the per-lane page boxes are serialised
    // through an ELECT/R2UR loop, ~100 cycles each), so the 32 boxes of a tile are split over four
    // warps: group g = {K,V} x {64-column chunk}.  The chunk-0 warp of each tensor arms expect_tx
    // with the byte count of the whole tile; complete_tx from the sibling warp may land first, which
    // is legal (the phase cannot complete before the arming arrive).
    const int g = (warp == 0) ? 0 : (warp == 2 ? 1 : (warp == 3 ? 2 : 3));
    const int is_v = g >> 1;
    const int chunk = g & 1;
    const int nstages = is_v ? S::kStagesV : S::kStagesK;
    uint64_t* full_bars = is_v ? v_full : k_full;
    uint64_t* empty_bars = is_v ? v_empty : k_empty;
    const CUtensorMap* tm = is_v ? &tmV : &tmK;
    // 16-bit cache: warp `chunk` loads column chunk `chunk` of every page box; fp8 cache (one chunk per row): the two
    // warps of a tensor split the page boxes by parity instead
    uint8_t* ring = smem + (is_v ? S::kOffV : S::kOffK) + (kKV8 ? 0 : chunk * S::kChunkBytes);
    int st = 0;
    uint32_t ph = 0;
    for (int seg = seg_begin; seg < seg_end; ++seg) {
      const int32_t* si = p.seg_info + seg * kSegInts;
      const int kv_head = si[1], t0 = si[2], t1 = si[3], page_start = si[8], num_pages = si[9];
      const int first_new = si[7] - si[6];  // kv_len - q_len: position of the oldest token the preceding kernel may have written
      for (int ti = t0; ti < t1; ++ti) {
        const TileGeom g2 = tile_geom(ti, ps);
        if (!dep_waited && g2.token_start + g2.rows > first_new) {
          ptx::grid_dep_wait();
          dep_waited = true;
        }
        int n_boxes, box_rows;
        if (ps <= kTileKV) {
          n_boxes = min(kTileKV / ps, num_pages - g2.first_page);
          box_rows = ps;
        } else {
          n_boxes = 1;
          box_rows = 0;
        }
        const uint32_t box_bytes = (ps <= kTileKV ? ps : kTileKV) * 128;
        const uint32_t tx = uint32_t(n_boxes) * box_bytes * S::kChunks;
        int my_pages[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int bi = lane + j * 32;
          my_pages[j] = (bi < n_boxes && (!kKV8 || (bi & 1) == chunk)) ? __ldg(p.kv_indices + page_start + g2.first_page + bi) : -1;
        }
        if (lane == 0) {
          ptx::mbar_wait(&empty_bars[st], ph ^ 1);
          if (chunk == 0) ptx::mbar_arrive_expect_tx(&full_bars[st], tx);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (my_pages[j] >= 0) {
            uint8_t* dst = ring + st * S::kTileBytes + (lane + j * 32) * box_rows * 128;
            if (p.layout_hnd)
              ptx::tma_load_4d(dst, tm, &full_bars[st], kKV8 ? 0 : chunk * 64, g2.page_off, kv_head, my_pages[j], ptx::kEvictFirst);
            else
              ptx::tma_load_4d(dst, tm, &full_bars[st], kKV8 ? 0 : chunk * 64, kv_head, g2.page_off, my_pages[j], ptx::kEvictFirst);
          }
        }
        if (++st == nstages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    uint32_t gt = 0;      // global tile counter (buffer parity)
    uint32_t seg_par = 0;
    const uint32_t q_base = ptx::smem_u32(smem + S::kOffQ);
    const uint32_t p_addr = ptx::smem_u32(smem + S::kOffP);
    auto issue_pv = [&](uint32_t tile_id) {
      const uint32_t b = tile_id & 1;
      ptx::mbar_wait(&p_ready[b], (tile_id >> 1) & 1);
      ptx::mbar_wait(&v_full[vs], vph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t v_addr = ptx::smem_u32(smem + S::kOffV + vs * S::kTileBytes);
        // A = V^T : MN-major SW128, LBO = stride between 64-wide d chunks, SBO = 8 kv rows
        const uint64_t da = ptx::make_smem_desc(v_addr, S::kChunkBytes, 1024, ptx::kSwz128);
        // B = P^T : K-major, no swizzle, core matrices [8 q][8 kv]
        const uint64_t db = ptx::make_smem_desc(p_addr + b * S::kPBytes, S::kLBO, 128, ptx::kSwzNone);
        const uint32_t d_tmem = tmem_base + 2 * NQ + b * NQ;
        if constexpr (kKV8) {
#pragma unroll
          for (int k = 0; k < kTileKV / 32; ++k)
            ptx::mma_f8f6f4_ss<1>(d_tmem, ptx::desc_advance(da, k * 32 * 128), ptx::desc_advance(db, k * 2 * S::kLBO),
                                  idesc_pv, k > 0 ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < kTileKV / 16; ++k)
            ptx::mma_f16_ss<1>(d_tmem, ptx::desc_advance(da, k * 16 * 128), ptx::desc_advance(db, k * 2 * S::kLBO),
                               idesc_pv, k > 0 ? 1u : 0u);
        }
        ptx::mma_commit(&v_empty[vs]);
        ptx::mma_commit(&o_full[b]);
      }
      __syncwarp();
      if (++vs == S::kStagesV) {
        vs = 0;
        vph ^= 1;
      }
    };
    for (int seg = seg_begin; seg < seg_end; ++seg) {
      const int32_t* si = p.seg_info + seg * kSegInts;
      const int t0 = si[2], t1 = si[3];
      const int sl = seg - seg_begin;  // local segment index selects the Q buffer
      ptx::mbar_wait(&q_full[sl & 1], (sl >> 1) & 1);
      const uint32_t q_addr = q_base + (sl & 1) * S::kQBytes;
      (void)seg_par;
      for (int ti = t0; ti < t1; ++ti, ++gt) {
        const uint32_t b = gt & 1;
        ptx::mbar_wait(&k_full[ks], kph);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t k_addr = ptx::smem_u32(smem + S::kOffK + ks * S::kTileBytes);
          const uint64_t db = ptx::make_smem_desc(q_addr, S::kLBO, 128, ptx::kSwzNone);
          const uint32_t d_tmem = tmem_base + b * NQ;
          if constexpr (kKV8) {
#pragma unroll
            for (int k = 0; k < D / 32; ++k) {  // K = 32 fp8 elements = 32 bytes per MMA
              const uint64_t da = ptx::make_smem_desc(k_addr + k * 32, 16, 1024, ptx::kSwz128);
              ptx::mma_f8f6f4_ss<1>(d_tmem, da, ptx::desc_advance(db, k * 2 * S::kLBO), idesc_qk, k > 0 ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
              const uint64_t da =
                  ptx::make_smem_desc(k_addr + (k / 4) * S::kChunkBytes + (k % 4) * 32, 16, 1024, ptx::kSwz128);
              ptx::mma_f16_ss<1>(d_tmem, da, ptx::desc_advance(db, k * 2 * S::kLBO), idesc_qk, k > 0 ? 1u : 0u);
            }
          }
          ptx::mma_commit(&k_empty[ks]);
          ptx::mma_commit(&s_full[b]);
        }
        __syncwarp();
        if (++ks == S::kStagesK) {
          ks = 0;
          kph ^= 1;
        }
        if (ti > t0) issue_pv(gt - 1);
      }
      issue_pv(gt - 1);
    }
  } else if (warp >= 4 && warp < 8) {
    // ============================ softmax / accumulate ============================
    const int q4 = warp - 4;              // TMEM lane quadrant
    const int row = q4 * 32 + lane;       // kv row inside the tile for S^T; head-dim index for O^T
    const uint32_t lane_addr = uint32_t(q4 * 32) << 16;
    uint32_t gt = 0;
    int vs = 0;
    uint32_t vph = 0;
    uint32_t seg_par = 0;
    (void)seg_par;
    const T* qbase = reinterpret_cast<const T*>(p.q);
    T* obase = reinterpret_cast<T*>(p.out);
    const int G = p.group;

    for (int seg = seg_begin; seg < seg_end; ++seg) {
      const int32_t* si = p.seg_info + seg * kSegInts;
      const int kv_head = si[1], t0 = si[2], t1 = si[3], slot = si[4], q_start = si[5], q_len = si[6],
                kv_len = si[7];
      const int nq = q_len * G;  // valid columns

      // ---- stage Q^T (B operand of QK) : element (c, d) at (d/8)*LBO + (c/8)*128 + (c%8)*16 + (d%8)*2
      // Double-buffered: segment i stages the Q of segment i+1 right away, so the MMA warp never waits for a Q
      // round trip at a segment boundary (buffer (i+1)&1 was last read by segment i-1, whose S tiles have all
      // been consumed by these warps already).
      auto stage_q = [&](int sg) {
        const int32_t* sj = p.seg_info + sg * kSegInts;
        const int kvh = sj[1], qs = sj[5], nqv = sj[6] * G;
        uint8_t* qbuf = smem + S::kOffQ + ((sg - seg_begin) & 1) * S::kQBytes;
        const int tid = threadIdx.x - 128;
        if constexpr (kKV8) {
          // 16 query elements -> 16 e4m3 bytes = one core-matrix row
          for (int v = tid; v < NV * (D / 16); v += 128) {
            const int c = v / (D / 16), dg = v % (D / 16);
            int4 val = make_int4(0, 0, 0, 0);
            if (c < nqv) {
              const int qi = c / G, g = c % G;
              const T* src = qbase + int64_t(qs + qi) * p.q_stride_n + int64_t(kvh * G + g) * p.q_stride_h + dg * 16;
              const int4 lo = __ldg(reinterpret_cast<const int4*>(src)), hi = __ldg(reinterpret_cast<const int4*>(src + 8));
              const T* l8 = reinterpret_cast<const T*>(&lo);
              const T* h8 = reinterpret_cast<const T*>(&hi);
              uint8_t* ob = reinterpret_cast<uint8_t*>(&val);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const __nv_fp8_e4m3 a(fminf(fmaxf(to_f32(l8[e]), -448.f), 448.f)), b(fminf(fmaxf(to_f32(h8[e]), -448.f), 448.f));
                ob[e] = *reinterpret_cast<const uint8_t*>(&a);
                ob[8 + e] = *reinterpret_cast<const uint8_t*>(&b);
              }
            }
            *reinterpret_cast<int4*>(qbuf + dg * S::kLBO + (c >> 3) * 128 + (c & 7) * 16) = val;
          }
        } else {
        for (int v = tid; v < NV * (D / 8); v += 128) {
          const int c = v / (D / 8), dg = v % (D / 8);
          int4 val = make_int4(0, 0, 0, 0);
          if (c < nqv) {
            const int qi = c / G, g = c % G;
            val = __ldg(reinterpret_cast<const int4*>(qbase + int64_t(qs + qi) * p.q_stride_n +
                                                      int64_t(kvh * G + g) * p.q_stride_h + dg * 8));
          }
          *reinterpret_cast<int4*>(qbuf + dg * S::kLBO + (c >> 3) * 128 + (c & 7) * 16) = val;
        }
        }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&q_full[(sg - seg_begin) & 1]);
      };
      if (seg == seg_begin) stage_q(seg);
      if (seg + 1 < seg_end) stage_q(seg + 1);

      float m[NV], l[NV], o[NV], alpha_saved[NV];
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        m[c] = -INFINITY;
        l[c] = 0.f;
        o[c] = 0.f;
        alpha_saved[c] = 1.f;
      }

      auto consume_o = [&](uint32_t tile_id) {
        const uint32_t b = tile_id & 1;
        ptx::mbar_wait(&o_full[b], (tile_id >> 1) & 1);
        ptx::tc_fence_after();
        uint32_t r[NV];
        tmem_ld_n<NV>(tmem_base + lane_addr + 2 * NQ + b * NQ, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < NV; ++c) o[c] = o[c] * alpha_saved[c] + __uint_as_float(r[c]) * (1.f / kPScale);
      };

      for (int ti = t0; ti < t1; ++ti, ++gt) {
        const uint32_t b = gt & 1;
        const TileGeom g = tile_geom(ti, ps);
        const int valid = min(g.rows, kv_len - g.token_start);
        const int kv_pos = g.token_start + row;

        // ---- step A: softmax of tile gt ----
        ptx::mbar_wait(&s_full[b], (gt >> 1) & 1);
        ptx::tc_fence_after();
        uint32_t sr[NV];
        tmem_ld_n<NV>(tmem_base + lane_addr + b * NQ, sr);
        ptx::tmem_ld_wait();
        float s[NV], red[NV];
#pragma unroll
        for (int c = 0; c < NV; ++c) {
          float x = __uint_as_float(sr[c]);
          if (p.soft_cap > 0.f) {
            x = p.soft_cap * ptx::tanh_approx(x * p.sm_scale / p.soft_cap) * 1.4426950408889634f;
          } else {
            x *= p.sm_scale_log2;
          }
          bool ok = (row < valid) && (c < nq);
          if (p.causal | (p.window_left >= 0)) {
            const int q_pos = kv_len - q_len + c / G;
            if (p.causal) ok = ok && (kv_pos <= q_pos);
            if (p.window_left >= 0) ok = ok && (kv_pos >= q_pos - p.window_left);
          }
          s[c] = ok ? x : -INFINITY;
          red[c] = s[c];
        }
        const int col = butterfly_reduce<NV, true>(red, lane);
        float* rm = red_max + b * NQ * 4;
        rm[col * 4 + q4] = red[0];
        ptx::named_bar_sync(1, 128);
        float alpha[NV];
#pragma unroll
        for (int c = 0; c < NV; ++c) {
          const float4 w = *reinterpret_cast<const float4*>(rm + c * 4);
          const float tmax = fmaxf(fmaxf(w.x, w.y), fmaxf(w.z, w.w));
          const float m_new = fmaxf(m[c], tmax);
          const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
          alpha[c] = ptx::ex2(m[c] - m_use);  // m == -inf -> 0 (o,l are 0 anyway)
          const float pv = ptx::ex2(s[c] - m_use);
          l[c] = l[c] * alpha[c] + pv;
          m[c] = m_new;
          // P^T element (c, row)
          if constexpr (kKV8) {
            const __nv_fp8_e4m3 p8(pv * kPScale);
            smem[S::kOffP + b * S::kPBytes + (row >> 4) * S::kLBO + (c >> 3) * 128 + (c & 7) * 16 + (row & 15)] =
                *reinterpret_cast<const uint8_t*>(&p8);
          } else {
            *reinterpret_cast<T*>(smem + S::kOffP + b * S::kPBytes + (row >> 3) * S::kLBO + (c >> 3) * 128 + (c & 7) * 16 +
                                  (row & 7) * 2) = from_f32<T>(pv);
          }
        }
        if (valid < kTileKV) {
          // rows past the end of the sequence: V may hold stale / uninitialised data -> zero it so
          // that 0 * garbage can never produce NaN.
          ptx::mbar_wait(&v_full[vs], vph);
          if (row >= valid) {
            uint8_t* vrow = smem + S::kOffV + vs * S::kTileBytes + row * 128;
#pragma unroll
            for (int c = 0; c < S::kChunks; ++c)
#pragma unroll
              for (int j = 0; j < 8; ++j) reinterpret_cast<int4*>(vrow + c * S::kChunkBytes)[j] = make_int4(0, 0, 0, 0);
          }
        }
        if (++vs == S::kStagesV) {
          vs = 0;
          vph ^= 1;
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&p_ready[b]);

        // ---- step B: fold O_tile of the previous tile ----
        if (ti > t0) consume_o(gt - 1);
#pragma unroll
        for (int c = 0; c < NV; ++c) alpha_saved[c] = alpha[c];
      }
      consume_o(gt - 1);
      ptx::tc_fence_before();

      // ---- segment epilogue: reduce l over the 128 kv-row threads, normalise, write ----
      float ls[NV];
#pragma unroll
      for (int c = 0; c < NV; ++c) ls[c] = l[c];
      const int col = butterfly_reduce<NV, false>(ls, lane);
      red_sum[col * 4 + q4] = ls[0];
      ptx::named_bar_sync(1, 128);
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        if (c < nq) {
          const float4 w = *reinterpret_cast<const float4*>(red_sum + c * 4);
          float lt = w.x + w.y + w.z + w.w;
          float mc = m[c];
          const int qi = c / G, g = c % G;
          if (slot < 0 && p.sinks) {  // final (unsplit) state: fold the attention sink into the denominator
            const float s2 = p.sinks[kv_head * G + g] * 1.4426950408889634f;
            if (!(lt > 0.f)) {
              mc = s2;
              lt = 1.f;
            } else {
              lt += ptx::ex2(s2 - mc);
            }
          }
          const float inv = lt > 0.f ? 1.f / lt : 0.f;
          const float val = o[c] * inv;
          const float lse = lt > 0.f ? mc + ptx::lg2(lt) : -INFINITY;
          if (slot < 0) {
            obase[int64_t(q_start + qi) * p.o_stride_n + int64_t(kv_head * G + g) * p.o_stride_h + row] = from_f32<T>(val);
            if (p.lse && row == 0) p.lse[int64_t(q_start + qi) * p.num_qo_heads + kv_head * G + g] = lse;
          } else {
            p.partial_o[(int64_t(slot) * p.rows_per_slot + c) * D + row] = val;
            if (row == 0) p.partial_lse[int64_t(slot) * p.rows_per_slot + c] = lse;
          }
        }
      }
      if (slot >= 0 && p.merge_counters) {
        // ---- in-kernel merge: the last part of this (request, kv head) to arrive folds all partials ----
        const int first_slot = si[10], nparts = si[11];
        __threadfence();
        ptx::named_bar_sync(1, 128);
        if (threadIdx.x == 128) {
          const int old = atomicAdd(&p.merge_counters[first_slot], 1);
          const int last = (old == nparts - 1);
          if (last) p.merge_counters[first_slot] = 0;
          reinterpret_cast<volatile int*>(red_sum)[0] = last;
        }
        ptx::named_bar_sync(1, 128);
        const bool last = reinterpret_cast<volatile int*>(red_sum)[0] != 0;
        if (last) {
          __threadfence();
          for (int c = 0; c < nq; ++c) {
            float mx = -INFINITY;
            for (int sidx = 0; sidx < nparts; ++sidx)
              mx = fmaxf(mx, __ldcg(p.partial_lse + int64_t(first_slot + sidx) * p.rows_per_slot + c));
            float acc = 0.f, den = 0.f;
            for (int sidx = 0; sidx < nparts; ++sidx) {
              const float ls = __ldcg(p.partial_lse + int64_t(first_slot + sidx) * p.rows_per_slot + c);
              const float w = (mx == -INFINITY) ? 0.f : ptx::ex2(ls - mx);
              acc += w * __ldcg(p.partial_o + (int64_t(first_slot + sidx) * p.rows_per_slot + c) * D + row);
              den += w;
            }
            const int qi = c / G, g = c % G;
            if (p.sinks) {
              const float s2 = p.sinks[kv_head * G + g] * 1.4426950408889634f;
              if (mx == -INFINITY) {
                mx = s2;
                den = 1.f;
              } else {
                den += ptx::ex2(s2 - mx);
              }
            }
            const float val = den > 0.f ? acc / den : 0.f;
            obase[int64_t(q_start + qi) * p.o_stride_n + int64_t(kv_head * G + g) * p.o_stride_h + row] = from_f32<T>(val);
            if (p.lse && row == 0)
              p.lse[int64_t(q_start + qi) * p.num_qo_heads + kv_head * G + g] = den > 0.f ? mx + ptx::lg2(den) : -INFINITY;
          }
        }
      }
      ptx::named_bar_sync(1, 128);  // red_sum / Q smem reuse across segments
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
  }
}

template <int NQ, int NV, int D, typename T, int KVB>
__global__ void __launch_bounds__(288, 1)
decode_paged_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const DecodeParams p, uint32_t idesc_qk, uint32_t idesc_pv) {
  decode_body<NQ, NV, D, T, KVB>(tmK, tmV, p, idesc_qk, idesc_pv, blockIdx.x);
}

// Merge split-KV partials: one CTA per merge item, thread = head-dim index.
// item = {slot0, nparts, q_start, q_len, kv_head, pad...} (8 ints)
template <int D, typename T>
__global__ void __launch_bounds__(D)
decode_merge_kernel(const int32_t* __restrict__ items, int num_items, const float* __restrict__ partial_o,
                    const float* __restrict__ partial_lse, T* __restrict__ out, float* __restrict__ lse_out,
                    int rows_per_slot, int group, int num_qo_heads, int64_t o_stride_n, int64_t o_stride_h,
                    const float* __restrict__ sinks) {
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();  // early trigger: dependents overlap their prologue, they still wait for our completion
  for (int it = blockIdx.x; it < num_items; it += gridDim.x) {
    const int32_t* mi = items + it * 8;
    const int slot0 = mi[0], nparts = mi[1], q_start = mi[2], q_len = mi[3], kv_head = mi[4];
    const int nq = q_len * group;
    const int d = threadIdx.x;
    for (int c = 0; c < nq; ++c) {
      float mx = -INFINITY;
      for (int s = 0; s < nparts; ++s) mx = fmaxf(mx, partial_lse[int64_t(slot0 + s) * rows_per_slot + c]);
      float acc = 0.f, den = 0.f;
      for (int s = 0; s < nparts; ++s) {
        const float ls = partial_lse[int64_t(slot0 + s) * rows_per_slot + c];
        const float w = (mx == -INFINITY) ? 0.f : ptx::ex2(ls - mx);
        acc += w * partial_o[(int64_t(slot0 + s) * rows_per_slot + c) * D + d];
        den += w;
      }
      const int qi = c / group, g = c % group;
      if (sinks) {
        const float s2 = sinks[kv_head * group + g] * 1.4426950408889634f;
        if (mx == -INFINITY) {
          mx = s2;
          den = 1.f;
        } else {
          den += ptx::ex2(s2 - mx);
        }
      }
      const float val = den > 0.f ? acc / den : 0.f;
      out[int64_t(q_start + qi) * o_stride_n + int64_t(kv_head * group + g) * o_stride_h + d] = from_f32<T>(val);
      if (lse_out && d == 0)
        lse_out[int64_t(q_start + qi) * num_qo_heads + kv_head * group + g] = den > 0.f ? mx + ptx::lg2(den) : -INFINITY;
    }
  }
}

template <int NQ, int NV, typename T, int KVB>
int launch_decode(const CUtensorMap& tmK, const CUtensorMap& tmV, const DecodeParams& p, int grid, bool f16, bool pdl,
                  cudaStream_t stream, int kv_fmt) {
  using S = DecodeSmem<NQ, 128, KVB>;
  auto kern = decode_paged_kernel<NQ, NV, 128, T, KVB>;
  static bool attr_set = false;
  if (!attr_set) {
    FIB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set = true;
  }
  const uint32_t fmt = f16 ? ptx::kFmtF16 : ptx::kFmtBF16;
  // fp8 cache: A = K / V in the cache format (e4m3 = 0, e5m2 = 1), B = Q / P converted to e4m3
  const uint32_t idesc_qk = KVB == 1 ? ptx::make_idesc_f8((uint32_t)kv_fmt, ptx::kFmtE4M3, 128, NQ, 0, 0) : ptx::make_idesc_f16(fmt, 128, NQ, 0, 0);
  const uint32_t idesc_pv = KVB == 1 ? ptx::make_idesc_f8((uint32_t)kv_fmt, ptx::kFmtE4M3, 128, NQ, 1, 0) : ptx::make_idesc_f16(fmt, 128, NQ, 1, 0);
#ifdef FIB_POD_TU
  if (pod_stage().armed) {
    // POD: a prefill launch is parked -> run both bodies in ONE grid (prefill CTAs first, then the decode CTAs)
    if (!pod_stage().have_prefill) return set_error("pod: decode launched before the prefill side was staged");
    if constexpr (KVB == 2 && NQ == 16 && (NV == 1 || NV == 4 || NV == 8))
      return pod_launch<NV, T>(tmK, tmV, &p, sizeof(p), idesc_qk, idesc_pv, grid, (int)S::kTotal, pdl, stream);
    else
      return set_error("pod: this decode variant (fp8 KV or q rows not in {1, 4, 8}) has no fused POD kernel");
  }
#endif
  LaunchCfg lc(dim3(grid), dim3(288), S::kTotal, stream, pdl);
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, tmK, tmV, p, idesc_qk, idesc_pv));
  return 0;
}

}  // namespace

// Returns the dynamic smem / launch geometry facts the planner needs.
extern "C" int decode_paged_info(int64_t* tile_kv, int64_t* seg_ints, int64_t* max_rows) {
  *tile_kv = kTileKV;
  *seg_ints = kSegInts;
  *max_rows = 32;
  return 0;
}

// q: [total_q, Hq, D]; k_cache/v_cache: paged, strides given in elements
// (stride_page, stride_n (token in page), stride_h); kv dtype == q dtype (f16/bf16).
extern "C" int decode_paged_run(void* q, void* k_cache, void* v_cache, void* out, void* lse, void* kv_indices,
                                void* seg_info, void* cta_seg_indptr, void* merge_items, int64_t num_merge_items,
                                void* partial_o, void* partial_lse, void* merge_counters, int64_t grid, int64_t max_q_rows,
                                int64_t num_qo_heads, int64_t num_kv_heads, int64_t head_dim, int64_t page_size,
                                int64_t num_pages_total, int64_t kv_stride_page, int64_t kv_stride_n,
                                int64_t kv_stride_h, int64_t layout_hnd, int64_t q_stride_n, int64_t q_stride_h,
                                int64_t o_stride_n, int64_t o_stride_h, double sm_scale, double soft_cap,
                                int64_t window_left, int64_t causal, void* sinks, int64_t dtype, int64_t kv_dtype, int64_t pdl,
                                int64_t stream_) {
  FIB_CHECK(head_dim == 128, "decode_sm100: only head_dim 128 is specialised");
  FIB_CHECK(dtype == kF16 || dtype == kBF16, "decode_sm100: q dtype must be f16/bf16");
  const bool kv8 = kv_dtype == kE4M3 || kv_dtype == kE5M2;
  FIB_CHECK(kv8 || kv_dtype == dtype, "decode_sm100: kv dtype must equal the q dtype or be fp8 (e4m3 / e5m2)");
  FIB_CHECK(max_q_rows >= 1 && max_q_rows <= 32, "decode_sm100: q_len*group must be <= 32");
  FIB_CHECK(kv_stride_n % (kv8 ? 16 : 8) == 0 && kv_stride_h % (kv8 ? 16 : 8) == 0 && kv_stride_page % (kv8 ? 16 : 8) == 0,
            "kv strides must be 16B multiples");
  FIB_CHECK(page_size <= 128 || true, "");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const CUtensorMapDataType dt = kv8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                      : (dtype == kF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  const uint64_t esz = kv8 ? 1 : 2;
  const uint32_t box_cols = kv8 ? 128 : 64;  // 128 bytes either way
  CUtensorMap tmK, tmV;
  const uint32_t box_rows = page_size <= kTileKV ? (uint32_t)page_size : (uint32_t)kTileKV;
  for (int i = 0; i < 2; ++i) {
    const void* base = i == 0 ? k_cache : v_cache;
    CUtensorMap* tm = i == 0 ? &tmK : &tmV;
    if (layout_hnd) {
      uint64_t dims[4] = {(uint64_t)head_dim, (uint64_t)page_size, (uint64_t)num_kv_heads, (uint64_t)num_pages_total};
      uint64_t str[3] = {(uint64_t)kv_stride_n * esz, (uint64_t)kv_stride_h * esz, (uint64_t)kv_stride_page * esz};
      uint32_t box[4] = {box_cols, box_rows, 1, 1};
      if (make_tmap(tm, dt, 4, base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    } else {
      uint64_t dims[4] = {(uint64_t)head_dim, (uint64_t)num_kv_heads, (uint64_t)page_size, (uint64_t)num_pages_total};
      uint64_t str[3] = {(uint64_t)kv_stride_h * esz, (uint64_t)kv_stride_n * esz, (uint64_t)kv_stride_page * esz};
      uint32_t box[4] = {box_cols, 1, box_rows, 1};
      if (make_tmap(tm, dt, 4, base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    }
  }
  DecodeParams p;
  p.q = q;
  p.out = out;
  p.lse = (float*)lse;
  p.kv_indices = (const int32_t*)kv_indices;
  p.seg_info = (const int32_t*)seg_info;
  p.cta_seg_indptr = (const int32_t*)cta_seg_indptr;
  p.partial_o = (float*)partial_o;
  p.partial_lse = (float*)partial_lse;
  p.merge_counters = (int*)merge_counters;
  p.q_stride_n = q_stride_n;
  p.q_stride_h = q_stride_h;
  p.o_stride_n = o_stride_n;
  p.o_stride_h = o_stride_h;
  p.num_qo_heads = (int)num_qo_heads;
  p.num_kv_heads = (int)num_kv_heads;
  p.group = (int)(num_qo_heads / num_kv_heads);
  p.page_size = (int)page_size;
  p.layout_hnd = (int)layout_hnd;
  p.window_left = (int)window_left;
  p.causal = (int)causal;
  p.kv_early = (pdl == 2) ? 1 : 0;
  p.sinks = reinterpret_cast<const float*>(sinks);
  p.sm_scale = (float)sm_scale;
  p.sm_scale_log2 = (float)(sm_scale * 1.4426950408889634);
  p.soft_cap = (float)soft_cap;
  const bool f16 = dtype == kF16;
  int NV = 1;
  while (NV < max_q_rows) NV *= 2;
  p.rows_per_slot = NV;
#define FIB_DEC(NQ_, NV_)                                                                                     \
  {                                                                                                           \
    const int kvf = kv_dtype == kE5M2 ? 1 : 0;                                                                \
    int rc;                                                                                                   \
    if (kv8)                                                                                                  \
      rc = f16 ? launch_decode<NQ_, NV_, __half, 1>(tmK, tmV, p, (int)grid, true, pdl != 0, stream, kvf)      \
               : launch_decode<NQ_, NV_, __nv_bfloat16, 1>(tmK, tmV, p, (int)grid, false, pdl != 0, stream, kvf); \
    else                                                                                                      \
      rc = f16 ? launch_decode<NQ_, NV_, __half, 2>(tmK, tmV, p, (int)grid, true, pdl != 0, stream, 0)        \
               : launch_decode<NQ_, NV_, __nv_bfloat16, 2>(tmK, tmV, p, (int)grid, false, pdl != 0, stream, 0); \
    if (rc) return rc;                                                                                        \
  }
  switch (NV) {
    case 1: FIB_DEC(16, 1); break;
    case 2: FIB_DEC(16, 2); break;
    case 4: FIB_DEC(16, 4); break;
    case 8: FIB_DEC(16, 8); break;
    case 16: FIB_DEC(16, 16); break;
    default: FIB_DEC(32, 32); break;
  }
#undef FIB_DEC
  if (num_merge_items > 0 && merge_counters == nullptr) {
    int blocks = (int)num_merge_items;
    LaunchCfg lc(dim3(blocks), dim3(128), 0, stream, pdl != 0);
    if (f16) {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, decode_merge_kernel<128, __half>, (const int32_t*)merge_items,
                                        (int)num_merge_items, (const float*)partial_o, (const float*)partial_lse,
                                        (__half*)out, (float*)lse, p.rows_per_slot, p.group, p.num_qo_heads, o_stride_n,
                                        o_stride_h, p.sinks));
    } else {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, decode_merge_kernel<128, __nv_bfloat16>, (const int32_t*)merge_items,
                                        (int)num_merge_items, (const float*)partial_o, (const float*)partial_lse,
                                        (__nv_bfloat16*)out, (float*)lse, p.rows_per_slot, p.group, p.num_qo_heads,
                                        o_stride_n, o_stride_h, p.sinks));
    }
  }
  return 0;
}
