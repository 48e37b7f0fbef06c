// Prefill / append attention (ragged or paged KV) for sm_100a: persistent warp-specialised FMHA forward
// on tcgen05 + TMEM + TMA.
//
// Capability parity: reference BatchPrefillWith{Ragged,Paged}KVCacheWrapper.run
// (flashinfer/prefill.py:2184-2505, :3246-3555); the Blackwell bars to beat are the CUTLASS sm100
// FMHA (include/flashinfer/attention/blackwell/, ragged only) and the closed trtllm-gen cubins.
//
// B200-first design:
//  * work unit = (request, 256 query rows, q head); units are LPT-balanced over a persistent grid by
//    the C++ planner (runtime/planner.cpp: prefill_plan).
//  * 12 warps: warp0 TMA producer, warp1 tcgen05.mma issuer, warp2 TMEM allocator, warps 4-7 and 8-11
//    two softmax warpgroups, each owning one 128-row Q tile (ping-pong: while WG0 does softmax the
//    tensor core runs WG1's PV and next QK).
//  * TMEM map (512 cols): S0 | S1 | O0 | O1 (fp32, 128 cols each); P (bf16) is written back in place
//    over S and consumed as the A operand of the PV MMA straight from TMEM.
//  * O stays in TMEM across KV tiles; the running max is updated lazily (rescale only when it grows
//    by more than 2^8), so the common case does no O read-modify-write at all.
//  * paged KV is gathered by TMA page boxes (same scheme as decode_sm100.cu); ragged KV is one
//    big page.  Causal / sliding-window / soft-cap handled in the softmax pass; masks are only
//    evaluated on boundary tiles.
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>
#include <type_traits>

using namespace fib200;

#ifndef FIB_POD_TU
FIB_EXPORT_LAST_ERROR()
#endif

namespace {

constexpr int kTileQ = 128;    // rows per softmax warpgroup
constexpr int kUnitQ = 256;    // rows per work unit (two Q tiles)
constexpr int kTileKV = 128;
constexpr int kWorkInts = 8;
constexpr int D = 128;  // head_dim_vo (and the default head_dim_qk)

struct PrefillParams {
  void* out;
  float* lse;
  const int32_t* kv_indices;      // paged only
  const int32_t* kv_page_indptr;  // paged only [B+1]
  const int32_t* work_info;       // [nwork][8] {req, q0, rows, head, kv_len, q_len, qo_start, kv_start}
  const int32_t* cta_work_indptr;
  int64_t o_stride_n, o_stride_h;
  int num_qo_heads, group, page_size, layout_hnd, paged;
  int window_left, causal;
  float sm_scale_log2, soft_cap, sm_scale;
};

// DQK = head_dim_qk (128, or 192 for DeepSeek-style prefill with head_dim_vo = 128).  The 192 variant trades pipeline
// depth for the larger Q / K tiles (2 K stages, 1 V stage) to stay inside 227 KB.
template <int DQK>
struct Smem {
  static constexpr int kStagesK = DQK == 128 ? 3 : 2, kStagesV = DQK == 128 ? 2 : 1;
  static constexpr int kQChunks = DQK / 64;
  static constexpr int kQTileBytes = kTileQ * DQK * 2;   // 32 / 48 KB
  static constexpr int kKTileBytes = kTileKV * DQK * 2;
  static constexpr int kVTileBytes = kTileKV * D * 2;    // 32 KB
  static constexpr int kChunkBytes = kTileKV * 128;      // 16 KB (64 columns)
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = 2 * kQTileBytes;
  static constexpr int kOffV = kOffK + kStagesK * kKTileBytes;
  static constexpr int kOffBar = kOffV + kStagesV * kVTileBytes;
  static constexpr int kNumBars = 2 * kStagesK + 2 * kStagesV + 2 /*q_full,q_empty*/ + 2 /*s_full*/ + 2 /*p_ready*/ +
                                  2 /*o_done*/ + 2 /*o_free*/;
  static constexpr int kTotal = kOffBar + kNumBars * 8 + 16 + 1024;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---- softmax passes, specialised on (mask, soft-cap) so the common tile is pure FFMA/EX2/FADD/CVT ----
template <bool kMask, bool kCap>
__device__ __forceinline__ float row_max_pass(uint32_t s_tmem, int kv0, int lo, int hi, float cap, float cap_scale) {
  float tmax = -INFINITY;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    ptx::tmem_ld_x32(s_tmem + c * 32, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      float x = __uint_as_float(r[e]);
      if constexpr (kCap) x = cap * ptx::tanh_approx(x * cap_scale);
      if constexpr (kMask) {
        const int kvp = kv0 + c * 32 + e;
        if (kvp > hi || kvp < lo) x = -INFINITY;
      }
      tmax = fmaxf(tmax, x);
    }
  }
  return tmax;
}

template <bool kMask, bool kCap, bool kBf16>
__device__ __forceinline__ float exp_pass(uint32_t s_tmem, int kv0, int lo, int hi, float cap, float cap_scale,
                                          float scale, float m_ref) {
  float l = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    ptx::tmem_ld_x32(s_tmem + c * 32, r);
    ptx::tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
      float x0 = __uint_as_float(r[e]), x1 = __uint_as_float(r[e + 1]);
      if constexpr (kCap) {
        x0 = cap * ptx::tanh_approx(x0 * cap_scale);
        x1 = cap * ptx::tanh_approx(x1 * cap_scale);
      }
      float p0 = ptx::ex2(fmaf(x0, scale, -m_ref));
      float p1 = ptx::ex2(fmaf(x1, scale, -m_ref));
      if constexpr (kMask) {
        const int kvp = kv0 + c * 32 + e;
        if (kvp > hi || kvp < lo) p0 = 0.f;
        if (kvp + 1 > hi || kvp + 1 < lo) p1 = 0.f;
      }
      l += p0 + p1;
      pk[e / 2] = kBf16 ? pack_bf16x2(p0, p1) : pack_f16x2(p0, p1);
    }
    ptx::tmem_st_x16(s_tmem + c * 16, pk);
  }
  return l;
}

// The kernel body is a device function of the CTA index so that the fused POD kernel (pod_sm100.cu) can run it on a subset
// of the grid next to the decode body; tensor maps are passed by reference to the __grid_constant__ kernel parameters.
template <typename T, int DQK>
__device__ __forceinline__ void prefill_body(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                                             const PrefillParams& p, uint32_t idesc_qk, uint32_t idesc_pv, int cta) {
  using S = Smem<DQK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
  uint64_t* k_full = bars;
  uint64_t* k_empty = k_full + S::kStagesK;
  uint64_t* v_full = k_empty + S::kStagesK;
  uint64_t* v_empty = v_full + S::kStagesV;
  uint64_t* q_full = v_empty + S::kStagesV;  // [1]
  uint64_t* q_empty = q_full + 1;            // [1]
  uint64_t* s_full = q_empty + 1;            // [2]
  uint64_t* p_ready = s_full + 2;            // [2]
  uint64_t* o_done = p_ready + 2;            // [2]
  uint64_t* o_free = o_done + 2;             // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int w_begin = p.cta_work_indptr[cta];
  const int w_end = p.cta_work_indptr[cta + 1];
  if (w_begin >= w_end) return;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmQ);
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
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_ready[i], 128);
      ptx::mbar_init(&o_done[i], 1);
      ptx::mbar_init(&o_free[i], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_ptr, 512);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int ps = p.page_size;

  ptx::grid_dep_wait();

  // kv tile range of a unit (shared by every role)
  auto unit_tiles = [&](const int32_t* wi, int& t_lo, int& t_hi) {
    const int q0 = wi[1], rows = wi[2], kv_len = wi[4], q_len = wi[5];
    int kv_hi = kv_len;
    if (p.causal) kv_hi = min(kv_len, kv_len - q_len + q0 + rows);
    if (kv_hi < 0) kv_hi = 0;
    int kv_lo = 0;
    if (p.window_left >= 0) kv_lo = max(0, kv_len - q_len + q0 - p.window_left);
    t_lo = kv_lo / kTileKV;
    t_hi = (kv_hi + kTileKV - 1) / kTileKV;
    if (t_hi < t_lo) t_hi = t_lo;
  };

  if (warp < 4) {
    ptx::setmaxnreg_dec<80>();
    if (warp == 2) {
      // ============================ Q loader ============================
      uint32_t qph = 0;
      for (int w = w_begin; w < w_end; ++w) {
        const int32_t* wi = p.work_info + w * kWorkInts;
        const int q0 = wi[1], head = wi[3], qo_start = wi[6];
        if (lane == 0) {
          ptx::mbar_wait(q_empty, qph ^ 1);
          ptx::mbar_arrive_expect_tx(q_full, 2 * S::kQTileBytes);
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int c = 0; c < S::kQChunks; ++c)
              ptx::tma_load_3d(smem + S::kOffQ + t * S::kQTileBytes + c * S::kChunkBytes, &tmQ, q_full, c * 64, head,
                               qo_start + q0 + t * kTileQ, ptx::kEvictFirst);
        }
        qph ^= 1;
      }
    } else if (warp == 0 || warp == 3) {
      // ============================ K loader (warp 0) / V loader (warp 3) ============================
      // (one warp per tensor: paged gathers issue one TMA per page and UTMALDG issue is the limiter)
      const int is_v = (warp == 3);
      const int nstages = is_v ? S::kStagesV : S::kStagesK;
      uint64_t* full_bars = is_v ? v_full : k_full;
      uint64_t* empty_bars = is_v ? v_empty : k_empty;
      const CUtensorMap* tm = is_v ? &tmV : &tmK;
      uint8_t* ring = smem + (is_v ? S::kOffV : S::kOffK);
      const int tile_bytes = is_v ? S::kVTileBytes : S::kKTileBytes;
      const int nchunks = is_v ? 2 : S::kQChunks;
      int st = 0;
      uint32_t ph = 0;
      for (int w = w_begin; w < w_end; ++w) {
        const int32_t* wi = p.work_info + w * kWorkInts;
        const int req = wi[0], head = wi[3], kv_start = wi[7];
        const int kv_head = head / p.group;
        int t_lo, t_hi;
        unit_tiles(wi, t_lo, t_hi);
        int page_start = 0, num_pages = 0;
        if (p.paged) {
          page_start = p.kv_page_indptr[req];
          num_pages = p.kv_page_indptr[req + 1] - page_start;
        }
        for (int ti = t_lo; ti < t_hi; ++ti) {
          // geometry: ragged = one giant page starting at kv_start; paged = page boxes
          int n_boxes, box_rows, first_page = 0, page_off = 0;
          uint32_t box_bytes;
          if (!p.paged) {
            n_boxes = 1;
            box_rows = 0;
            page_off = kv_start + ti * kTileKV;
            box_bytes = kTileKV * 128;
          } else if (ps <= kTileKV) {
            const int ppt = kTileKV / ps;
            first_page = ti * ppt;
            n_boxes = min(ppt, num_pages - first_page);
            box_rows = ps;
            box_bytes = ps * 128;
          } else {
            const int tpp = ps / kTileKV;  // page_size > 128 must be a multiple of 128 for prefill
            first_page = ti / tpp;
            page_off = (ti % tpp) * kTileKV;
            n_boxes = 1;
            box_rows = 0;
            box_bytes = kTileKV * 128;
          }
          const uint32_t tx = uint32_t(n_boxes) * box_bytes * uint32_t(nchunks);
          int my_pages[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int bi = lane + j * 32;
            if (!p.paged)
              my_pages[j] = (bi == 0) ? 0 : -1;
            else
              my_pages[j] = (bi < n_boxes) ? __ldg(p.kv_indices + page_start + first_page + bi) : -1;
          }
          if (lane == 0) {
            ptx::mbar_wait(&empty_bars[st], ph ^ 1);
            ptx::mbar_arrive_expect_tx(&full_bars[st], tx);
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (my_pages[j] >= 0) {
              uint8_t* dst = ring + st * tile_bytes + (lane + j * 32) * box_rows * 128;
              for (int c = 0; c < nchunks; ++c) {
                if (p.layout_hnd)
                  ptx::tma_load_4d(dst + c * S::kChunkBytes, tm, &full_bars[st], c * 64, page_off, kv_head, my_pages[j]);
                else
                  ptx::tma_load_4d(dst + c * S::kChunkBytes, tm, &full_bars[st], c * 64, kv_head, page_off, my_pages[j]);
              }
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
      uint32_t kph = 0, vph = 0, qph = 0;
      uint32_t s_cnt[2] = {0, 0};   // number of QK issued per tile (for nothing but symmetry)
      uint32_t p_cnt[2] = {0, 0};   // P_ready phases consumed
      uint32_t of_cnt[2] = {0, 0};  // o_free phases consumed
      (void)s_cnt;
      const uint32_t q_addr = ptx::smem_u32(smem + S::kOffQ);
      auto issue_qk = [&](int t, uint32_t k_addr) {
        if (ptx::elect_one()) {
          const uint32_t d_tmem = tmem_base + t * 128;
#pragma unroll
          for (int k = 0; k < DQK / 16; ++k) {
            const uint32_t off = (k / 4) * S::kChunkBytes + (k % 4) * 32;
            const uint64_t da = ptx::make_smem_desc(q_addr + t * S::kQTileBytes + off, 16, 1024, ptx::kSwz128);
            const uint64_t db = ptx::make_smem_desc(k_addr + off, 16, 1024, ptx::kSwz128);
            ptx::mma_f16_ss<1>(d_tmem, da, db, idesc_qk, k > 0 ? 1u : 0u);
          }
          ptx::mma_commit(&s_full[t]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, uint32_t v_addr, bool first_tile) {
        if (ptx::elect_one()) {
          const uint32_t d_tmem = tmem_base + 256 + t * 128;
          const uint32_t a_tmem = tmem_base + t * 128;
          const uint64_t db = ptx::make_smem_desc(v_addr, S::kChunkBytes, 1024, ptx::kSwz128);
#pragma unroll
          for (int k = 0; k < kTileKV / 16; ++k)
            ptx::mma_f16_ts<1>(d_tmem, a_tmem + k * 8, ptx::desc_advance(db, k * 16 * 128), idesc_pv,
                               (first_tile && k == 0) ? 0u : 1u);
          ptx::mma_commit(&o_done[t]);
        }
        __syncwarp();
      };
      for (int w = w_begin; w < w_end; ++w) {
        const int32_t* wi = p.work_info + w * kWorkInts;
        int t_lo, t_hi;
        unit_tiles(wi, t_lo, t_hi);
        const int n = t_hi - t_lo;
        ptx::mbar_wait(q_full, qph);
        qph ^= 1;
        if (n == 0) {
          // nothing to attend to: release Q, tell softmax WGs via s_full with zero tiles is not needed
          if (ptx::elect_one()) ptx::mma_commit(q_empty);
          __syncwarp();
          continue;
        }
        // O accumulators must have been drained by the previous unit's epilogue
        for (int t = 0; t < 2; ++t) {
          if (of_cnt[t] > 0) ptx::mbar_wait(&o_free[t], (of_cnt[t] - 1) & 1);
        }
        // prologue: QK(0,0), QK(1,0)
        ptx::mbar_wait(&k_full[ks], kph);
        ptx::tc_fence_after();
        {
          const uint32_t k_addr = ptx::smem_u32(smem + S::kOffK + ks * S::kKTileBytes);
          issue_qk(0, k_addr);
          issue_qk(1, k_addr);
          if (ptx::elect_one()) {
            ptx::mma_commit(&k_empty[ks]);
            if (n == 1) ptx::mma_commit(q_empty);
          }
          __syncwarp();
          if (++ks == S::kStagesK) { ks = 0; kph ^= 1; }
        }
        for (int j = 0; j < n; ++j) {
          const bool has_next = (j + 1 < n);
          ptx::mbar_wait(&v_full[vs], vph);
          const uint32_t v_addr = ptx::smem_u32(smem + S::kOffV + vs * S::kVTileBytes);
          uint32_t k_addr = 0;
          if (has_next) {
            ptx::mbar_wait(&k_full[ks], kph);
            k_addr = ptx::smem_u32(smem + S::kOffK + ks * S::kKTileBytes);
          }
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            ptx::mbar_wait(&p_ready[t], p_cnt[t] & 1);
            ++p_cnt[t];
            ptx::tc_fence_after();
            issue_pv(t, v_addr, j == 0);
            if (has_next) issue_qk(t, k_addr);
          }
          if (ptx::elect_one()) {
            ptx::mma_commit(&v_empty[vs]);
            if (has_next) {
              ptx::mma_commit(&k_empty[ks]);
              if (j + 2 == n) ptx::mma_commit(q_empty);  // last QK of the unit issued
            }
          }
          __syncwarp();
          if (++vs == S::kStagesV) { vs = 0; vph ^= 1; }
          if (has_next) {
            if (++ks == S::kStagesK) { ks = 0; kph ^= 1; }
          }
        }
        ++of_cnt[0];
        ++of_cnt[1];
      }
    }
  } else {
    // ============================ softmax warpgroups ============================
    ptx::setmaxnreg_inc<208>();
    const int t = (warp - 4) >> 2;        // Q tile owned by this warpgroup
    const int q4 = warp & 3;              // TMEM lane quadrant
    const int row = q4 * 32 + lane;       // row inside the Q tile
    const uint32_t lane_addr = uint32_t(q4 * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_addr + t * 128;
    const uint32_t o_tmem = tmem_base + lane_addr + 256 + t * 128;
    uint32_t s_cnt = 0, od_cnt = 0;
    int vs = 0;
    uint32_t vph = 0;
    T* obase = reinterpret_cast<T*>(p.out);
    constexpr bool kIsBf16 = std::is_same<T, __nv_bfloat16>::value;

    for (int w = w_begin; w < w_end; ++w) {
      const int32_t* wi = p.work_info + w * kWorkInts;
      const int q0 = wi[1], rows = wi[2], head = wi[3], kv_len = wi[4], q_len = wi[5], qo_start = wi[6];
      int t_lo, t_hi;
      unit_tiles(wi, t_lo, t_hi);
      const int n = t_hi - t_lo;
      const int q_row = q0 + t * kTileQ + row;          // row index inside the request
      const bool row_valid = (t * kTileQ + row) < rows;
      const int q_pos = kv_len - q_len + q_row;         // absolute position of this query token
      T* orow = obase + int64_t(qo_start + q_row) * p.o_stride_n + int64_t(head) * p.o_stride_h;
      if (n == 0) {
        if (row_valid) {
#pragma unroll
          for (int c = 0; c < D; c += 8) *reinterpret_cast<int4*>(orow + c) = make_int4(0, 0, 0, 0);
          if (p.lse) p.lse[int64_t(qo_start + q_row) * p.num_qo_heads + head] = -INFINITY;
        }
        continue;
      }
      float m_used = -INFINITY;  // log2-domain reference max actually used for exp2
      float l = 0.f;
      const int first_row_pos = kv_len - q_len + q0 + t * kTileQ;  // q_pos of row 0 of this tile

      for (int j = 0; j < n; ++j) {
        const int kv0 = (t_lo + j) * kTileKV;
        bool need_mask = (kv0 + kTileKV > kv_len);
        if (p.causal) need_mask = need_mask || (kv0 + kTileKV - 1 > first_row_pos);
        if (p.window_left >= 0) need_mask = need_mask || (kv0 < first_row_pos + kTileQ - 1 - p.window_left);
        int hi = kv_len - 1;                       // highest visible kv index for this row
        if (p.causal) hi = min(hi, q_pos);
        const int lo = (p.window_left >= 0) ? q_pos - p.window_left : 0;

        ptx::mbar_wait(&s_full[t], s_cnt & 1);
        ++s_cnt;
        ptx::tc_fence_after();

        // ---- pass 1: row max ----
        const bool cap = p.soft_cap > 0.f;
        const float cap_scale = cap ? p.sm_scale / p.soft_cap : 0.f;
        float tmax;
        if (!cap) {
          tmax = need_mask ? row_max_pass<true, false>(s_tmem, kv0, lo, hi, 0.f, 0.f)
                           : row_max_pass<false, false>(s_tmem, kv0, lo, hi, 0.f, 0.f);
        } else {
          tmax = need_mask ? row_max_pass<true, true>(s_tmem, kv0, lo, hi, p.soft_cap, cap_scale)
                           : row_max_pass<false, true>(s_tmem, kv0, lo, hi, p.soft_cap, cap_scale);
        }
        const float scale = (p.soft_cap > 0.f) ? 1.4426950408889634f : p.sm_scale_log2;
        const float m_tile = tmax * scale;  // scale > 0
        // ---- lazy rescale: only when the max grew by more than 2^8 ----
        bool grow = (m_tile > m_used + 8.f) || (m_used == -INFINITY && m_tile > -INFINITY);
        if (j > 0 && __any_sync(0xffffffffu, grow && l > 0.f)) {
          // O must be stable: wait for PV(t, j-1)
          ptx::mbar_wait(&o_done[t], (od_cnt - 1) & 1);
          ptx::tc_fence_after();
          const float alpha = (grow && m_used > -INFINITY) ? ptx::ex2(m_used - m_tile) : 1.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            ptx::tmem_ld_x32(o_tmem + c * 32, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
            ptx::tmem_st_x32(o_tmem + c * 32, r);
          }
          ptx::tmem_st_wait();
          l *= alpha;
        } else if (grow && m_used > -INFINITY) {
          // j == 0 cannot get here (m_used == -inf); rows with l == 0 need no O fix-up
          l *= ptx::ex2(m_used - m_tile);
        }
        if (grow) m_used = m_tile;
        const float m_ref = (m_used == -INFINITY) ? 0.f : m_used;

        // ---- pass 2: P = exp2(S*scale - m), row sum, write P (bf16) in place over S ----
        if (!cap) {
          l += need_mask ? exp_pass<true, false, kIsBf16>(s_tmem, kv0, lo, hi, 0.f, 0.f, scale, m_ref)
                         : exp_pass<false, false, kIsBf16>(s_tmem, kv0, lo, hi, 0.f, 0.f, scale, m_ref);
        } else {
          l += need_mask ? exp_pass<true, true, kIsBf16>(s_tmem, kv0, lo, hi, p.soft_cap, cap_scale, scale, m_ref)
                         : exp_pass<false, true, kIsBf16>(s_tmem, kv0, lo, hi, p.soft_cap, cap_scale, scale, m_ref);
        }
        ptx::tmem_st_wait();
        // paged caches may hold uninitialised rows past kv_len: zero them so 0 * garbage != NaN
        if (t == 0 && p.paged && kv0 + kTileKV > kv_len) {
          ptx::mbar_wait(&v_full[vs], vph);
          if (kv0 + row >= kv_len) {
            uint8_t* vrow = smem + S::kOffV + vs * S::kVTileBytes + row * 128;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
              for (int e = 0; e < 8; ++e) reinterpret_cast<int4*>(vrow + c * S::kChunkBytes)[e] = make_int4(0, 0, 0, 0);
          }
          ptx::fence_proxy_async_smem();
        }
        if (++vs == S::kStagesV) { vs = 0; vph ^= 1; }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&p_ready[t]);
        ++od_cnt;  // PV(t, j) will complete phase od_cnt-1 of o_done[t]
      }

      // ---- epilogue: wait last PV, normalise, store ----
      ptx::mbar_wait(&o_done[t], (od_cnt - 1) & 1);
      ptx::tc_fence_after();
      const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_x32(o_tmem + c * 32, r);
        ptx::tmem_ld_wait();
        if (row_valid) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            Vec16<T> v;
#pragma unroll
            for (int u = 0; u < 8; ++u) v.v[u] = from_f32<T>(__uint_as_float(r[e + u]) * inv);
            st16(orow + c * 32 + e, v);
          }
        }
      }
      if (row_valid && p.lse)
        p.lse[int64_t(qo_start + q_row) * p.num_qo_heads + head] = l > 0.f ? m_used + ptx::lg2(l) : -INFINITY;
      ptx::tc_fence_before();
      ptx::mbar_arrive(&o_free[t]);
    }
  }

  ptx::grid_dep_launch();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, 512);
  }
}

template <typename T, int DQK>
__global__ void __launch_bounds__(384, 1)
prefill_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const PrefillParams p, uint32_t idesc_qk, uint32_t idesc_pv) {
  prefill_body<T, DQK>(tmQ, tmK, tmV, p, idesc_qk, idesc_pv, blockIdx.x);
}

}  // namespace

extern "C" int prefill_run(void* q, void* k, void* v, void* out, void* lse, void* kv_indices, void* kv_page_indptr,
                           void* work_info, void* cta_work_indptr, int64_t grid, int64_t total_q, int64_t num_qo_heads,
                           int64_t num_kv_heads, int64_t head_dim, int64_t paged, int64_t page_size,
                           int64_t num_pages_total, int64_t kv_stride_page, int64_t kv_stride_n, int64_t kv_stride_h,
                           int64_t v_stride_page, int64_t v_stride_n, int64_t v_stride_h, int64_t layout_hnd, int64_t q_stride_n, int64_t q_stride_h, int64_t o_stride_n,
                           int64_t o_stride_h, double sm_scale, double soft_cap, int64_t window_left, int64_t causal,
                           int64_t dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(head_dim == 128 || head_dim == 192, "prefill_sm100: head_dim_qk must be 128 or 192 (head_dim_vo = 128)");
  FIB_CHECK(dtype == kF16 || dtype == kBF16, "prefill_sm100: dtype must be f16/bf16");
  FIB_CHECK(!paged || page_size <= 128 || page_size % 128 == 0, "prefill_sm100: page_size > 128 must be a multiple of 128");
  FIB_CHECK(q_stride_n % 8 == 0 && q_stride_h % 8 == 0, "q strides must be 16B multiples");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const CUtensorMapDataType dt = dtype == kF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmQ, tmK, tmV;
  {
    // q [total_q, Hq, D] -> dims (D, Hq, total_q); rows past total_q are zero-filled by TMA
    uint64_t dims[3] = {(uint64_t)head_dim, (uint64_t)num_qo_heads, (uint64_t)total_q};
    uint64_t str[2] = {(uint64_t)q_stride_h * 2, (uint64_t)q_stride_n * 2};
    uint32_t box[3] = {64, 1, (uint32_t)kTileQ};
    if (make_tmap(&tmQ, dt, 3, q, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  const uint32_t box_rows = (!paged || page_size > kTileKV) ? (uint32_t)kTileKV : (uint32_t)page_size;
  for (int i = 0; i < 2; ++i) {
    const void* base = i == 0 ? k : v;
    CUtensorMap* tm = i == 0 ? &tmK : &tmV;
    const uint64_t hd = i == 0 ? (uint64_t)head_dim : (uint64_t)D;
    const int64_t sp = i == 0 ? kv_stride_page : v_stride_page, sn = i == 0 ? kv_stride_n : v_stride_n,
                  sh = i == 0 ? kv_stride_h : v_stride_h;
    if (layout_hnd && paged) {
      uint64_t dims[4] = {hd, (uint64_t)page_size, (uint64_t)num_kv_heads, (uint64_t)num_pages_total};
      uint64_t str[3] = {(uint64_t)sn * 2, (uint64_t)sh * 2, (uint64_t)sp * 2};
      uint32_t box[4] = {64, box_rows, 1, 1};
      if (make_tmap(tm, dt, 4, base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    } else {
      // NHD pages, or ragged [nnz, H, D] as one giant NHD page (page_size := nnz, 1 page)
      const uint64_t psz = paged ? (uint64_t)page_size : (uint64_t)num_pages_total;
      const uint64_t npg = paged ? (uint64_t)num_pages_total : 1;
      uint64_t dims[4] = {hd, (uint64_t)num_kv_heads, psz, npg};
      uint64_t str[3] = {(uint64_t)sh * 2, (uint64_t)sn * 2, (uint64_t)(paged ? sp : sn * (int64_t)psz) * 2};
      uint32_t box[4] = {64, 1, box_rows, 1};
      if (make_tmap(tm, dt, 4, base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    }
  }
  PrefillParams p;
  p.out = out;
  p.lse = (float*)lse;
  p.kv_indices = (const int32_t*)kv_indices;
  p.kv_page_indptr = (const int32_t*)kv_page_indptr;
  p.work_info = (const int32_t*)work_info;
  p.cta_work_indptr = (const int32_t*)cta_work_indptr;
  p.o_stride_n = o_stride_n;
  p.o_stride_h = o_stride_h;
  p.num_qo_heads = (int)num_qo_heads;
  p.group = (int)(num_qo_heads / num_kv_heads);
  p.page_size = (int)page_size;
  p.layout_hnd = (int)(layout_hnd && paged);
  p.paged = (int)paged;
  p.window_left = (int)window_left;
  p.causal = (int)causal;
  p.sm_scale = (float)sm_scale;
  p.sm_scale_log2 = (float)(sm_scale * 1.4426950408889634);
  p.soft_cap = (float)soft_cap;
  const bool f16 = dtype == kF16;
  const uint32_t fmt = f16 ? ptx::kFmtF16 : ptx::kFmtBF16;
  const uint32_t idesc_qk = ptx::make_idesc_f16(fmt, 128, 128, 0, 0);
  const uint32_t idesc_pv = ptx::make_idesc_f16(fmt, 128, 128, 0, 1);
  auto go = [&](auto kern, int smem_bytes) -> int {
#ifdef FIB_POD_TU
    // POD: park the fully-built prefill launch; the decode launcher that follows fuses both into one grid
    if (pod_stage().armed) {
      PodStage& st = pod_stage();
      static_assert(sizeof(PrefillParams) <= sizeof(st.params), "PodStage params buffer too small");
      st.tm[0] = tmQ;
      st.tm[1] = tmK;
      st.tm[2] = tmV;
      memcpy(st.params, &p, sizeof(p));
      st.idesc[0] = idesc_qk;
      st.idesc[1] = idesc_pv;
      st.grid = (int)grid;
      st.smem = smem_bytes;
      st.f16 = f16 ? 1 : 0;
      st.dqk = (int)head_dim;
      st.have_prefill = true;
      return 0;
    }
#endif
    FIB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    LaunchCfg lc(dim3((unsigned)grid), dim3(384), smem_bytes, stream, pdl != 0);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, tmQ, tmK, tmV, p, idesc_qk, idesc_pv));
    return 0;
  };
  if (head_dim == 128) {
    if (f16) return go(prefill_kernel<__half, 128>, Smem<128>::kTotal);
    return go(prefill_kernel<__nv_bfloat16, 128>, Smem<128>::kTotal);
  }
  if (f16) return go(prefill_kernel<__half, 192>, Smem<192>::kTotal);
  return go(prefill_kernel<__nv_bfloat16, 192>, Smem<192>::kTotal);
  return 0;
}
