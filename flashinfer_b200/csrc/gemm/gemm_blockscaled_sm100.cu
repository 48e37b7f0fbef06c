// Low-precision tcgen05 GEMMs for sm_100a:  C[b] = alpha * (A[b] * B[b]^T)   (A [M,K], B [N,K], both K-major)
//
//   kind 0  fp8 (e4m3 / e5m2) with per-tensor scales       tcgen05.mma.kind::f8f6f4
//   kind 1  mxfp8:  fp8 data + UE8M0 scale per 32 elems     kind::mxf8f6f4.block_scale (scale_vec::1X)
//   kind 2  nvfp4:  e2m1 data + UE4M3 scale per 16 elems    kind::mxf4nvf4.block_scale.scale_vec::4X
//   kind 3  mxfp4:  e2m1 data + UE8M0 scale per 32 elems    kind::mxf4nvf4.block_scale.scale_vec::2X
//
// Parity: reference mm_fp8 / bmm_fp8 (flashinfer/gemm/gemm_base.py:2262-2533), mm_mxfp8 / bmm_mxfp8 (:3003-3420),
// mm_fp4 (:4567-4839) and the CUTLASS / cuDNN / trtllm-gen back ends behind them.
//
// Design: persistent warp-specialised kernel (TMA producer warp, single-thread MMA issuer, 4 epilogue warps).
// Every smem stage carries one 128-byte-wide K slab of A and B (128 fp8 or 256 fp4 elements = 4 MMAs) plus the
// scale factors of that slab, which live in global memory in the 128x4 "swizzled" layout (512-byte blocks holding
// 128 rows x 4 scale bytes: byte (m%32)*16 + (m%128/32)*4 + k%4).  A 512-byte block is exactly one
// tcgen05.cp.32x128b.warpx4 source, so the scales go gmem -(bulk copy)-> smem -(tcgen05.cp)-> TMEM with no
// register traffic; tcgen05.cp and tcgen05.mma execute in issue order, so one TMEM scale buffer suffices.
// The N tile width is a run-time multiple of 32: a tile that does not start on a 128-row scale block simply
// offsets the TMEM column of SFB by (n0 % 128) / 32 (tiles are multiples of 64 rows -> offsets 0 or 2).
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

constexpr int BM = 128;
constexpr int BKB = 128;  // K slab in bytes per stage (one SWIZZLE_128B row)

enum Kind { kFp8 = 0, kMxFp8 = 1, kNvFp4 = 2, kMxFp4 = 3, kDense16 = 4 /* bf16 / f16 operands, CTA-pair kernel only */ };
__host__ __device__ constexpr bool kind_scaled(int k) { return k == kMxFp8 || k == kNvFp4 || k == kMxFp4; }

struct Geo {
  int stages, stage_bytes, a_bytes, b_bytes, sfa_bytes, sfb_bytes, bar_offset, total;
  int nchunk;      // 512-byte scale blocks per row block per K slab
  int rb;          // 128-row scale blocks that can overlap one N tile
  int acc_stages;  // TMEM accumulator buffers
  int sfa_col, sfb_col, tmem_cols;
  __host__ __device__ static Geo make(int BN, int kind) {
    Geo g;
    g.nchunk = kind == kNvFp4 ? 4 : (kind == kMxFp4 ? 2 : (kind == kMxFp8 ? 1 : 0));
    g.rb = (BN % 128 == 0) ? BN / 128 : (BN + 64 + 127) / 128;
    g.a_bytes = BM * BKB;
    g.b_bytes = BN * BKB;
    g.sfa_bytes = g.nchunk * 512;
    g.sfb_bytes = g.rb * g.nchunk * 512;
    g.stage_bytes = (g.a_bytes + g.b_bytes + g.sfa_bytes + g.sfb_bytes + 1023) / 1024 * 1024;  // SWIZZLE_128B tiles need 1 KB alignment
    int st = (218 * 1024) / g.stage_bytes;
    g.stages = st > 8 ? 8 : st;
    g.bar_offset = g.stages * g.stage_bytes;
    g.total = g.bar_offset + 320 + 1024;
    const int sf_cols = g.nchunk * 4 + g.rb * g.nchunk * 4;
    g.acc_stages = (2 * BN + sf_cols <= 512) ? 2 : 1;
    g.sfa_col = g.acc_stages * BN;
    g.sfb_col = g.sfa_col + g.nchunk * 4;
    int need = g.sfb_col + g.rb * g.nchunk * 4;
    int c = 32;
    while (c < need) c <<= 1;
    g.tmem_cols = c;
    return g;
  }
};

struct Params {
  const uint8_t* sfa;
  const uint8_t* sfb;
  const float* alpha_a;  // optional device scalars; alpha = alpha_a * alpha_b
  const float* alpha_b;
  int64_t sfa_batch_stride, sfb_batch_stride, c_batch_stride, ldc;
  int M, N, Kb;  // Kb = K in bytes
  int batch, BN;
  const int32_t* tile_expert;  // grouped (MoE) mode: expert of every 128-row tile of A (-1 = skip); B / SFB / alpha are per expert
  const int32_t* meta;         // grouped mode: meta[0] = number of live row tiles (device side)
  const int32_t* row_map;      // grouped mode (optional): rows with row_map < 0 are padding, their results are not stored
  int a_box_rows;              // rows per A TMA box: 128, or 32 in grouped mode with a row_map (only live row groups are loaded)
  int tab_tiles, tab_experts;  // capacity of the shared-memory tile->expert / alpha tables (grouped mode)
  int split;          // cluster split-K factor (1 or 2): both CTAs of a cluster own the same tile, half of K each
  int sf_k_tiles;     // 512-byte blocks along K in the scale tensors
  int sfb_row_tiles;  // 128-row blocks in SFB
  uint32_t idesc;
};

template <int KIND, typename OutT>
__global__ void __launch_bounds__(256, 1)
bs_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, OutT* __restrict__ C,
               const Params p) {
  const Geo G = Geo::make(p.BN, KIND);
  const int BN = p.BN;
  const int kStages = G.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G.bar_offset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* go_bar = tmem_empty + 2;       // split-K: leader's pipeline smem is free, peer may deposit its partial
  uint64_t* partials_bar = go_bar + 1;     // split-K: peer's partial has landed in the leader's smem
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(partials_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = p.split;
  const int crank = split > 1 ? int(ptx::cluster_ctarank()) : 0;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 4);
    }
    ptx::mbar_init(go_bar, 1);
    ptx::mbar_init(partials_bar, 1);
    ptx::fence_mbar_init();
  }
  if constexpr (KIND != kFp8) {
    // scale staging areas start finite (zero): slabs past the K / N edge are never loaded but still multiplied
    for (int s = 0; s < kStages; ++s) {
      uint32_t* sf = reinterpret_cast<uint32_t*>(smem + s * G.stage_bytes + G.a_bytes + G.b_bytes);
      for (int i = threadIdx.x; i < (G.sfa_bytes + G.sfb_bytes) / 4; i += blockDim.x) sf[i] = 0;
    }
    ptx::fence_proxy_async_smem();
  }
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_ptr, G.tmem_cols);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (split > 1) ptx::cluster_sync();  // the peer's mbarriers must exist before any remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  ptx::grid_dep_wait();
  ptx::grid_dep_launch();

  const bool grouped = p.tile_expert != nullptr;
  int tiles_m = (p.M + BM - 1) / BM;
  if (grouped && p.meta) tiles_m = min(tiles_m, p.meta[0]);  // written by the MoE sort kernel
  // grouped mode: the tile -> expert table and the per-expert alphas are staged in shared memory once; reading them from
  // global memory per tile put an L2 round trip (~1 us) on the critical path of every role (ncu: long-scoreboard stalls)
  int32_t* s_expert = reinterpret_cast<int32_t*>(smem + G.bar_offset + 320);
  float* s_alpha = reinterpret_cast<float*>(s_expert + p.tab_tiles);
  // MoE with few tokens per expert: most of a 128-row tile is padding.  The A tile is fetched in 32-row boxes and only the
  // boxes that contain live rows are loaded (live rows are a prefix of the tile): s_nbox[tile] in 1..4.  The stale rows left
  // in shared memory only feed accumulator rows that the epilogue never stores.  Per-SM ingest was the limiter of these
  // weight-streaming GEMMs (A was as many bytes as B), so this is worth ~25 % on the tiny-batch MoE shapes.
  int32_t* s_nbox = reinterpret_cast<int32_t*>(s_alpha + p.tab_experts);
  if (grouped) {
    for (int i = threadIdx.x; i < tiles_m && i < p.tab_tiles; i += blockDim.x) {
      s_expert[i] = p.tile_expert[i];
      int nb = BM / 32;
      if (p.a_box_rows == 32 && p.row_map) {
        nb = 1;
        for (int bx = 1; bx < BM / 32; ++bx)
          if (i * BM + bx * 32 < p.M && p.row_map[i * BM + bx * 32] >= 0) nb = bx + 1;
      }
      s_nbox[i] = nb;
    }
    if (p.alpha_a)
      for (int i = threadIdx.x; i < p.batch && i < p.tab_experts; i += blockDim.x) s_alpha[i] = p.alpha_a[i];
    __syncthreads();
  }
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_per_batch = tiles_m * tiles_n;
  const int num_tiles = tiles_per_batch * (grouped ? 1 : p.batch);
  const int num_kb_total = (p.Kb + BKB - 1) / BKB;
  // split-K: one tile per cluster, rank r multiplies K slabs [kb_lo, kb_hi)
  const int kb_lo = (crank * num_kb_total) / split, kb_hi = ((crank + 1) * num_kb_total) / split;
  const int t_first = split > 1 ? int(blockIdx.x) / split : int(blockIdx.x);
  const int t_stride = split > 1 ? num_tiles : int(gridDim.x);
  // tile -> (a_batch, b_batch, row tile, col tile).  grouped: n-fastest so consecutive CTAs share an expert's weights
  auto decode = [&](int t, int& ba, int& bb, int& tm, int& tn) -> bool {
    if (grouped) {
      tm = t / tiles_n;
      tn = t % tiles_n;
      ba = 0;
      bb = tm < p.tab_tiles ? s_expert[tm] : p.tile_expert[tm];
      return bb >= 0;
    }
    const int r = t % tiles_per_batch;
    ba = bb = t / tiles_per_batch;
    tm = r % tiles_m;
    tn = r / tiles_m;
    return true;
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_first; t < num_tiles; t += t_stride) {
        int b, bb, tm, tn;
        if (!decode(t, b, bb, tm, tn)) continue;
        const int n0 = tn * BN;
        const int rb0 = n0 / 128;
        int rbn = (n0 + BN + 127) / 128 - rb0;
        if (rb0 + rbn > p.sfb_row_tiles) rbn = p.sfb_row_tiles - rb0;
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * G.stage_bytes;
          uint8_t* sb = sa + G.a_bytes;
          const int abr = p.a_box_rows;
          const int nbox = abr == BM ? 1 : (tm < p.tab_tiles ? s_nbox[tm] : BM / 32);
          uint32_t tx = uint32_t(nbox * abr) * BKB + G.b_bytes;
          int nch = 0;
          if constexpr (KIND != kFp8) {
            nch = p.sf_k_tiles - kb * G.nchunk;
            nch = nch > G.nchunk ? G.nchunk : nch;
            tx += uint32_t(nch * 512) * uint32_t(1 + rbn);
          }
          ptx::mbar_arrive_expect_tx(&full_bar[stage], tx);
          for (int bx = 0; bx < nbox; ++bx)
            ptx::tma_load_3d(sa + bx * abr * BKB, &tmA, &full_bar[stage], kb * BKB, tm * BM + bx * abr, b, ptx::kEvictNormal);
          ptx::tma_load_3d(sb, &tmB, &full_bar[stage], kb * BKB, n0, bb, ptx::kEvictNormal);
          if constexpr (KIND != kFp8) {
            uint8_t* ssfa = sb + G.b_bytes;
            uint8_t* ssfb = ssfa + G.sfa_bytes;
            const uint8_t* ga = p.sfa + int64_t(b) * p.sfa_batch_stride +
                                (int64_t(tm) * p.sf_k_tiles + int64_t(kb) * G.nchunk) * 512;
            ptx::bulk_load(ssfa, ga, nch * 512, &full_bar[stage]);
            for (int rr = 0; rr < rbn; ++rr) {
              const uint8_t* gb = p.sfb + int64_t(bb) * p.sfb_batch_stride +
                                  (int64_t(rb0 + rr) * p.sf_k_tiles + int64_t(kb) * G.nchunk) * 512;
              ptx::bulk_load(ssfb + rr * G.nchunk * 512, gb, nch * 512, &full_bar[stage]);
            }
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = t_first; t < num_tiles; t += t_stride) {
      int b_, bb_, tm_, tn_;
      if (!decode(t, b_, bb_, tm_, tn_)) continue;
      const int n0 = tn_ * BN;
      const uint32_t sfb_off = uint32_t((n0 % 128) / 32);
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb_lo; kb < kb_hi; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + stage * G.stage_bytes);
          const uint32_t sb = sa + G.a_bytes;
          const uint64_t da = ptx::make_smem_desc(sa, 16, 1024, ptx::kSwz128);
          const uint64_t db = ptx::make_smem_desc(sb, 16, 1024, ptx::kSwz128);
          if constexpr (KIND != kFp8) {
            const uint32_t ssfa = sb + G.b_bytes, ssfb = ssfa + G.sfa_bytes;
            for (int c = 0; c < G.nchunk; ++c) {
              ptx::tmem_cp_32x128b_warpx4(tmem_base + G.sfa_col + c * 4,
                                          ptx::make_smem_desc(ssfa + c * 512, 0, 128, ptx::kSwzNone));
              for (int rr = 0; rr < G.rb; ++rr)
                ptx::tmem_cp_32x128b_warpx4(tmem_base + G.sfb_col + (c * G.rb + rr) * 4,
                                            ptx::make_smem_desc(ssfb + (rr * G.nchunk + c) * 512, 0, 128, ptx::kSwzNone));
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t accum = (kb > kb_lo || k > 0) ? 1u : 0u;
            const uint64_t dak = ptx::desc_advance(da, k * 32), dbk = ptx::desc_advance(db, k * 32);
            if constexpr (KIND == kFp8) {
              ptx::mma_f8f6f4_ss<1>(d_tmem, dak, dbk, p.idesc, accum);
            } else if constexpr (KIND == kMxFp8) {
              const uint32_t id = p.idesc | (uint32_t(k) << 4) | (uint32_t(k) << 29);
              ptx::mma_mxf8f6f4_ss(d_tmem, dak, dbk, id, tmem_base + G.sfa_col, tmem_base + G.sfb_col + sfb_off, accum);
            } else if constexpr (KIND == kNvFp4) {
              ptx::mma_mxf4nvf4_ss(d_tmem, dak, dbk, p.idesc, tmem_base + G.sfa_col + k * 4,
                                   tmem_base + G.sfb_col + k * G.rb * 4 + sfb_off, accum);
            } else {
              const uint32_t sid = uint32_t(k & 1) * 2;
              const uint32_t id = p.idesc | (sid << 4) | (sid << 29);
              ptx::mma_mxf4_2x_ss(d_tmem, dak, dbk, id, tmem_base + G.sfa_col + (k >> 1) * 4,
                                  tmem_base + G.sfb_col + (k >> 1) * G.rb * 4 + sfb_off, accum);
            }
          }
          ptx::mma_commit(&empty_bar[stage]);
          if (kb == kb_hi - 1) ptx::mma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (G.acc_stages == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    float alpha0 = 1.f;
    if (p.alpha_a && !grouped) alpha0 *= *p.alpha_a;
    if (p.alpha_b) alpha0 *= *p.alpha_b;
    for (int t = t_first; t < num_tiles; t += t_stride) {
      int b, bb, tm, tn;
      if (!decode(t, b, bb, tm, tn)) continue;
      const float alpha = (grouped && p.alpha_a) ? alpha0 * (bb < p.tab_experts ? s_alpha[bb] : p.alpha_a[bb]) : alpha0;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = tm * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + acc * BN + (uint32_t(q * 32) << 16);
      const int r_in_tile = q * 32 + lane;
      if (split > 1) {
        // exchange buffer = the LEADER's pipeline smem (idle once its accumulator is complete):
        // [4-float column group][row] x 16 B  -> conflict-free for both the remote stores and the local loads
        if (crank == 0) {
          if (threadIdx.x == 128) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(go_bar), 1));
          ptx::mbar_wait_cluster(partials_bar, 0);
        } else {
          ptx::mbar_wait_cluster(go_bar, 0);
          const uint32_t remote = ptx::mapa(ptx::smem_u32(smem), 0);
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            ptx::tmem_ld_x32(taddr + c0, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              ptx::st_dsmem_v4(remote + uint32_t(((c0 + j) / 4 * BM + r_in_tile) * 16),
                               make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                           __uint_as_float(v[j + 3])));
          }
          ptx::named_bar_sync(1, 128);
          if (threadIdx.x == 128) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(partials_bar), 0));
          continue;  // only the leader writes the tile
        }
      }
      OutT* crow = C + int64_t(b) * p.c_batch_stride + int64_t(row) * p.ldc;
      const bool vec_ok = (p.ldc % (16 / sizeof(OutT)) == 0);
      const bool row_live = !p.row_map || (row < p.M && p.row_map[row] >= 0);
      const bool warp_live = __any_sync(0xffffffffu, row_live);  // a warp of pure padding rows skips its TMEM reads too
#pragma unroll 1
      for (int c0 = 0; c0 < BN && warp_live; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_x32(taddr + c0, v);
        ptx::tmem_ld_wait();
        if (split > 1) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 o4 = *reinterpret_cast<const float4*>(smem + ((c0 + j) / 4 * BM + r_in_tile) * 16);
            v[j] = __float_as_uint(__uint_as_float(v[j]) + o4.x);
            v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + o4.y);
            v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + o4.z);
            v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + o4.w);
          }
        }
        const int col0 = tn * BN + c0;
        if (row < p.M && row_live) {
          if (col0 + 32 <= p.N && vec_ok) {
            constexpr int VN = 16 / sizeof(OutT);
#pragma unroll
            for (int j = 0; j < 32; j += VN) {
              Vec16<OutT> o;
#pragma unroll
              for (int e = 0; e < VN; ++e) o.v[e] = from_f32<OutT>(__uint_as_float(v[j + e]) * alpha);
              st16(crow + col0 + j, o);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) crow[col0 + j] = from_f32<OutT>(__uint_as_float(v[j]) * alpha);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (G.acc_stages == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (split > 1) ptx::cluster_sync();  // nobody leaves while the peer may still touch its shared memory
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, G.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fp8 GEMM with fp32 group scales (DeepSeek-V3 style: 1x128 scales on A, 128x128 on B).  fp32 scales cannot be fed to
// the block-scaled tensor-core path, so every 128-wide K slab is multiplied into a fresh TMEM buffer (2 buffers,
// ping-pong) and the epilogue warps promote it into fp32 register accumulators:  acc += sa[row,kb] * sb[nblk,kb] * D.
// The promotion of slab kb overlaps the MMAs of slab kb+1.  Parity: reference gemm_fp8_nt_groupwise
// (include/flashinfer/gemm/gemm_groupwise_sm100.cuh:40-140) / DeepGEMM's per-block promotion.
// ---------------------------------------------------------------------------------------------------------------
struct GwParams {
  const float* sa;  // A scales, element strides (row, kblock)
  const float* sb;  // B scales, element strides (n-block of 128, kblock)
  int64_t sa_row, sa_k, sb_n, sb_k, ldc;
  int M, N, K;
  uint32_t idesc;
  // m-grouped contiguous mode (MoE / DeepGEMM): A [M, K] is tile-aligned per expert, B is [E * N, K], row tile tm uses the
  // weights + scales of expert tile_expert[tm]; meta[0] = number of live row tiles (device-side, produced by the sort).
  const int32_t* tile_expert;
  const int32_t* meta;
  const int32_t* row_map;  // optional: rows with row_map < 0 are padding (not stored)
  int a_box_rows;          // 128, or 32 with a row_map: only the 32-row boxes that hold live rows are loaded
  int64_t sb_e;  // expert stride of the B scales
};

template <int BN, typename OutT>
__global__ void __launch_bounds__(256, 1)
fp8_groupwise_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, OutT* __restrict__ C,
                     const GwParams p) {
  const Geo G = Geo::make(BN, kFp8);
  const int kStages = G.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G.bar_offset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 4);
    }
    ptx::fence_mbar_init();
  }
  constexpr uint32_t kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const bool grouped = p.tile_expert != nullptr;
  if (warp != 0 || grouped) ptx::grid_dep_wait();
  ptx::grid_dep_launch();
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = grouped ? (p.meta ? p.meta[0] : (p.M + BM - 1) / BM) : (p.M + BM - 1) / BM;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + BKB - 1) / BKB;
  // dense: m fastest (neighbouring CTAs share the weight tile); grouped: n fastest (they share the expert's A tile)
  auto decode = [&](int t, int& tm, int& tn, int& e) {
    if (grouped) {
      tn = t % tiles_n;
      tm = t / tiles_n;
      e = p.tile_expert[tm];
    } else {
      tm = t % tiles_m;
      tn = t / tiles_m;
      e = 0;
    }
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      bool first = !grouped;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int tm, tn, e;
        decode(t, tm, tn, e);
        if (e < 0) continue;
        const int brow = e * p.N + tn * BN;
        const int abr = p.a_box_rows;
        int nbox = 1;
        if (abr != BM) {
          for (int bx = 1; bx < BM / 32; ++bx)
            if (tm * BM + bx * 32 < p.M && p.row_map[tm * BM + bx * 32] >= 0) nbox = bx + 1;
        }
        const uint32_t a_tx = uint32_t(nbox * abr) * BKB;
        int kb_start = 0;
        if (first) {  // weights before griddepcontrol.wait
          first = false;
          const int npre = num_kb < kStages ? num_kb : kStages;
          for (int i = 0; i < npre; ++i) {
            ptx::mbar_arrive_expect_tx(&full_bar[i], a_tx + G.b_bytes);
            ptx::tma_load_2d(smem + i * G.stage_bytes + G.a_bytes, &tmB, &full_bar[i], i * BKB, brow, ptx::kEvictFirst);
          }
          ptx::grid_dep_wait();
          for (int i = 0; i < npre; ++i)
            for (int bx = 0; bx < nbox; ++bx)
              ptx::tma_load_2d(smem + i * G.stage_bytes + bx * abr * BKB, &tmA, &full_bar[i], i * BKB, tm * BM + bx * abr,
                               ptx::kEvictNormal);
          stage = npre == kStages ? 0 : npre;
          phase = npre == kStages ? 1 : 0;
          kb_start = npre;
        }
        for (int kb = kb_start; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * G.stage_bytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], a_tx + G.b_bytes);
          for (int bx = 0; bx < nbox; ++bx)
            ptx::tma_load_2d(sa + bx * abr * BKB, &tmA, &full_bar[stage], kb * BKB, tm * BM + bx * abr, ptx::kEvictNormal);
          ptx::tma_load_2d(sa + G.a_bytes, &tmB, &full_bar[stage], kb * BKB, brow, ptx::kEvictFirst);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    int stage = 0, buf = 0;
    uint32_t phase = 0, bphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      if (grouped && p.tile_expert[t / tiles_n] < 0) continue;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&tmem_empty[buf], bphase ^ 1);
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + stage * G.stage_bytes);
          const uint64_t da = ptx::make_smem_desc(sa, 16, 1024, ptx::kSwz128);
          const uint64_t db = ptx::make_smem_desc(sa + G.a_bytes, 16, 1024, ptx::kSwz128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f8f6f4_ss<1>(tmem_base + buf * BN, ptx::desc_advance(da, k * 32), ptx::desc_advance(db, k * 32), p.idesc,
                                  k > 0 ? 1u : 0u);
          ptx::mma_commit(&empty_bar[stage]);
          ptx::mma_commit(&tmem_full[buf]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
        buf ^= 1;
        if (buf == 0) bphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    int buf = 0;
    uint32_t bphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tm, tn, e;
      decode(t, tm, tn, e);
      if (e < 0) continue;
      const int row = tm * BM + q * 32 + lane;
      const int rowc = row < p.M ? row : p.M - 1;
      const int nblk = (tn * BN) / 128;
      const float* sbp = p.sb + e * p.sb_e + nblk * p.sb_n;
      float acc[BN];
#pragma unroll
      for (int c = 0; c < BN; ++c) acc[c] = 0.f;
      float sc_next = p.sa[rowc * p.sa_row] * sbp[0];
      for (int kb = 0; kb < num_kb; ++kb) {
        const float sc = sc_next;
        if (kb + 1 < num_kb) sc_next = p.sa[rowc * p.sa_row + (kb + 1) * p.sa_k] * sbp[(kb + 1) * p.sb_k];
        ptx::mbar_wait(&tmem_full[buf], bphase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + buf * BN + (uint32_t(q * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          ptx::tmem_ld_x32(taddr + c0, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c0 + j] = fmaf(__uint_as_float(v[j]), sc, acc[c0 + j]);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
        buf ^= 1;
        if (buf == 0) bphase ^= 1;
      }
      if (row < p.M && (!p.row_map || p.row_map[row] >= 0)) {
        OutT* crow = C + int64_t(row) * p.ldc + tn * BN;
        constexpr int VN = 16 / sizeof(OutT);
        const bool vec_ok = (p.ldc % VN == 0);
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += VN) {
          if (tn * BN + c0 + VN <= p.N && vec_ok) {
            Vec16<OutT> o;
#pragma unroll
            for (int e = 0; e < VN; ++e) o.v[e] = from_f32<OutT>(acc[c0 + e]);
            st16(crow + c0, o);
          } else {
#pragma unroll
            for (int e = 0; e < VN; ++e)
              if (tn * BN + c0 + e < p.N) crow[c0 + e] = from_f32<OutT>(acc[c0 + e]);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
  }
}

template <int BN, typename OutT>
int launch_gw(const CUtensorMap& tmA, const CUtensorMap& tmB, void* C, const GwParams& p, int grid, bool pdl, cudaStream_t stream) {
  static bool set = false;
  if (!set) {
    FIB_CUDA_CHECK(cudaFuncSetAttribute(fp8_groupwise_kernel<BN, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    set = true;
  }
  const Geo G = Geo::make(BN, kFp8);
  LaunchCfg lc(dim3(grid), dim3(256), G.total, stream, pdl);
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, fp8_groupwise_kernel<BN, OutT>, tmA, tmB, (OutT*)C, p));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variant for large problems.  A 2-CTA cluster owns a 256 x BN output tile: CTA r holds rows
// [r*128, r*128+128) of A (and of the accumulator, in its own TMEM) and HALF of the B tile (BN/2 rows); the leader's
// single MMA thread issues tcgen05.mma.cta_group::2 with M = 256, which reads A / B from both CTAs' shared memory.
// Per-SM operand traffic drops from (128 + BN) to (128 + BN/2) rows per slab, which is what lets the fp4 pipe run:
// at 1-CTA 128x256 tiles the kernel was latency x bandwidth bound (3 stages of 56 KB = 768 clk of look-ahead, tensor
// pipe 32 % busy in ncu); here a stage is 34 KB, 6 stages deep, with double-buffered 192-column accumulators.
// Synchronisation: both producers credit their TMA bytes to the LEADER's full barrier (cta_group::2 TMA with the peer
// bit cleared), tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs, the peer's epilogue warps
// arrive remotely on the leader's tmem_empty barrier.  Scale factors ride on 2-D tensor maps over the 512-byte blocks.
// ---------------------------------------------------------------------------------------------------------------
struct Geo2 {
  int stages, stage_bytes, a_bytes, b_bytes, sfa_bytes, sfb_bytes, bar_offset, total, nchunk, rb, sfa_col, sfb_col;
  __host__ __device__ static Geo2 make(int BN, int kind) {
    Geo2 g;
    g.nchunk = kind == kNvFp4 ? 4 : (kind == kMxFp4 ? 2 : (kind == kMxFp8 ? 1 : 0));
    g.rb = (BN % 128 == 0) ? BN / 128 : (BN + 64 + 127) / 128;
    g.a_bytes = BM * BKB;
    g.b_bytes = (BN / 2) * BKB;
    g.sfa_bytes = g.nchunk * 512;
    g.sfb_bytes = g.rb * g.nchunk * 512;
    g.stage_bytes = (g.a_bytes + g.b_bytes + g.sfa_bytes + g.sfb_bytes + 1023) / 1024 * 1024;
    int st = (216 * 1024) / g.stage_bytes;
    g.stages = st > 8 ? 8 : st;
    g.bar_offset = g.stages * g.stage_bytes;
    g.total = g.bar_offset + 320 + 1024;
    g.sfa_col = 2 * BN;
    g.sfb_col = g.sfa_col + g.nchunk * 4;
    return g;
  }
};

struct Params2 {
  const float* alpha_a;
  const float* alpha_b;
  int64_t c_batch_stride, ldc;
  int M, N, Kb, batch, BN;
  int sf_k_tiles, sfa_row_tiles, sfb_row_tiles;  // scale tensors: [batch][row tiles][sf_k_tiles] blocks of 512 B
  uint32_t idesc;                                // M = 256
};

template <int KIND, typename OutT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
bs_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmSFA, const __grid_constant__ CUtensorMap tmSFB, OutT* __restrict__ C,
                const Params2 p) {
  const Geo2 G = Geo2::make(p.BN, KIND);
  const int BN = p.BN;
  // fp4 slabs need 12 tcgen05.cp: a dedicated warp issues them one slab ahead.  mxfp8 (3 copies) keeps them inline:
  // the extra hand-shake costs more than it hides (measured 2.45 -> 2.09 PFLOP/s with the copy warp).
  constexpr bool kCopyWarp = (KIND == kNvFp4 || KIND == kMxFp4);
  const int kStages = G.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G.bar_offset);  // used on the leader only
  uint64_t* empty_bar = full_bar + kStages;                              // both CTAs (multicast commit)
  uint64_t* tmem_full = empty_bar + kStages;                             // both CTAs (multicast commit)
  uint64_t* tmem_empty = tmem_full + 2;                                  // leader only, 8 arrivals
  uint64_t* sf_ready = tmem_empty + 2;                                   // leader: scale copies of a slab landed in TMEM
  uint64_t* sf_free = sf_ready + 2;                                      // leader: MMAs that read a scale buffer retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sf_free + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = int(ptx::cluster_ctarank());
  const bool leader = crank == 0;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if constexpr (kind_scaled(KIND)) {
      ptx::prefetch_tmap(&tmSFA);
      ptx::prefetch_tmap(&tmSFB);
    }
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 8);
      ptx::mbar_init(&sf_ready[i], 1);
      ptx::mbar_init(&sf_free[i], 1);
    }
    ptx::fence_mbar_init();
  }
  constexpr uint32_t kTmemCols = 512;
  if (warp == 2) {
    ptx::tmem_alloc<2>(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish<2>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // both CTAs' barriers and TMEM exist before any cross-CTA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  ptx::grid_dep_wait();
  ptx::grid_dep_launch();

  const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM), tiles_n = (p.N + BN - 1) / BN;
  const int tiles_per_batch = tiles_m * tiles_n;
  const int num_tiles = tiles_per_batch * p.batch;
  const int num_kb = (p.Kb + BKB - 1) / BKB;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  // grouped rasterisation: the pairs that run concurrently cover a compact (8 M-tiles x n) block of the output, so the A rows
  // they stream stay L2 resident across the N sweep (walking all of M for one N column re-reads A from DRAM per column)
  auto tile_coords = [&](int r, int& tm, int& tn) {
    constexpr int kGroup = 8;
    const int per_group = kGroup * tiles_n;
    const int g = r / per_group;
    const int first = g * kGroup;
    const int rows = min(kGroup, tiles_m - first);
    const int in = r - g * per_group;
    tm = first + in % rows;
    tn = in / rows;
  };
  const uint32_t stage_tx = 2u * uint32_t(G.a_bytes + G.b_bytes + G.sfa_bytes + G.sfb_bytes);

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        const int b = t / tiles_per_batch, r = t % tiles_per_batch;
        int tm, tn;
        tile_coords(r, tm, tn);
        const int n0 = tn * BN;
        const int row0 = tm * 2 * BM + crank * BM;
        const int sfa_row = (tm * 2 + crank);
        const int rb0 = n0 / 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * G.stage_bytes;
          uint8_t* sb = sa + G.a_bytes;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
          ptx::tma2_load_3d(sa, &tmA, &full_bar[stage], kb * BKB, row0, b, ptx::kEvictNormal);
          ptx::tma2_load_3d(sb, &tmB, &full_bar[stage], kb * BKB, n0 + crank * (BN / 2), b, ptx::kEvictNormal);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 2) {
    // ---- second producer warp: scale-factor tiles (UTMALDG issue costs ~100 clk each; with A, B, SFA, SFB on one warp
    //      an nvfp4 slab (393 clk of MMA) was issue-bound).  Same full barrier; warp 0 arms the byte count.
    if constexpr (kind_scaled(KIND)) {
      if (ptx::elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = pair; t < num_tiles; t += num_pairs) {
          const int b = t / tiles_per_batch, r = t % tiles_per_batch;
          int tm, tn;
          tile_coords(r, tm, tn);
          const int sfa_row = tm * 2 + crank;
          const int rb0 = (tn * BN) / 128;
          for (int kb = 0; kb < num_kb; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* ssfa = smem + stage * G.stage_bytes + G.a_bytes + G.b_bytes;
            uint8_t* ssfb = ssfa + G.sfa_bytes;
            // scale tensors are [batch * row_tiles][sf_k_tiles][128 x u32]; out-of-range blocks are zero-filled by TMA
            ptx::tma2_load_3d(ssfa, &tmSFA, &full_bar[stage], 0, kb * G.nchunk, b * p.sfa_row_tiles + sfa_row, ptx::kEvictNormal);
            ptx::tma2_load_3d(ssfb, &tmSFB, &full_bar[stage], 0, kb * G.nchunk, b * p.sfb_row_tiles + rb0, ptx::kEvictNormal);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int sfb_i = 0;              // scale buffer (ping-pong), filled by the copy warp
    uint32_t sf_phase = 0;
    const uint32_t sf_stride = uint32_t(G.nchunk * 4 * (1 + G.rb));
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int r = t % tiles_per_batch;
      int tm_, tn_;
      tile_coords(r, tm_, tn_);
      const int n0 = tn_ * BN;
      const uint32_t sfb_off = uint32_t((n0 % 128) / 32);
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        if constexpr (kCopyWarp) ptx::mbar_wait(&sf_ready[sfb_i], sf_phase);
        ptx::tc_fence_after();
        const uint32_t sfa_t = tmem_base + G.sfa_col + sfb_i * sf_stride, sfb_t = sfa_t + G.nchunk * 4;
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + stage * G.stage_bytes);
          const uint32_t sb = sa + G.a_bytes;
          const uint64_t da = ptx::make_smem_desc(sa, 16, 1024, ptx::kSwz128);
          const uint64_t db = ptx::make_smem_desc(sb, 16, 1024, ptx::kSwz128);
          if constexpr (kind_scaled(KIND) && !kCopyWarp) {
            const uint32_t ssfa = sb + G.b_bytes, ssfb = ssfa + G.sfa_bytes;
            for (int c = 0; c < G.nchunk; ++c) {
              ptx::tmem_cp2_32x128b_warpx4(sfa_t + c * 4, ptx::make_smem_desc(ssfa + c * 512, 0, 128, ptx::kSwzNone));
              for (int rr = 0; rr < G.rb; ++rr)
                ptx::tmem_cp2_32x128b_warpx4(sfb_t + (c * G.rb + rr) * 4,
                                             ptx::make_smem_desc(ssfb + (rr * G.nchunk + c) * 512, 0, 128, ptx::kSwzNone));
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
            const uint64_t dak = ptx::desc_advance(da, k * 32), dbk = ptx::desc_advance(db, k * 32);
            if constexpr (KIND == kFp8) {
              ptx::mma_f8f6f4_ss<2>(d_tmem, dak, dbk, p.idesc, accum);
            } else if constexpr (KIND == kDense16) {  // a 128-byte slab is 64 16-bit elements: 4 MMAs of K = 16, same 32-byte steps
              ptx::mma_f16_ss<2>(d_tmem, dak, dbk, p.idesc, accum);
            } else if constexpr (KIND == kMxFp8) {
              const uint32_t id = p.idesc | (uint32_t(k) << 4) | (uint32_t(k) << 29);
              ptx::mma2_mxf8f6f4_ss(d_tmem, dak, dbk, id, sfa_t, sfb_t + sfb_off, accum);
            } else if constexpr (KIND == kNvFp4) {
              ptx::mma2_mxf4nvf4_ss(d_tmem, dak, dbk, p.idesc, sfa_t + k * 4, sfb_t + k * G.rb * 4 + sfb_off, accum);
            } else {
              const uint32_t sid = uint32_t(k & 1) * 2;
              const uint32_t id = p.idesc | (sid << 4) | (sid << 29);
              ptx::mma2_mxf4_2x_ss(d_tmem, dak, dbk, id, sfa_t + (k >> 1) * 4, sfb_t + (k >> 1) * G.rb * 4 + sfb_off, accum);
            }
          }
          ptx::mma_commit_2cta(&empty_bar[stage], 3);                      // slot free in BOTH CTAs
          if constexpr (kCopyWarp) ptx::mma_commit_2cta(&sf_free[sfb_i], 1);  // scale buffer reusable (leader's barrier)
          if (kb == num_kb - 1) ptx::mma_commit_2cta(&tmem_full[acc], 3);  // accumulator ready in BOTH CTAs
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
        if constexpr (kCopyWarp) {
          sfb_i ^= 1;
          if (sfb_i == 0) sf_phase ^= 1;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp == 3 && leader) {
    // ---- scale-factor copy warp: smem -> TMEM (both CTAs) one slab ahead of the MMA issuer, so that the 12 tcgen05.cp
    //      of an nvfp4 slab overlap the previous slab's MMAs instead of serialising with them on one thread
    if constexpr (kCopyWarp) {
      int stage = 0, sfb_i = 0;
      uint32_t phase = 0, sf_phase = 0;
      const uint32_t sf_stride = uint32_t(G.nchunk * 4 * (1 + G.rb));
      for (int t = pair; t < num_tiles; t += num_pairs) {
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&sf_free[sfb_i], sf_phase ^ 1);
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t sb = ptx::smem_u32(smem + stage * G.stage_bytes) + G.a_bytes;
            const uint32_t ssfa = sb + G.b_bytes, ssfb = ssfa + G.sfa_bytes;
            const uint32_t sfa_t = tmem_base + G.sfa_col + sfb_i * sf_stride, sfb_t = sfa_t + G.nchunk * 4;
            for (int c = 0; c < G.nchunk; ++c) {
              ptx::tmem_cp2_32x128b_warpx4(sfa_t + c * 4, ptx::make_smem_desc(ssfa + c * 512, 0, 128, ptx::kSwzNone));
              for (int rr = 0; rr < G.rb; ++rr)
                ptx::tmem_cp2_32x128b_warpx4(sfb_t + (c * G.rb + rr) * 4,
                                             ptx::make_smem_desc(ssfb + (rr * G.nchunk + c) * 512, 0, 128, ptx::kSwzNone));
            }
            ptx::mma_commit_2cta(&sf_ready[sfb_i], 1);  // cta_group::2 copies -> cta_group::2 commit, leader's barrier only
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
          sfb_i ^= 1;
          if (sfb_i == 0) sf_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    float alpha = 1.f;
    if (p.alpha_a) alpha *= *p.alpha_a;
    if (p.alpha_b) alpha *= *p.alpha_b;
    const uint32_t leader_empty = ptx::mapa(ptx::smem_u32(&tmem_empty[0]), 0);
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int b = t / tiles_per_batch, r = t % tiles_per_batch;
      int tm, tn;
      tile_coords(r, tm, tn);
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = tm * 2 * BM + crank * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + acc * BN + (uint32_t(q * 32) << 16);
      OutT* crow = C + int64_t(b) * p.c_batch_stride + int64_t(row) * p.ldc;
      const bool vec_ok = (p.ldc % (16 / sizeof(OutT)) == 0);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_x32(taddr + c0, v);
        ptx::tmem_ld_wait();
        const int col0 = tn * BN + c0;
        if (row < p.M) {
          if (col0 + 32 <= p.N && vec_ok) {
            constexpr int VN = 16 / sizeof(OutT);
#pragma unroll
            for (int j = 0; j < 32; j += VN) {
              Vec16<OutT> o;
#pragma unroll
              for (int e = 0; e < VN; ++e) o.v[e] = from_f32<OutT>(__uint_as_float(v[j + e]) * alpha);
              st16(crow + col0 + j, o);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) crow[col0 + j] = from_f32<OutT>(__uint_as_float(v[j]) * alpha);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(leader_empty + acc * 8);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();  // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<2>(tmem_base, kTmemCols);
  }
}

template <int KIND, typename OutT>
int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmSFA, const CUtensorMap& tmSFB, void* C,
            const Params2& p, int grid, int smem, bool pdl, cudaStream_t stream) {
  static bool set = false;
  if (!set) {
    FIB_CUDA_CHECK(cudaFuncSetAttribute(bs_gemm2_kernel<KIND, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    set = true;
  }
  LaunchCfg lc(dim3(grid), dim3(256), smem, stream, pdl);  // cluster dims are compiled in (__cluster_dims__)
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, bs_gemm2_kernel<KIND, OutT>, tmA, tmB, tmSFA, tmSFB, (OutT*)C, p));
  return 0;
}

template <int KIND, typename OutT>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, void* C, const Params& p, int grid, int smem, bool pdl,
           cudaStream_t stream) {
  const int cluster = p.split;
  static bool set = false;
  if (!set) {
    FIB_CUDA_CHECK(cudaFuncSetAttribute(bs_gemm_kernel<KIND, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    set = true;
  }
  LaunchCfg lc(dim3(grid), dim3(256), smem, stream, pdl, cluster);
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, bs_gemm_kernel<KIND, OutT>, tmA, tmB, (OutT*)C, p));
  return 0;
}

}  // namespace

// A [batch, M, Kbytes] (row stride lda bytes), B [batch, N, Kbytes] (ldb bytes): raw bytes (fp8: 1 elem / byte, fp4:
// 2 elems / byte).  sfa / sfb: 128x4 swizzled scale tensors ([row_tiles, sf_k_tiles, 512] per batch) or null (kind 0).
// a_fmt / b_fmt (kind 0/1): 0 = e4m3, 1 = e5m2.  bn: forced N tile (0 = heuristic).
extern "C" int gemm_lowp_nt(void* A, void* B, void* C, void* sfa, void* sfb, void* alpha_a, void* alpha_b, int64_t batch,
                            int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                            int64_t a_batch_stride, int64_t b_batch_stride, int64_t c_batch_stride,
                            int64_t sfa_batch_stride, int64_t sfb_batch_stride, int64_t kind, int64_t a_fmt,
                            int64_t b_fmt, int64_t out_dtype, int64_t bn, void* tile_expert, void* meta, void* row_map, int64_t pdl,
                            int64_t stream_) {
  FIB_CHECK(kind >= 0 && kind <= 4, "gemm_lowp: kind must be 0..4");
  FIB_CHECK(out_dtype == kF16 || out_dtype == kBF16, "gemm_lowp: output must be f16/bf16");
  const bool fp4 = kind == kNvFp4 || kind == kMxFp4;
  FIB_CHECK(K % (fp4 ? 32 : (kind == kDense16 ? 8 : 16)) == 0, "gemm_lowp: K must be a multiple of 16 (fp8) / 32 (fp4) / 8 (16-bit)");
  FIB_CHECK(lda % 16 == 0 && ldb % 16 == 0, "gemm_lowp: row strides must be multiples of 16 bytes");
  if (M == 0 || N == 0 || batch == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int64_t Kb = fp4 ? K / 2 : (kind == kDense16 ? 2 * K : K);
  const int tiles_m = int((M + BM - 1) / BM);
  const int64_t eb = tile_expert ? 1 : batch;  // grouped mode: `batch` counts experts (B / SFB), A and C are one matrix
  int BN = (int)bn;
  if (BN == 0) {
    const int sms = num_sms();
    const int64_t want = (N * tiles_m * eb + sms - 1) / sms;  // columns per CTA for one full wave
    if (want >= 256) {
      BN = (kind == kFp8) ? 256 : 192;
      // prefer the width with the least tile-quantisation waste over full waves
      auto waves = [&](int w) {
        const int64_t tiles = ((N + w - 1) / w) * tiles_m * eb;
        return double((tiles + sms - 1) / sms) * w;
      };
      const int cands[3] = {256, 192, 128};
      double best = 1e30;
      for (int c : cands) {
        if (kind != kFp8 && 2 * c + Geo::make(c, (int)kind).nchunk * 4 * (1 + Geo::make(c, (int)kind).rb) > 512 && c != 256) continue;
        const double cost = waves(c) * ((c == 256 && kind != kFp8) ? 1.15 : 1.0) * (c == 128 ? 1.1 : 1.0);
        if (cost < best) {
          best = cost;
          BN = c;
        }
      }
    } else {
      // fp8 per-tensor: any multiple of 16; block-scaled: multiples of 64 (scale columns are consumed in 64-row units)
      const int q = kind == kFp8 ? 32 : 64;
      BN = int((want + q - 1) / q * q);
      if (BN < q) BN = q;
    }
  }
  // ---- CTA-pair (cta_group::2) path for large problems ----
  {
    const char* env2 = getenv("FIB200_LOWP_2CTA");
    const int BN2 = bn ? (int)bn : ((kind == kFp8 || kind == kDense16) ? 256 : 192);
    const int64_t tiles2 = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN2 - 1) / BN2) * batch;
    bool use2 = !tile_expert && M >= 512 && tiles2 >= num_sms() / 2 && K >= 512;
    if (env2) use2 = atoi(env2) != 0 && !tile_expert;
    if (kind == kDense16 && !env2) use2 = !tile_expert;
    if (use2 && BN2 % 64 == 0 && BN2 >= 64 && BN2 <= 256 && (!kind_scaled((int)kind) || 2 * BN2 + 2 * Geo2::make(BN2, (int)kind).nchunk * 4 * (1 + Geo2::make(BN2, (int)kind).rb) <= 512)) {
      const Geo2 G2 = Geo2::make(BN2, (int)kind);
      CUtensorMap tmA, tmB, tmSFA, tmSFB;
      {
        uint64_t dims[3] = {(uint64_t)Kb, (uint64_t)M, (uint64_t)batch};
        uint64_t str[2] = {(uint64_t)lda, (uint64_t)a_batch_stride};
        uint32_t box[3] = {BKB, BM, 1};
        if (make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
      }
      {
        uint64_t dims[3] = {(uint64_t)Kb, (uint64_t)N, (uint64_t)batch};
        uint64_t str[2] = {(uint64_t)ldb, (uint64_t)b_batch_stride};
        uint32_t box[3] = {BKB, (uint32_t)(BN2 / 2), 1};
        if (make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, B, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
      }
      Params2 p2;
      const int vec2 = kind == kNvFp4 ? 16 : 32;
      p2.sf_k_tiles = !kind_scaled((int)kind) ? 0 : int(((K + vec2 - 1) / vec2 + 3) / 4);
      p2.sfa_row_tiles = int((M + 127) / 128);
      p2.sfb_row_tiles = int((N + 127) / 128);
      if (kind_scaled((int)kind)) {
        FIB_CHECK(sfa && sfb, "gemm_lowp: block-scaled kinds need scale tensors");
        FIB_CHECK(sfa_batch_stride == int64_t(p2.sfa_row_tiles) * p2.sf_k_tiles * 512 || batch == 1, "gemm_lowp: SFA must be contiguous per batch");
        FIB_CHECK(sfb_batch_stride == int64_t(p2.sfb_row_tiles) * p2.sf_k_tiles * 512 || batch == 1, "gemm_lowp: SFB must be contiguous per batch");
        // [rows = batch * row_tiles][sf_k_tiles][128 x u32]: one TMA box = (all scale blocks of a K slab) x (row tiles of the tile)
        uint64_t dimsa[3] = {128, (uint64_t)p2.sf_k_tiles, (uint64_t)(batch * int64_t(p2.sfa_row_tiles))};
        uint64_t dimsb[3] = {128, (uint64_t)p2.sf_k_tiles, (uint64_t)(batch * int64_t(p2.sfb_row_tiles))};
        uint64_t str[2] = {512, (uint64_t)p2.sf_k_tiles * 512};
        uint32_t boxa[3] = {128, (uint32_t)G2.nchunk, 1};
        uint32_t boxb[3] = {128, (uint32_t)G2.nchunk, (uint32_t)G2.rb};
        if (make_tmap(&tmSFA, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, sfa, dimsa, str, boxa, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
        if (make_tmap(&tmSFB, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, sfb, dimsb, str, boxb, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
      } else {
        tmSFA = tmA;
        tmSFB = tmB;
      }
      p2.alpha_a = (const float*)alpha_a;
      p2.alpha_b = (const float*)alpha_b;
      p2.c_batch_stride = c_batch_stride;
      p2.ldc = ldc;
      p2.M = (int)M; p2.N = (int)N; p2.Kb = (int)Kb; p2.batch = (int)batch; p2.BN = BN2;
      if (kind == kFp8) p2.idesc = ptx::make_idesc_f8((uint32_t)a_fmt, (uint32_t)b_fmt, 2 * BM, BN2, 0, 0);
      else if (kind == kDense16) p2.idesc = ptx::make_idesc_f16(a_fmt == 0 ? ptx::kFmtF16 : ptx::kFmtBF16, 2 * BM, BN2, 0, 0);
      else if (kind == kMxFp8) p2.idesc = ptx::make_idesc_blockscaled((uint32_t)a_fmt, (uint32_t)b_fmt, 2 * BM, BN2, 1, 0, 0);
      else p2.idesc = ptx::make_idesc_blockscaled(1, 1, 2 * BM, BN2, kind == kMxFp4 ? 1 : 0, 0, 0);
      const int pairs = (int)(tiles2 < num_sms() / 2 ? tiles2 : num_sms() / 2);
      const bool f16o = out_dtype == kF16;
      switch (kind) {
        case kDense16:
          return f16o ? launch2<kDense16, __half>(tmA, tmB, tmSFA, tmSFB, C, p2, 2 * pairs, G2.total, pdl != 0, stream)
                      : launch2<kDense16, __nv_bfloat16>(tmA, tmB, tmSFA, tmSFB, C, p2, 2 * pairs, G2.total, pdl != 0, stream);
        case kFp8:
          return f16o ? launch2<kFp8, __half>(tmA, tmB, tmSFA, tmSFB, C, p2, 2 * pairs, G2.total, pdl != 0, stream)
                      : launch2<kFp8, __nv_bfloat16>(tmA, tmB, tmSFA, tmSFB, C, p2, 2 * pairs, G2.total, pdl != 0, stream);
        case kMxFp8:
          return f16o ? launch2<kMxFp8, __half>(tmA, tmB, tmSFA, tmSFB, C, p2, 2 * pairs, G2.total, pdl != 0, stream)
                      : launch2<kMxFp8, __nv_bfloat16>(tmA, tmB, tmSFA, tmSFB, C, p2, 2 * pairs, G2.total, pdl != 0, stream);
        case kNvFp4:
          return f16o ? launch2<kNvFp4, __half>(tmA, tmB, tmSFA, tmSFB, C, p2, 2 * pairs, G2.total, pdl != 0, stream)
                      : launch2<kNvFp4, __nv_bfloat16>(tmA, tmB, tmSFA, tmSFB, C, p2, 2 * pairs, G2.total, pdl != 0, stream);
        default:
          return f16o ? launch2<kMxFp4, __half>(tmA, tmB, tmSFA, tmSFB, C, p2, 2 * pairs, G2.total, pdl != 0, stream)
                      : launch2<kMxFp4, __nv_bfloat16>(tmA, tmB, tmSFA, tmSFB, C, p2, 2 * pairs, G2.total, pdl != 0, stream);
      }
    }
  }
  FIB_CHECK(kind != kDense16, "gemm_lowp: 16-bit operands run on the CTA-pair kernel only (N tile multiple of 64)");
  FIB_CHECK(BN % 32 == 0 && BN >= 32 && BN <= 256, "gemm_lowp: N tile must be a multiple of 32 in [32, 256]");
  FIB_CHECK(kind == kFp8 || BN % 64 == 0, "gemm_lowp: block-scaled N tile must be a multiple of 64");
  const Geo G = Geo::make(BN, (int)kind);
  FIB_CHECK(G.stages >= 2 && G.tmem_cols <= 512, "gemm_lowp: tile does not fit");

  const int a_box_rows = (tile_expert && row_map) ? 32 : BM;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[3] = {(uint64_t)Kb, (uint64_t)M, (uint64_t)(tile_expert ? 1 : batch)};
    uint64_t str[2] = {(uint64_t)lda, (uint64_t)(tile_expert ? lda * M : a_batch_stride)};
    uint32_t box[3] = {BKB, (uint32_t)a_box_rows, 1};
    if (make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)Kb, (uint64_t)N, (uint64_t)batch};
    uint64_t str[2] = {(uint64_t)ldb, (uint64_t)b_batch_stride};
    uint32_t box[3] = {BKB, (uint32_t)BN, 1};
    if (make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, B, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  Params p;
  p.sfa = (const uint8_t*)sfa;
  p.tile_expert = (const int32_t*)tile_expert;
  p.meta = (const int32_t*)meta;
  p.row_map = (const int32_t*)row_map;
  p.tab_tiles = tile_expert ? (tiles_m < 1024 ? tiles_m : 1024) : 0;
  p.tab_experts = tile_expert ? (int)(batch < 1024 ? batch : 1024) : 0;
  p.a_box_rows = a_box_rows;
  const int tab_bytes = (2 * p.tab_tiles + p.tab_experts) * 4;
  p.sfb = (const uint8_t*)sfb;
  p.alpha_a = (const float*)alpha_a;
  p.alpha_b = (const float*)alpha_b;
  p.sfa_batch_stride = sfa_batch_stride;
  p.sfb_batch_stride = sfb_batch_stride;
  p.c_batch_stride = c_batch_stride;
  p.ldc = ldc;
  p.M = (int)M;
  p.N = (int)N;
  p.Kb = (int)Kb;
  p.batch = (int)batch;
  p.BN = BN;
  const int vec = kind == kNvFp4 ? 16 : 32;
  p.sf_k_tiles = kind == kFp8 ? 0 : int(((K + vec - 1) / vec + 3) / 4);
  p.sfb_row_tiles = int((N + 127) / 128);
  if (kind != kFp8) FIB_CHECK(sfa && sfb, "gemm_lowp: block-scaled kinds need scale tensors");
  if (kind == kFp8) {
    p.idesc = ptx::make_idesc_f8((uint32_t)a_fmt, (uint32_t)b_fmt, BM, BN, 0, 0);
  } else if (kind == kMxFp8) {
    p.idesc = ptx::make_idesc_blockscaled((uint32_t)a_fmt, (uint32_t)b_fmt, BM, BN, 1, 0, 0);
  } else {
    p.idesc = ptx::make_idesc_blockscaled(1, 1, BM, BN, kind == kMxFp4 ? 1 : 0, 0, 0);
  }
  const int64_t tiles = int64_t(tiles_m) * ((N + BN - 1) / BN) * eb;
  // cluster split-K for latency-bound shapes: few tiles, long K -> two CTAs per tile, DSMEM reduction
  const int64_t num_kb = (Kb + BKB - 1) / BKB;
  const char* env_split = getenv("FIB200_LOWP_SPLIT");
  p.split = env_split ? atoi(env_split) : ((tiles * 2 <= num_sms() && num_kb >= 8) ? 2 : 1);  // m=512,n=1024,k=7168 nvfp4: 18.0 -> 15.2 us
  if (tile_expert) p.split = 1;
  if (p.split != 2 || tiles * 2 > num_sms() || num_kb < 2 || BM * BN * 4 > G.stages * G.stage_bytes) p.split = 1;
  const int grid = p.split > 1 ? (int)(tiles * p.split) : (int)(tiles < num_sms() ? tiles : num_sms());
  const bool f16 = out_dtype == kF16;
  switch (kind) {
    case kFp8:
      return f16 ? launch<kFp8, __half>(tmA, tmB, C, p, grid, G.total + tab_bytes, pdl != 0, stream)
                 : launch<kFp8, __nv_bfloat16>(tmA, tmB, C, p, grid, G.total + tab_bytes, pdl != 0, stream);
    case kMxFp8:
      return f16 ? launch<kMxFp8, __half>(tmA, tmB, C, p, grid, G.total + tab_bytes, pdl != 0, stream)
                 : launch<kMxFp8, __nv_bfloat16>(tmA, tmB, C, p, grid, G.total + tab_bytes, pdl != 0, stream);
    case kNvFp4:
      return f16 ? launch<kNvFp4, __half>(tmA, tmB, C, p, grid, G.total + tab_bytes, pdl != 0, stream)
                 : launch<kNvFp4, __nv_bfloat16>(tmA, tmB, C, p, grid, G.total + tab_bytes, pdl != 0, stream);
    default:
      return f16 ? launch<kMxFp4, __half>(tmA, tmB, C, p, grid, G.total + tab_bytes, pdl != 0, stream)
                 : launch<kMxFp4, __nv_bfloat16>(tmA, tmB, C, p, grid, G.total + tab_bytes, pdl != 0, stream);
  }
}

// A [M, K] fp8 (lda bytes), B [N, K] fp8; sa fp32 with element strides (sa_row, sa_k) over [M, K/128];
// sb fp32 with element strides (sb_n, sb_k) over [N/128, K/128].
extern "C" int gemm_fp8_groupwise_nt(void* A, void* B, void* C, void* sa, void* sb, int64_t M, int64_t N, int64_t K, int64_t lda,
                                     int64_t ldb, int64_t ldc, int64_t sa_row, int64_t sa_k, int64_t sb_n, int64_t sb_k,
                                     int64_t a_fmt, int64_t b_fmt, int64_t out_dtype, int64_t bn, void* tile_expert, void* meta,
                                     int64_t num_experts, int64_t sb_e, void* row_map, int64_t pdl, int64_t stream_) {
  // tile_expert != null: m-grouped contiguous mode, B is [num_experts * N, K], sb has an expert stride sb_e, M = padded rows
  FIB_CHECK(out_dtype == kF16 || out_dtype == kBF16, "gemm_fp8_groupwise: output must be f16/bf16");
  FIB_CHECK(K % 128 == 0, "gemm_fp8_groupwise: K must be a multiple of 128");
  FIB_CHECK(lda % 16 == 0 && ldb % 16 == 0, "gemm_fp8_groupwise: row strides must be multiples of 16 bytes");
  if (M == 0 || N == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int tiles_m = int((M + BM - 1) / BM);
  int BN = (int)bn;
  if (BN == 0) {
    const int64_t want = (N * tiles_m + num_sms() - 1) / num_sms();
    BN = want <= 32 ? 32 : (want <= 64 ? 64 : 128);
  }
  FIB_CHECK(BN == 32 || BN == 64 || BN == 128, "gemm_fp8_groupwise: N tile must be 32 / 64 / 128");
  if (tile_expert) {
    while (BN > 32 && N % BN) BN >>= 1;
    FIB_CHECK(N % BN == 0 && num_experts >= 1, "gemm_fp8_groupwise (grouped): N must be a multiple of 32");
  }
  const int64_t b_rows = tile_expert ? num_experts * N : N;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda};
    uint32_t box[2] = {BKB, (uint32_t)((tile_expert && row_map) ? 32 : BM)};
    if (make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)b_rows};
    uint64_t str[1] = {(uint64_t)ldb};
    uint32_t box[2] = {BKB, (uint32_t)BN};
    if (make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, B, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  GwParams p;
  p.tile_expert = (const int32_t*)tile_expert;
  p.meta = (const int32_t*)meta;
  p.row_map = (const int32_t*)row_map;
  p.a_box_rows = (tile_expert && row_map) ? 32 : BM;
  p.sb_e = sb_e;
  p.sa = (const float*)sa;
  p.sb = (const float*)sb;
  p.sa_row = sa_row; p.sa_k = sa_k; p.sb_n = sb_n; p.sb_k = sb_k; p.ldc = ldc;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.idesc = ptx::make_idesc_f8((uint32_t)a_fmt, (uint32_t)b_fmt, BM, BN, 0, 0);
  const int64_t tiles = int64_t(tiles_m) * ((N + BN - 1) / BN);
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  const bool f16 = out_dtype == kF16;
  if (BN == 32) return f16 ? launch_gw<32, __half>(tmA, tmB, C, p, grid, pdl != 0, stream) : launch_gw<32, __nv_bfloat16>(tmA, tmB, C, p, grid, pdl != 0, stream);
  if (BN == 64) return f16 ? launch_gw<64, __half>(tmA, tmB, C, p, grid, pdl != 0, stream) : launch_gw<64, __nv_bfloat16>(tmA, tmB, C, p, grid, pdl != 0, stream);
  return f16 ? launch_gw<128, __half>(tmA, tmB, C, p, grid, pdl != 0, stream) : launch_gw<128, __nv_bfloat16>(tmA, tmB, C, p, grid, pdl != 0, stream);
}
