// Grouped (per-expert) bf16/fp16 GEMM for sm_100a:  C[rows of expert e, :] = A[rows of expert e, :] * W[e]^T
//
// Parity: reference SegmentGEMMWrapper (flashinfer/gemm/gemm_base.py:1736-1992), grouped_mm_bf16
// (flashinfer/grouped_mm/core.py), the m-grouped contiguous DeepGEMM layout (flashinfer/deep_gemm.py:1425-1585)
// and the MoE grouped GEMMs of M1-M3.
//
// Layout contract ("m-grouped contiguous"): rows of A/C are grouped by expert and every group starts at a
// multiple of 128 rows, so a 128-row M tile never straddles experts; `tile_expert[m_tile]` names the expert
// (-1 = unused tile) and `meta[0]` holds the number of live M tiles ON THE DEVICE (CUDA-graph friendly: the
// grid is fixed, CTAs read the count).  W is [E, N, K] (K-major) and is addressed through a 3-D TMA map,
// so the expert id is just a TMA coordinate.  Same warp-specialised tcgen05 pipeline as the dense GEMM.
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

constexpr int BM = 128, BK = 64;

struct GSmem {
  int stages, stage_bytes, a_bytes, bar_offset, total;
  __host__ __device__ static GSmem make(int BN) {
    GSmem g;
    g.a_bytes = BM * BK * 2;
    g.stage_bytes = g.a_bytes + BN * BK * 2;
    int st = (216 * 1024) / g.stage_bytes;
    g.stages = st > 8 ? 8 : st;
    g.bar_offset = g.stages * g.stage_bytes;
    g.total = g.bar_offset + 320 + 1024;
    return g;
  }
};

template <typename OutT>
__global__ void __launch_bounds__(256, 1)
grouped_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    OutT* __restrict__ C, const int32_t* __restrict__ tile_expert, const int32_t* __restrict__ meta,
                    int max_m_tiles, int N, int K, int64_t ldc, int BN, uint32_t idesc, int tab_tiles,
                    const int32_t* __restrict__ row_map) {
  const GSmem S = GSmem::make(BN);
  const int kStages = S.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S.bar_offset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
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
  uint32_t tmem_cols = 32;
  while (tmem_cols < uint32_t(2 * BN)) tmem_cols <<= 1;
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_ptr, tmem_cols);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  ptx::grid_dep_wait();
  int num_m = meta ? meta[0] : max_m_tiles;
  if (num_m > max_m_tiles) num_m = max_m_tiles;
  // stage the tile -> expert table in shared memory (a global read per tile is an L2 round trip on every role's critical path)
  int32_t* s_expert = reinterpret_cast<int32_t*>(smem + S.bar_offset + 320);
  const int tab = num_m < tab_tiles ? num_m : tab_tiles;
  // live-row A loading (see gemm_blockscaled_sm100.cu): with a row_map the A tile is fetched in 32-row boxes and only the
  // boxes holding live rows are loaded; stale smem rows only feed accumulator rows that are never stored
  int32_t* s_nbox = s_expert + tab_tiles;
  const int abr = row_map ? 32 : BM;
  for (int i = threadIdx.x; i < tab; i += blockDim.x) {
    s_expert[i] = tile_expert[i];
    int nb = BM / 32;
    if (row_map) {
      nb = 1;
      for (int bx = 1; bx < BM / 32; ++bx)
        if (row_map[i * BM + bx * 32] >= 0) nb = bx + 1;
    }
    s_nbox[i] = nb;
  }
  __syncthreads();
  auto expert_of = [&](int tm) { return tm < tab ? s_expert[tm] : tile_expert[tm]; };
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int tm = t / tiles_n, tn = t % tiles_n;
        const int e = expert_of(tm);
        if (e < 0) continue;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S.stage_bytes;
          uint8_t* sb = sa + S.a_bytes;
          const int nbox = abr == BM ? 1 : (tm < tab ? s_nbox[tm] : BM / 32);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], S.stage_bytes - S.a_bytes + nbox * abr * BK * 2);
          for (int bx = 0; bx < nbox; ++bx)
            ptx::tma_load_2d(sa + bx * abr * BK * 2, &tmA, &full_bar[stage], kb * BK, tm * BM + bx * abr, ptx::kEvictFirst);
          ptx::tma_load_3d(sb, &tmW, &full_bar[stage], kb * BK, tn * BN, e, ptx::kEvictNormal);
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
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int tm = t / tiles_n;
      if (expert_of(tm) < 0) continue;
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + stage * S.stage_bytes);
          const uint32_t sb = sa + S.a_bytes;
          const uint64_t da = ptx::make_smem_desc(sa, 16, 1024, ptx::kSwz128);
          const uint64_t db = ptx::make_smem_desc(sb, 16, 1024, ptx::kSwz128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            ptx::mma_f16_ss<1>(d_tmem, ptx::desc_advance(da, k * 32), ptx::desc_advance(db, k * 32), idesc,
                               (kb > 0 || k > 0) ? 1u : 0u);
          ptx::mma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) ptx::mma_commit(&tmem_full[acc]);
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
    const int q = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int tm = t / tiles_n, tn = t % tiles_n;
      if (expert_of(tm) < 0) continue;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = tm * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + acc * BN + (uint32_t(q * 32) << 16);
      // MoE: padding rows of the tile-aligned layout (row_map < 0) are never read back -> do not spend HBM writes on them
      const bool live = !row_map || row_map[row] >= 0;
      const bool warp_live = __any_sync(0xffffffffu, live);
#pragma unroll 1
      for (int c0 = 0; c0 < BN && warp_live; c0 += 16) {
        uint32_t r[16];
        ptx::tmem_ld_x16(taddr + c0, r);
        ptx::tmem_ld_wait();
        const int col0 = tn * BN + c0;
        OutT* dst = C + int64_t(row) * ldc + col0;
        if (!live) continue;
        if (col0 + 16 <= N) {
          constexpr int VN = 16 / sizeof(OutT);
#pragma unroll
          for (int j = 0; j < 16; j += VN) {
            Vec16<OutT> o;
#pragma unroll
            for (int e2 = 0; e2 < VN; ++e2) o.v[e2] = from_f32<OutT>(__uint_as_float(r[j + e2]));
            st16(dst + j, o);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (col0 + j < N) dst[j] = from_f32<OutT>(__uint_as_float(r[j]));
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, tmem_cols);
  }
}

}  // namespace

// A [max_m_tiles*128, K] (lda), W [E, N, K] contiguous, C [max_m_tiles*128, N] (ldc).
extern "C" int grouped_gemm_nt(void* A, void* W, void* C, void* tile_expert, void* meta, int64_t max_m_tiles, int64_t N,
                               int64_t K, int64_t E, int64_t lda, int64_t ldc, void* row_map, int64_t dtype, int64_t pdl,
                               int64_t stream_) {
  FIB_CHECK(K % 8 == 0 && lda % 8 == 0 && ldc % 8 == 0 && N % 8 == 0, "grouped_gemm: K/N/lda/ldc must be multiples of 8");
  FIB_CHECK(dtype == kF16 || dtype == kBF16, "grouped_gemm: dtype must be f16/bf16");
  if (max_m_tiles == 0 || N == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const CUtensorMapDataType dt = dtype == kF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int BN = (N % 256 == 0 && max_m_tiles * (N / 256) >= 2 * num_sms()) ? 256 : 128;
  CUtensorMap tmA, tmW;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)(max_m_tiles * BM)};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {BK, (uint32_t)(row_map ? 32 : BM)};
    if (make_tmap(&tmA, dt, 2, A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)E};
    uint64_t str[2] = {(uint64_t)K * 2, (uint64_t)N * K * 2};
    uint32_t box[3] = {BK, (uint32_t)BN, 1};
    if (make_tmap(&tmW, dt, 3, W, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  const GSmem S = GSmem::make(BN);
  const uint32_t idesc = ptx::make_idesc_f16(dtype == kF16 ? ptx::kFmtF16 : ptx::kFmtBF16, BM, BN, 0, 0);
  const int64_t tiles = max_m_tiles * ((N + BN - 1) / BN);
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  const int tab_tiles = (int)(max_m_tiles < 2048 ? max_m_tiles : 2048);
  LaunchCfg lc(dim3(grid), dim3(256), S.total + tab_tiles * 8, stream, pdl != 0);
  if (dtype == kF16) {
    static bool set = false;
    if (!set) {
      FIB_CUDA_CHECK(cudaFuncSetAttribute(grouped_gemm_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set = true;
    }
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, grouped_gemm_kernel<__half>, tmA, tmW, (__half*)C,
                                      (const int32_t*)tile_expert, (const int32_t*)meta, (int)max_m_tiles, (int)N, (int)K,
                                      ldc, BN, idesc, tab_tiles, (const int32_t*)row_map));
  } else {
    static bool set = false;
    if (!set) {
      FIB_CUDA_CHECK(cudaFuncSetAttribute(grouped_gemm_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          227 * 1024));
      set = true;
    }
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, grouped_gemm_kernel<__nv_bfloat16>, tmA, tmW, (__nv_bfloat16*)C,
                                      (const int32_t*)tile_expert, (const int32_t*)meta, (int)max_m_tiles, (int)N, (int)K,
                                      ldc, BN, idesc, tab_tiles, (const int32_t*)row_map));
  }
  return 0;
}
