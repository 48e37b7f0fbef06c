// Fused all-gather + GEMM for sequence-parallel column-parallel linears:  out = all_gather_rows(x) * W^T
// ONE kernel per rank: extra "comm" warps push this rank's activation shard to every peer (multimem.st: one store,
// replicated by the NVSwitch) chunk by chunk while the tcgen05 pipeline already multiplies the chunks that arrived.
//
// Parity: reference flashinfer/comm/all_gather_matmul (all_gather_matmul.py:52-75: comm-stream broadcast with
// per-chunk signals + persistent matmul that spin-waits before loading a remote chunk; cuTile / Triton kernels).
// Here both halves live in one kernel, so there is no stream-scheduling hazard: comm warps never wait, the TMA
// producer waits (ld.acquire.sys + fence.proxy.async) only for the 128-row tile it is about to load, and the tile
// order starts with the local shard.
//
// Roles (384 threads): warp 0 TMA producer | warp 1 MMA | warp 2 TMEM alloc | warps 4-7 epilogue | warps 8-11 comm.
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kMaxRanks = 16;

struct GSmem {
  int stages, stage_bytes, a_bytes, bar_offset, total;
  __host__ __device__ static GSmem make(int BN) {
    GSmem g;
    g.a_bytes = BM * BK * 2;
    g.stage_bytes = g.a_bytes + ((BN * BK * 2 + 1023) / 1024) * 1024;
    int st = (200 * 1024) / g.stage_bytes;
    g.stages = st > 8 ? 8 : st;
    g.bar_offset = g.stages * g.stage_bytes;
    g.total = g.bar_offset + 320 + 1024;
    return g;
  }
};

struct AGP {
  uint8_t* peer_gath[kMaxRanks];   // symmetric gathered-activation buffer of every rank [world*Ml, K]
  uint32_t* peer_flags[kMaxRanks]; // symmetric per-row-tile arrival counters
  uint8_t* mc_gath;                // multicast aliases (or null)
  uint32_t* mc_flags;
  uint32_t* expect;                // local expected counters [num_sms][max_row_tiles] (one private copy per CTA)
  int max_row_tiles;
  const uint8_t* x_local;          // this rank's shard [Ml, K]
  int64_t x_stride_bytes;
  int rank, world, Ml, slices;     // rows per rank (multiple of 128), column slices per chunk
};

template <typename OutT>
__global__ void __launch_bounds__(384, 1)
ag_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, OutT* __restrict__ C,
               int M, int N, int K, int64_t ldc, int BN, uint32_t idesc, const AGP ag) {
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

  const int tiles_m = M / BM, tiles_n = (N + BN - 1) / BN;
  const int tiles_per_rank = ag.Ml / BM;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;
  ptx::grid_dep_wait();
  // tile order: n fastest inside a row tile, row tiles rotated so that the local shard comes first
  auto row_tile = [&](int t) { return ((t / tiles_n) + ag.rank * tiles_per_rank) % tiles_m; };

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage_i = 0;
      uint32_t phase = 0;
      int ready_tile = -1;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int tm = row_tile(t), tn = t % tiles_n;
        if (tm != ready_tile) {
          const uint32_t want = ag.expect[int64_t(blockIdx.x) * ag.max_row_tiles + tm] + uint32_t(ag.slices);
          while (int32_t(ptx::ld_acquire_sys(ag.peer_flags[ag.rank] + tm) - want) < 0) {
          }
          ptx::fence_proxy_async();  // peer (generic-proxy) writes -> our TMA (async-proxy) reads
          ready_tile = tm;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage_i], phase ^ 1);
          uint8_t* sa = smem + stage_i * S.stage_bytes;
          uint8_t* sb = sa + S.a_bytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage_i], S.a_bytes + BN * BK * 2);
          ptx::tma_load_2d(sa, &tmA, &full_bar[stage_i], kb * BK, tm * BM, ptx::kEvictNormal);
          ptx::tma_load_2d(sb, &tmW, &full_bar[stage_i], kb * BK, tn * BN, ptx::kEvictNormal);
          if (++stage_i == kStages) {
            stage_i = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    int stage_i = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage_i], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + stage_i * S.stage_bytes);
          const uint32_t sb = sa + S.a_bytes;
          const uint64_t da = ptx::make_smem_desc(sa, 16, 1024, ptx::kSwz128);
          const uint64_t db = ptx::make_smem_desc(sb, 16, 1024, ptx::kSwz128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            ptx::mma_f16_ss<1>(d_tmem, ptx::desc_advance(da, k * 32), ptx::desc_advance(db, k * 32), idesc,
                               (kb > 0 || k > 0) ? 1u : 0u);
          ptx::mma_commit(&empty_bar[stage_i]);
          if (kb == num_kb - 1) ptx::mma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage_i == kStages) {
          stage_i = 0;
          phase ^= 1;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4 && warp < 8) {
    const int q = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int tm = row_tile(t), tn = t % tiles_n;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int row = tm * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + acc * BN + (uint32_t(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t r[16];
        ptx::tmem_ld_x16(taddr + c0, r);
        ptx::tmem_ld_wait();
        const int col0 = tn * BN + c0;
        if (col0 < N) {
          OutT* dst = C + int64_t(row) * ldc + col0;
          constexpr int VN = 16 / sizeof(OutT);
#pragma unroll
          for (int j = 0; j < 16; j += VN) {
            Vec16<OutT> o;
#pragma unroll
            for (int e2 = 0; e2 < VN; ++e2) o.v[e2] = from_f32<OutT>(__uint_as_float(r[j + e2]));
            st16(dst + j, o);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    ptx::grid_dep_launch();
  } else if (warp >= 8) {
    // ---------------- comm warps: push (chunk, slice) items of the local shard to every rank
    const int tid = threadIdx.x - 256;
    const int64_t row_vecs = int64_t(K) * sizeof(OutT) / 16;
    const int64_t slice_vecs = (row_vecs + ag.slices - 1) / ag.slices;
    const int items = tiles_per_rank * ag.slices;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int c = it / ag.slices, s = it % ag.slices;
      const int64_t v0 = s * slice_vecs;
      const int64_t v1 = (v0 + slice_vecs < row_vecs) ? v0 + slice_vecs : row_vecs;
      const int64_t w = v1 - v0;
      for (int64_t i = tid; i < int64_t(BM) * w; i += 128) {
        const int64_t r = i / w, v = v0 + i % w;
        const int4 val = *reinterpret_cast<const int4*>(ag.x_local + (int64_t(c) * BM + r) * ag.x_stride_bytes + v * 16);
        const int64_t off = ((int64_t(ag.rank) * ag.Ml + int64_t(c) * BM + r) * row_vecs + v) * 16;
        if (ag.mc_gath) {
          ptx::multimem_st_v4(ag.mc_gath + off, val);
        } else {
          for (int p = 0; p < ag.world; ++p) *reinterpret_cast<int4*>(ag.peer_gath[p] + off) = val;
        }
      }
      __threadfence_system();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (tid == 0) {
        const int gt = ag.rank * tiles_per_rank + c;
        if (ag.mc_flags) {
          ptx::multimem_red_add_u32(ag.mc_flags + gt, 1u);
        } else {
          for (int p = 0; p < ag.world; ++p) ptx::red_add_release_sys(ag.peer_flags[p] + gt, 1u);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  // every CTA advances its private copy of the expected counters once per launch (the grid is always num_sms wide)
  for (int i = threadIdx.x; i < tiles_m; i += blockDim.x) ag.expect[int64_t(blockIdx.x) * ag.max_row_tiles + i] += uint32_t(ag.slices);
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, tmem_cols);
  }
}

}  // namespace

// x_local [Ml, K] (this rank's rows), W [N, K]; gathered buffer (symmetric) [world*Ml, K]; C [world*Ml, N] local.
extern "C" int allgather_gemm_nt(void* x_local, void* W, void* gathered, void* C, int64_t Ml, int64_t N, int64_t K,
                                 int64_t ldx, int64_t ldw, int64_t ldc, int64_t dtype, void* peer_gath_tab,
                                 void* peer_flags_tab, void* mc_gath, void* mc_flags, void* expect, int64_t max_row_tiles,
                                 int64_t rank, int64_t world, int64_t slices, int64_t bn, int64_t pdl, int64_t stream_) {
  FIB_CHECK(K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0 && N % 16 == 0, "allgather_gemm: K/ld* must be multiples of 8, N of 16");
  FIB_CHECK(Ml % BM == 0 && Ml > 0, "allgather_gemm: rows per rank must be a positive multiple of 128");
  FIB_CHECK(dtype == kF16 || dtype == kBF16, "allgather_gemm: dtype must be f16/bf16");
  FIB_CHECK(world >= 1 && world <= kMaxRanks && slices >= 1, "allgather_gemm: bad world / slices");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int64_t M = Ml * world;
  const CUtensorMapDataType dt = dtype == kF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  int BN = (int)bn;
  if (BN == 0) BN = (N % 256 == 0 && (M / BM) * (N / 256) >= num_sms()) ? 256 : 128;
  if (BN > N) BN = (int)N;
  FIB_CHECK(BN % 16 == 0 && BN >= 16 && BN <= 256, "allgather_gemm: bad N tile");
  CUtensorMap tmA, tmW;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {BK, BM};
    if (make_tmap(&tmA, dt, 2, gathered, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)BN};
    if (make_tmap(&tmW, dt, 2, W, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  AGP ag;
  for (int i = 0; i < world; ++i) {
    ag.peer_gath[i] = reinterpret_cast<uint8_t*>(((const int64_t*)peer_gath_tab)[i]);
    ag.peer_flags[i] = reinterpret_cast<uint32_t*>(((const int64_t*)peer_flags_tab)[i]);
  }
  ag.mc_gath = (uint8_t*)mc_gath;
  ag.mc_flags = (uint32_t*)mc_flags;
  ag.expect = (uint32_t*)expect;
  ag.max_row_tiles = (int)max_row_tiles;
  FIB_CHECK(M / BM <= max_row_tiles, "allgather_gemm: too many row tiles for the flag array");
  ag.x_local = (const uint8_t*)x_local;
  ag.x_stride_bytes = ldx * 2;
  ag.rank = (int)rank;
  ag.world = (int)world;
  ag.Ml = (int)Ml;
  ag.slices = (int)slices;
  const GSmem S = GSmem::make(BN);
  const uint32_t idesc = ptx::make_idesc_f16(dtype == kF16 ? ptx::kFmtF16 : ptx::kFmtBF16, BM, BN, 0, 0);
  const int grid = num_sms();  // fixed: every CTA keeps a private expectation table in step
  LaunchCfg lc(dim3(grid), dim3(384), S.total, stream, pdl != 0);
  if (dtype == kF16) {
    static bool set = false;
    if (!set) {
      FIB_CUDA_CHECK(cudaFuncSetAttribute(ag_gemm_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set = true;
    }
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, ag_gemm_kernel<__half>, tmA, tmW, (__half*)C, (int)M, (int)N, (int)K, ldc, BN, idesc, ag));
  } else {
    static bool set = false;
    if (!set) {
      FIB_CUDA_CHECK(cudaFuncSetAttribute(ag_gemm_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set = true;
    }
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, ag_gemm_kernel<__nv_bfloat16>, tmA, tmW, (__nv_bfloat16*)C, (int)M, (int)N, (int)K, ldc, BN, idesc, ag));
  }
  return 0;
}
