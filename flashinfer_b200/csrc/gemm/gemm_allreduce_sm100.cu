// Fused GEMM + all-reduce for tensor-parallel row-parallel linears:  out = sum_ranks( A_r[M,K_r] * W_r[N,K_r]^T )
// ONE kernel per rank: the tcgen05 GEMM and the NVLS reduction overlap tile by tile.
//
// Parity: reference flashinfer/cute_dsl/gemm_allreduce_two_shot.py (PersistentDenseGemmKernel(all_reduce="two_shot")
// :216-360: epilogue stores the tile to a symmetric C, multimem.red arrives on a per-tile flag, dedicated all-reduce
// warps wait for the flag and multimem.ld_reduce / multimem.st the tile).
//
// Roles (384 threads): warp 0 TMA producer | warp 1 MMA issuer | warp 2 TMEM allocator | warps 4-7 epilogue
// (TMEM -> bf16 -> symmetric staging buffer, then ONE multimem.red.release bumps the tile flag on every rank) |
// warps 8-11 all-reduce (spin on the local flag until all ranks have stored that tile, then pull the in-switch sum
// with multimem.ld_reduce and write the final tile; two-shot: only the owner rank reduces and multicast-stores).
// All ranks walk the tiles in the same order, so the waits are short and cannot deadlock (the GEMM side never waits
// on a peer).  Flags are monotonic counters with the expected value kept per tile in local memory (graph-replay safe).
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
    g.total = g.bar_offset + 320 + 512 /* per-row sum-of-squares scratch of the reduce-scatter mode */ + 1024;
    return g;
  }
};

struct ARP {
  uint8_t* peer_stage[kMaxRanks];  // symmetric staging C of every rank (P2P fallback)
  uint32_t* peer_flags[kMaxRanks]; // symmetric per-tile flags of every rank
  uint8_t* peer_out[kMaxRanks];    // symmetric out of every rank (two-shot P2P fallback)
  uint8_t* mc_stage;               // multicast alias of the staging C (or null)
  uint32_t* mc_flags;              // multicast alias of the flags (or null)
  uint8_t* mc_out;                 // multicast alias of the symmetric out (two-shot)
  uint32_t* expect;                // local: expected flag value per tile
  uint32_t* peer_done[kMaxRanks];  // two-shot end barrier slots [max_ctas][world]
  uint32_t* done_epoch;            // local [max_ctas]
  int rank, world, two_shot;      // two_shot: 0 one-shot all-reduce | 1 two-shot all-reduce | 2 reduce-scatter (row shards)
  int rows_per_rank;               // reduce-scatter: rank r owns rows [r * rows_per_rank, (r + 1) * rows_per_rank)
  int64_t ldo;                     // reduce-scatter: leading dimension of the local out / residual shard
  const void* residual;            // reduce-scatter: optional residual shard added to the reduced rows
  float* sumsq;                    // reduce-scatter: optional per-row sum of squares of (sum + residual) for the fused RMSNorm
};

// Reduce-scatter tail of one tile: only the rows this rank owns are pulled (in-switch sum), the optional residual shard is
// added, the result goes to the LOCAL out shard, and (optionally) the per-row sum of squares is accumulated for the RMSNorm
// that follows (rs_norm_kernel).  Ranks that own no row of the tile neither wait nor read.
template <typename OutT>
__device__ __forceinline__ void rs_reduce_tile(const ARP& ar, OutT* __restrict__ out, int tm, int tn, int M, int N, int BN,
                                               int64_t ldc, uint32_t want, int t, int tid, float* s_sq) {
  constexpr int VN = 16 / sizeof(OutT);  // s_sq: BM floats of dynamic smem (a static array would eat into the 227 KB opt-in)
  const int lo = max(tm * BM, ar.rank * ar.rows_per_rank);
  const int hi = min(min(tm * BM + BM, M), (ar.rank + 1) * ar.rows_per_rank);
  if (lo >= hi) return;  // uniform over the 128 all-reduce threads
  if (tid == 0) {
    while (int32_t(ptx::ld_acquire_sys(ar.peer_flags[ar.rank] + t) - want) < 0) {
    }
  }
  s_sq[tid] = 0.f;
  asm volatile("bar.sync 2, 128;" ::: "memory");
  const int vec_per_row = BN / VN;
  const int rows = hi - lo;
  constexpr int U = 8;  // in-switch reductions in flight per thread (a multimem.ld_reduce is a ~2 us round trip)
  for (int i0 = tid; i0 - (tid & 31) < rows * vec_per_row; i0 += 128 * U) {  // warp-uniform trip count (shuffles below)
    int4 red[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 128;
      const int r = i / vec_per_row, v = i % vec_per_row;
      const int col = tn * BN + v * VN;
      ok[u] = i < rows * vec_per_row && col < N;
      red[u] = make_int4(0, 0, 0, 0);
      if (ok[u] && ar.mc_stage) {
        const int64_t off = (int64_t(lo + r) * ldc + col) * sizeof(OutT);
        if constexpr (std::is_same<OutT, __half>::value) red[u] = ptx::multimem_ld_reduce_f16x8(ar.mc_stage + off);
        else red[u] = ptx::multimem_ld_reduce_bf16x8(ar.mc_stage + off);
      }
    }
    // lanes that share a row (vec_per_row consecutive lanes when it is a power of two <= 32) combine their squares with
    // shuffles; one shared-memory atomic per row group instead of one per lane (32-way serialised before: 3x the kernel time)
    const int gw = (vec_per_row & (vec_per_row - 1)) == 0 ? (vec_per_row < 32 ? vec_per_row : 32) : 1;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 128;
      const int r = i / vec_per_row, v = i % vec_per_row;
      const int col = tn * BN + v * VN;
      const int64_t off = (int64_t(lo + r) * ldc + col) * sizeof(OutT);
      float sq = 0.f;
      if (ok[u]) {
      float accv[VN];
      if (ar.mc_stage) {
        const OutT* h = reinterpret_cast<const OutT*>(&red[u]);
#pragma unroll
        for (int e = 0; e < VN; ++e) accv[e] = to_f32(h[e]);
      } else {
#pragma unroll
        for (int e = 0; e < VN; ++e) accv[e] = 0.f;
        for (int p = 0; p < ar.world; ++p) {
          int4 x;
          asm volatile("ld.global.relaxed.sys.v4.s32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                       : "l"(ar.peer_stage[p] + off)
                       : "memory");
          const OutT* h = reinterpret_cast<const OutT*>(&x);
#pragma unroll
          for (int e = 0; e < VN; ++e) accv[e] += to_f32(h[e]);
        }
      }
      const int64_t ooff = int64_t(lo + r - ar.rank * ar.rows_per_rank) * ar.ldo + col;
      if (ar.residual) {
        const Vec16<OutT> rv = ld16(reinterpret_cast<const OutT*>(ar.residual) + ooff);
#pragma unroll
        for (int e = 0; e < VN; ++e) accv[e] += to_f32(rv.v[e]);
      }
      Vec16<OutT> o;
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        o.v[e] = from_f32<OutT>(accv[e]);
        sq += accv[e] * accv[e];
      }
      st16(out + ooff, o);
      }
      if (ar.sumsq) {
        for (int o2 = gw >> 1; o2 > 0; o2 >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o2);
        if (ok[u] && (gw == 1 || (v & (gw - 1)) == 0)) atomicAdd(&s_sq[lo + r - tm * BM], sq);
      }
    }
  }
  if (ar.sumsq) {
    asm volatile("bar.sync 2, 128;" ::: "memory");
    const int row = tm * BM + tid;
    if (row >= lo && row < hi) atomicAdd(ar.sumsq + (row - ar.rank * ar.rows_per_rank), s_sq[tid]);
  }
  asm volatile("bar.sync 2, 128;" ::: "memory");  // s_sq is reused by the next tile
}

// RMSNorm over a reduce-scattered shard whose per-row sum of squares was accumulated by the GEMM kernel's reduce warps.
// x [rows, n] (already sum + residual), out = x * rsqrt(sumsq / n + eps) * weight; sumsq is reset for the next call.
template <typename T>
__global__ void __launch_bounds__(256)
rs_norm_kernel(const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ weight, float* __restrict__ sumsq, int n,
               int64_t ldx, int64_t ldy, float eps) {
  constexpr int VN = 16 / sizeof(T);
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();
  const int row = blockIdx.x;
  const float scale = rsqrtf(sumsq[row] / float(n) + eps);
  __syncthreads();
  if (threadIdx.x == 0) sumsq[row] = 0.f;
  for (int c = threadIdx.x * VN; c < n; c += blockDim.x * VN) {
    const Vec16<T> v = ld16(x + int64_t(row) * ldx + c);
    const Vec16<T> w = ld16(weight + c);
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < VN; ++e) o.v[e] = from_f32<T>(to_f32(v.v[e]) * scale * to_f32(w.v[e]));
    st16(y + int64_t(row) * ldy + c, o);
  }
}

template <typename OutT>
__global__ void __launch_bounds__(384, 1)
gemm_ar_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, OutT* __restrict__ stage,
               OutT* __restrict__ out, int M, int N, int K, int64_t ldc, int BN, uint32_t idesc, const ARP ar) {
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

  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;
  ptx::grid_dep_wait();

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage_i = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int tm = t % tiles_m, tn = t / tiles_m;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage_i], phase ^ 1);
          uint8_t* sa = smem + stage_i * S.stage_bytes;
          uint8_t* sb = sa + S.a_bytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage_i], S.a_bytes + BN * BK * 2);
          ptx::tma_load_2d(sa, &tmA, &full_bar[stage_i], kb * BK, tm * BM, ptx::kEvictNormal);
          ptx::tma_load_2d(sb, &tmW, &full_bar[stage_i], kb * BK, tn * BN, ptx::kEvictFirst);
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
    // ---------------- epilogue: TMEM -> staging C (symmetric) -> bump the tile flag on every rank
    const int q = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int tm = t % tiles_m, tn = t / tiles_m;
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
        if (row < M && col0 < N) {
          OutT* dst = stage + int64_t(row) * ldc + col0;
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
      __threadfence_system();
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the 4 epilogue warps
      if (warp == 4 && lane == 0) {
        ptx::mbar_arrive(&tmem_empty[acc]);
        if (ar.mc_flags) {
          ptx::multimem_red_add_u32(ar.mc_flags + t, 1u);
        } else {
          for (int p = 0; p < ar.world; ++p) ptx::red_add_release_sys(ar.peer_flags[p] + t, 1u);
        }
      } else if (lane == 0) {
        ptx::mbar_arrive(&tmem_empty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    ptx::grid_dep_launch();
  } else if (warp >= 8) {
    // ---------------- all-reduce warps
    constexpr int VN = 16 / sizeof(OutT);
    const int tid = threadIdx.x - 256;  // 0..127
    const int vec_per_row = BN / VN;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int tm = t % tiles_m, tn = t / tiles_m;
      const uint32_t want = ar.expect[t] + uint32_t(ar.world);
      if (ar.two_shot == 2) {
        rs_reduce_tile<OutT>(ar, out, tm, tn, M, N, BN, ldc, want, t, tid, reinterpret_cast<float*>(smem + S.bar_offset + 320));
        if (tid == 0) ar.expect[t] = want;
        continue;
      }
      if (tid == 0) {
        while (int32_t(ptx::ld_acquire_sys(ar.peer_flags[ar.rank] + t) - want) < 0) {
        }
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      const bool mine = !ar.two_shot || (t % ar.world) == ar.rank;
      if (mine) {
        const int rows = min(BM, M - tm * BM);
        constexpr int U = 8;  // switch round trips in flight per thread
        for (int i0 = tid; i0 < rows * vec_per_row; i0 += 128 * U) {
          int4 red[U];
          bool ok[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = i0 + u * 128;
            const int r = i / vec_per_row, v = i % vec_per_row;
            const int col = tn * BN + v * VN;
            ok[u] = i < rows * vec_per_row && col < N;
            red[u] = make_int4(0, 0, 0, 0);
            if (ok[u] && ar.mc_stage) {
              const int64_t off = (int64_t(tm * BM + r) * ldc + col) * sizeof(OutT);
              if constexpr (std::is_same<OutT, __half>::value) red[u] = ptx::multimem_ld_reduce_f16x8(ar.mc_stage + off);
              else red[u] = ptx::multimem_ld_reduce_bf16x8(ar.mc_stage + off);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const int i = i0 + u * 128;
            const int r = i / vec_per_row, v = i % vec_per_row;
            const int col = tn * BN + v * VN;
            const int64_t off = (int64_t(tm * BM + r) * ldc + col) * sizeof(OutT);
            int4 res = red[u];
            if (!ar.mc_stage) {
              float accv[VN];
#pragma unroll
              for (int e = 0; e < VN; ++e) accv[e] = 0.f;
              for (int p = 0; p < ar.world; ++p) {
                int4 x;
                asm volatile("ld.global.relaxed.sys.v4.s32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                             : "l"(ar.peer_stage[p] + off)
                             : "memory");
                const OutT* h = reinterpret_cast<const OutT*>(&x);
#pragma unroll
                for (int e = 0; e < VN; ++e) accv[e] += to_f32(h[e]);
              }
              OutT* h = reinterpret_cast<OutT*>(&res);
#pragma unroll
              for (int e = 0; e < VN; ++e) h[e] = from_f32<OutT>(accv[e]);
            }
            if (!ar.two_shot) {
              *reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(out) + off) = res;
            } else if (ar.mc_out) {
              ptx::multimem_st_v4(ar.mc_out + off, res);
            } else {
              for (int p = 0; p < ar.world; ++p) *reinterpret_cast<int4*>(ar.peer_out[p] + off) = res;
            }
          }
        }
      }
      if (tid == 0) ar.expect[t] = want;
    }
    if (ar.two_shot == 1) {
      // every rank must see all owner-written tiles before the kernel completes
      __threadfence_system();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (tid < ar.world) {
        const uint32_t epoch = ar.done_epoch[blockIdx.x] + 1;
        ptx::st_release_sys(ar.peer_done[tid] + blockIdx.x * ar.world + ar.rank, epoch);
        while (int32_t(ptx::ld_acquire_sys(ar.peer_done[ar.rank] + blockIdx.x * ar.world + tid) - epoch) < 0) {
        }
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (tid == 0) ar.done_epoch[blockIdx.x] += 1;
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

// tables: host int64 arrays [world].  stage / out: local pointers of the symmetric staging C and of the output
// (out must be symmetric for two_shot).  expect: local uint32[max_tiles]; done_epoch: local uint32[grid].
extern "C" int gemm_allreduce_nt(void* A, void* W, void* stage, void* out, int64_t M, int64_t N, int64_t K, int64_t lda,
                                 int64_t ldw, int64_t ldc, int64_t dtype, void* peer_stage_tab, void* peer_flags_tab,
                                 void* peer_out_tab, void* peer_done_tab, void* mc_stage, void* mc_flags, void* mc_out,
                                 void* expect, void* done_epoch, int64_t rank, int64_t world, int64_t two_shot,
                                 int64_t max_tiles, int64_t bn, int64_t rows_per_rank, int64_t ldo, void* residual,
                                 void* sumsq, int64_t pdl, int64_t stream_) {
  FIB_CHECK(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0 && N % 16 == 0, "gemm_allreduce: K/lda/ldw/ldc must be multiples of 8 and N of 16");
  FIB_CHECK(dtype == kF16 || dtype == kBF16, "gemm_allreduce: dtype must be f16/bf16");
  FIB_CHECK(world >= 1 && world <= kMaxRanks, "gemm_allreduce: world size must be in [1,16]");
  if (M == 0 || N == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const CUtensorMapDataType dt = dtype == kF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int tiles_m = int((M + BM - 1) / BM);
  int BN = (int)bn;
  if (BN == 0) {
    const int64_t want = (N * tiles_m + num_sms() - 1) / num_sms();
    BN = int((want + 15) / 16 * 16);
    if (BN < 32) BN = 32;
    if (BN > 256) BN = 256;
  }
  FIB_CHECK(BN % 16 == 0 && BN >= 16 && BN <= 256, "gemm_allreduce: bad N tile");
  const int64_t tiles = int64_t(tiles_m) * ((N + BN - 1) / BN);
  FIB_CHECK(tiles <= max_tiles, "gemm_allreduce: flag array too small for this problem");
  CUtensorMap tmA, tmW;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {BK, BM};
    if (make_tmap(&tmA, dt, 2, A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)BN};
    if (make_tmap(&tmW, dt, 2, W, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  ARP ar;
  for (int i = 0; i < world; ++i) {
    ar.peer_stage[i] = reinterpret_cast<uint8_t*>(((const int64_t*)peer_stage_tab)[i]);
    ar.peer_flags[i] = reinterpret_cast<uint32_t*>(((const int64_t*)peer_flags_tab)[i]);
    ar.peer_out[i] = peer_out_tab ? reinterpret_cast<uint8_t*>(((const int64_t*)peer_out_tab)[i]) : nullptr;
    ar.peer_done[i] = reinterpret_cast<uint32_t*>(((const int64_t*)peer_done_tab)[i]);
  }
  ar.mc_stage = (uint8_t*)mc_stage;
  ar.mc_flags = (uint32_t*)mc_flags;
  ar.mc_out = (uint8_t*)mc_out;
  ar.expect = (uint32_t*)expect;
  ar.done_epoch = (uint32_t*)done_epoch;
  ar.rank = (int)rank;
  ar.world = (int)world;
  ar.two_shot = (int)two_shot;
  ar.rows_per_rank = (int)rows_per_rank;
  ar.ldo = ldo;
  ar.residual = residual;
  ar.sumsq = (float*)sumsq;
  if (two_shot == 2) FIB_CHECK(rows_per_rank > 0 && ldo % 8 == 0, "gemm_reduce_scatter: bad shard geometry");
  const GSmem S = GSmem::make(BN);
  const uint32_t idesc = ptx::make_idesc_f16(dtype == kF16 ? ptx::kFmtF16 : ptx::kFmtBF16, BM, BN, 0, 0);
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  LaunchCfg lc(dim3(grid), dim3(384), S.total, stream, pdl != 0);
  if (dtype == kF16) {
    static bool set = false;
    if (!set) {
      FIB_CUDA_CHECK(cudaFuncSetAttribute(gemm_ar_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set = true;
    }
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, gemm_ar_kernel<__half>, tmA, tmW, (__half*)stage, (__half*)out, (int)M, (int)N,
                                      (int)K, ldc, BN, idesc, ar));
  } else {
    static bool set = false;
    if (!set) {
      FIB_CUDA_CHECK(cudaFuncSetAttribute(gemm_ar_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set = true;
    }
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, gemm_ar_kernel<__nv_bfloat16>, tmA, tmW, (__nv_bfloat16*)stage,
                                      (__nv_bfloat16*)out, (int)M, (int)N, (int)K, ldc, BN, idesc, ar));
  }
  return 0;
}

// Second half of GEMM -> reduce-scatter -> add-RMSNorm: normalise the local shard with the sums the GEMM kernel left behind.
extern "C" int rs_rmsnorm(void* x, void* y, void* weight, void* sumsq, int64_t rows, int64_t n, int64_t ldx, int64_t ldy,
                          double eps, int64_t dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(dtype == kF16 || dtype == kBF16, "rs_rmsnorm: dtype must be f16/bf16");
  FIB_CHECK(n % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "rs_rmsnorm: n / strides must be multiples of 8");
  if (rows == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LaunchCfg lc(dim3((unsigned)rows), dim3(256), 0, stream, pdl != 0);
  if (dtype == kF16) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rs_norm_kernel<__half>, (const __half*)x, (__half*)y, (const __half*)weight,
                                      (float*)sumsq, (int)n, ldx, ldy, (float)eps));
  } else {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rs_norm_kernel<__nv_bfloat16>, (const __nv_bfloat16*)x, (__nv_bfloat16*)y,
                                      (const __nv_bfloat16*)weight, (float*)sumsq, (int)n, ldx, ldy, (float)eps));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Prefill-size GEMM -> reduce-scatter: chunk-pipelined.  For M in the thousands the one-kernel version above is bound by
// how many in-switch reductions its 4 reduce warps keep in flight; here the row range of every rank is cut into chunks, the
// persistent tcgen05 GEMM (gemm_nt) writes chunk c of every rank's rows into the symmetric staging buffer on the main stream
// and THIS kernel pulls the chunk on a side stream while the GEMM of chunk c + 1 runs: all 148 SMs issue
// `multimem.ld_reduce` (4 per thread in flight; it co-resides with the GEMM CTAs: no shared memory, 512 threads), adds the
// residual shard, writes the local shard and accumulates the per-row sums of squares for rs_rmsnorm.  A per-CTA cross-rank
// epoch barrier at the top (with a watchdog) makes sure every rank has stored the chunk.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct RSPull {
  const uint8_t* mc_stage;          // multicast alias of the staging buffer (or null)
  const uint8_t* peer_stage[kMaxRanks];
  uint32_t* peer_sig[kMaxRanks];    // [max_ctas][world] epoch slots of every rank
  uint32_t* epochs;                 // local [max_ctas]
  int rank, world;
  int64_t ld;                       // staging row pitch (elements)
  int row_lo, rows, n;              // global rows [row_lo, row_lo + rows) of the staging buffer are mine in this chunk
  void* out;                        // local shard rows [out_row, out_row + rows)
  int64_t ldo;
  const void* residual;
  int64_t ldr;
  float* sumsq;                     // per local row (already offset to the chunk's first row) or null
};

template <typename T>
__global__ void __launch_bounds__(512) rs_pull_kernel(const RSPull p) {
  constexpr int VN = 16 / sizeof(T);
  constexpr int U = 4;
  // ---- every rank has finished storing this chunk (stream order on each rank + this barrier)
  __syncthreads();
  if (int(threadIdx.x) < p.world) {
    const int peer = threadIdx.x;
    const uint32_t epoch = p.epochs[blockIdx.x] + 1;
    __threadfence_system();
    ptx::st_release_sys(p.peer_sig[peer] + blockIdx.x * p.world + p.rank, epoch);
    ptx::spin_until_ge_sys(p.peer_sig[p.rank] + blockIdx.x * p.world + peer, epoch);
  }
  __syncthreads();
  if (threadIdx.x == 0) p.epochs[blockIdx.x] += 1;
  const int vpr = p.n / VN;
  const int64_t total = int64_t(p.rows) * vpr;
  T* out = reinterpret_cast<T*>(p.out);
  const T* residual = reinterpret_cast<const T*>(p.residual);
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i0 - (threadIdx.x & 31) < total; i0 += stride * U) {
    int4 red[U];
    bool ok[U];
    int r[U], col[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + int64_t(u) * stride;
      ok[u] = i < total;
      r[u] = ok[u] ? int(i / vpr) : 0;
      col[u] = ok[u] ? int(i - int64_t(r[u]) * vpr) * VN : 0;
      red[u] = make_int4(0, 0, 0, 0);
      if (ok[u] && p.mc_stage) {
        const uint8_t* a = p.mc_stage + (int64_t(p.row_lo + r[u]) * p.ld + col[u]) * sizeof(T);
        if constexpr (std::is_same<T, __half>::value) red[u] = ptx::multimem_ld_reduce_f16x8(a);
        else red[u] = ptx::multimem_ld_reduce_bf16x8(a);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float sq = 0.f;
      if (ok[u]) {
        float acc[VN];
        if (p.mc_stage) {
          const T* h = reinterpret_cast<const T*>(&red[u]);
#pragma unroll
          for (int e = 0; e < VN; ++e) acc[e] = to_f32(h[e]);
        } else {
#pragma unroll
          for (int e = 0; e < VN; ++e) acc[e] = 0.f;
          for (int q = 0; q < p.world; ++q) {
            const int4 x = ptx::ld_volatile_v4(p.peer_stage[(p.rank + q) % p.world] + (int64_t(p.row_lo + r[u]) * p.ld + col[u]) * sizeof(T));
            const T* h = reinterpret_cast<const T*>(&x);
#pragma unroll
            for (int e = 0; e < VN; ++e) acc[e] += to_f32(h[e]);
          }
        }
        if (residual) {
          const Vec16<T> rv = ld16(residual + int64_t(r[u]) * p.ldr + col[u]);
#pragma unroll
          for (int e = 0; e < VN; ++e) acc[e] += to_f32(rv.v[e]);
        }
        Vec16<T> o;
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          o.v[e] = from_f32<T>(acc[e]);
          sq += acc[e] * acc[e];
        }
        st16(out + int64_t(r[u]) * p.ldo + col[u], o);
      }
      if (p.sumsq) {
        // a warp covers 32 consecutive vectors: one row when vpr % 32 == 0 (checked on the host) -> one atomic per warp
        sq = warp_reduce_sum(sq);
        if ((threadIdx.x & 31) == 0 && ok[u]) atomicAdd(p.sumsq + r[u], sq);
      }
    }
  }
}

}  // namespace

// One chunk of the pipelined GEMM -> reduce-scatter: pull rows [row_lo, row_lo + rows) of the symmetric staging buffer.
extern "C" int rs_pull_rows(void* mc_stage, void* peer_stage_host, void* peer_sig_host, void* epochs, int64_t rank, int64_t world,
                            int64_t ld, int64_t row_lo, int64_t rows, int64_t n, void* out, int64_t ldo, void* residual, int64_t ldr,
                            void* sumsq, int64_t max_ctas, int64_t dtype, int64_t stream_) {
  FIB_CHECK(dtype == kF16 || dtype == kBF16, "rs_pull_rows: dtype must be f16 / bf16");
  FIB_CHECK(world >= 1 && world <= kMaxRanks && n % 256 == 0 && ld % 8 == 0 && ldo % 8 == 0 && ldr % 8 == 0,
            "rs_pull_rows: n must be a multiple of 256 (one row per warp step), strides multiples of 8");
  if (rows == 0) return 0;
  RSPull p;
  memset(&p, 0, sizeof(p));
  p.mc_stage = reinterpret_cast<const uint8_t*>(mc_stage);
  const int64_t* ps = reinterpret_cast<const int64_t*>(peer_stage_host);
  const int64_t* pg = reinterpret_cast<const int64_t*>(peer_sig_host);
  for (int i = 0; i < world; ++i) {
    p.peer_stage[i] = reinterpret_cast<const uint8_t*>(ps[i]);
    p.peer_sig[i] = reinterpret_cast<uint32_t*>(pg[i]);
  }
  p.epochs = reinterpret_cast<uint32_t*>(epochs);
  p.rank = int(rank); p.world = int(world); p.ld = ld; p.row_lo = int(row_lo); p.rows = int(rows); p.n = int(n);
  p.out = out; p.ldo = ldo; p.residual = residual; p.ldr = ldr; p.sumsq = reinterpret_cast<float*>(sumsq);
  int grid = num_sms();
  if (grid > max_ctas) grid = int(max_ctas);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LaunchCfg lc(dim3(grid), dim3(512), 0, stream, false);
  if (dtype == kF16) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rs_pull_kernel<__half>, p));
  } else {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rs_pull_kernel<__nv_bfloat16>, p));
  }
  return 0;
}
