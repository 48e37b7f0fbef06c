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
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

constexpr int BM = 128;  // MMA M (rows of the "A-side" operand)
constexpr int BK = 64;   // 64 x 2B = one 128B swizzle span

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN <= 64) ? 8 : (BN <= 128 ? 6 : 4);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // + barriers + alignment slack
};

template <int BN, bool kSwap, typename OutT, bool kSplitK>
__global__ void __launch_bounds__(256, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, OutT* __restrict__ C,
               float* __restrict__ partial, const OutT* __restrict__ bias, int rowsA, int rowsB, int K, int64_t ldc,
               int splits, uint32_t idesc) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_a = (rowsA + BM - 1) / BM;
  const int tiles_b = (rowsB + BN - 1) / BN;
  const int num_kb_total = (K + BK - 1) / BK;
  const int kb_per_split = (num_kb_total + splits - 1) / splits;
  const int num_tiles = tiles_a * tiles_b * splits;

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
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_ptr, (2 * BN < 32) ? 32 : 2 * BN);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  ptx::grid_dep_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int ta = t % tiles_a;
        const int rest = t / tiles_a;
        const int tb = rest % tiles_b;
        const int sp = rest / tiles_b;
        const int kb0 = sp * kb_per_split;
        const int kb1 = min(num_kb_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          // weights are streamed once (evict-first when they are the A side of a swapped GEMM)
          ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, ta * BM, kSwap ? ptx::kEvictFirst : ptx::kEvictNormal);
          ptx::tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, tb * BN, kSwap ? ptx::kEvictLast : ptx::kEvictFirst);
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
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int sp = (t / tiles_a) / tiles_b;
      const int kb0 = sp * kb_per_split;
      const int kb1 = min(num_kb_total, kb0 + kb_per_split);
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t sa = ptx::smem_u32(smem + stage * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
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
      if (kb1 <= kb0) {
        // empty split (cannot happen with the host-side split choice, kept for safety)
        if (ptx::elect_one()) ptx::mma_commit(&tmem_full[acc]);
        __syncwarp();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    ptx::grid_dep_launch();
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp - 4;  // TMEM lane quadrant == warp % 4
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int ta = t % tiles_a;
      const int rest = t / tiles_a;
      const int tb = rest % tiles_b;
      const int sp = rest / tiles_b;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int a_row = ta * BM + q * 32 + lane;  // row of the A-side operand owned by this thread
      const uint32_t taddr = tmem_base + acc * BN + (uint32_t(q * 32) << 16);
      constexpr int CH = (BN >= 32) ? 32 : 16;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CH) {
        uint32_t r[CH];
        if constexpr (CH == 32)
          ptx::tmem_ld_x32(taddr + c0, r);
        else
          ptx::tmem_ld_x16(taddr + c0, r);
        ptx::tmem_ld_wait();
        const int b_row0 = tb * BN + c0;
        if constexpr (kSplitK) {
          // fp32 partials laid out [split][rowsB][rowsA] (swap) or [split][rowsA][rowsB]
          if (a_row < rowsA) {
            if constexpr (kSwap) {
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (b_row0 + j < rowsB)
                  partial[(int64_t(sp) * rowsB + (b_row0 + j)) * rowsA + a_row] = __uint_as_float(r[j]);
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (b_row0 + j < rowsB)
                  partial[(int64_t(sp) * rowsA + a_row) * rowsB + b_row0 + j] = __uint_as_float(r[j]);
            }
          }
        } else if constexpr (kSwap) {
          // C[b_row][a_row]: consecutive lanes -> consecutive a_row -> coalesced 2B stores
          if (a_row < rowsA) {
            const float bv = bias ? to_f32(bias[a_row]) : 0.f;
#pragma unroll
            for (int j = 0; j < CH; ++j) {
              if (b_row0 + j < rowsB) C[int64_t(b_row0 + j) * ldc + a_row] = from_f32<OutT>(__uint_as_float(r[j]) + bv);
            }
          }
        } else {
          // C[a_row][b_row..]: each thread owns one output row; 16B vector stores
          if (a_row < rowsA) {
            OutT* dst = C + int64_t(a_row) * ldc + b_row0;
            if (b_row0 + CH <= rowsB && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
              constexpr int VN = 16 / sizeof(OutT);
#pragma unroll
              for (int j = 0; j < CH; j += VN) {
                Vec16<OutT> v;
#pragma unroll
                for (int e = 0; e < VN; ++e) {
                  float x = __uint_as_float(r[j + e]);
                  if (bias) x += to_f32(bias[b_row0 + j + e]);
                  v.v[e] = from_f32<OutT>(x);
                }
                st16(dst + j, v);
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
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, (2 * BN < 32) ? 32 : 2 * BN);
  }
}

// Reduce split-K fp32 partials into the output (+bias).  partial layout [splits][R0][R1].
template <typename OutT>
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, OutT* __restrict__ C, const OutT* __restrict__ bias,
                                     int splits, int R0, int R1, int64_t ldc, int swap) {
  ptx::grid_dep_wait();
  const int64_t total = int64_t(R0) * R1;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[int64_t(s) * total + i];
    const int r0 = int(i / R1), r1 = int(i % R1);
    // both layouts: r0 = output row (token), r1 = output col (feature)
    if (bias) acc += to_f32(bias[r1]);
    C[int64_t(r0) * ldc + r1] = from_f32<OutT>(acc);
  }
  (void)swap;
}

template <int BN, bool kSwap, typename OutT, bool kSplitK>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, OutT* C, float* partial, const OutT* bias, int rowsA,
                int rowsB, int K, int64_t ldc, int splits, bool f16, bool pdl, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  auto kern = gemm_nt_kernel<BN, kSwap, OutT, kSplitK>;
  static bool attr_set = false;
  if (!attr_set) {
    FIB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set = true;
  }
  const int tiles = ((rowsA + BM - 1) / BM) * ((rowsB + BN - 1) / BN) * splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  const uint32_t idesc = ptx::make_idesc_f16(f16 ? ptx::kFmtF16 : ptx::kFmtBF16, BM, BN, 0, 0);
  LaunchCfg lc(dim3(grid), dim3(256), S::kTotal, stream, pdl);
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, tmA, tmB, C, partial, bias, rowsA, rowsB, K, ldc, splits, idesc));
  return 0;
}

template <typename OutT>
int gemm_dispatch(const void* A, const void* B, OutT* C, const OutT* bias, int M, int N, int K, int64_t lda, int64_t ldb,
                  int64_t ldc, bool f16, float* workspace, int64_t workspace_bytes, bool pdl, cudaStream_t stream) {
  const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const bool swap = (M <= 128);
  // A-side = 128-row operand.  normal: activations; swap: weights.
  const void* pa = swap ? B : A;
  const void* pb = swap ? A : B;
  const int rowsA = swap ? N : M, rowsB = swap ? M : N;
  const int64_t ldA = swap ? ldb : lda, ldB = swap ? lda : ldb;
  int BN;
  if (swap) {
    BN = M <= 16 ? 16 : (M <= 32 ? 32 : (M <= 64 ? 64 : 128));
  } else {
    BN = (N >= 256 && (int64_t(M) * N >= int64_t(256) * 256 * 64)) ? 256 : 128;
  }
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)rowsA};
    uint64_t str[1] = {(uint64_t)ldA * 2};
    uint32_t box[2] = {BK, BM};
    if (make_tmap(&tmA, dt, 2, pa, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)rowsB};
    uint64_t str[1] = {(uint64_t)ldB * 2};
    uint32_t box[2] = {BK, (uint32_t)BN};
    if (make_tmap(&tmB, dt, 2, pb, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  // split-K heuristic: fill the machine when there are too few output tiles and K is long.
  const int tiles = ((rowsA + BM - 1) / BM) * ((rowsB + BN - 1) / BN);
  const int num_kb = (K + BK - 1) / BK;
  int splits = 1;
  if (workspace && tiles * 2 <= num_sms() && num_kb >= 16) {
    splits = num_sms() / tiles;
    if (splits > num_kb / 4) splits = num_kb / 4;
    if (splits > 16) splits = 16;
    while (splits > 1 && int64_t(splits) * M * N * 4 > workspace_bytes) --splits;
    // make every split non-empty
    while (splits > 1 && ((num_kb + splits - 1) / splits) * (splits - 1) >= num_kb) --splits;
    if (splits < 1) splits = 1;
  }
#define FIB_LAUNCH(BN_, SWAP_)                                                                                       \
  if (splits > 1) {                                                                                                  \
    if (launch_gemm<BN_, SWAP_, OutT, true>(tmA, tmB, C, workspace, bias, rowsA, rowsB, K, ldc, splits, f16, pdl,    \
                                            stream))                                                                 \
      return 1;                                                                                                      \
  } else {                                                                                                           \
    if (launch_gemm<BN_, SWAP_, OutT, false>(tmA, tmB, C, workspace, bias, rowsA, rowsB, K, ldc, 1, f16, pdl,        \
                                             stream))                                                                \
      return 1;                                                                                                      \
  }
  if (swap) {
    switch (BN) {
      case 16: FIB_LAUNCH(16, true); break;
      case 32: FIB_LAUNCH(32, true); break;
      case 64: FIB_LAUNCH(64, true); break;
      default: FIB_LAUNCH(128, true); break;
    }
  } else {
    if (BN == 256) {
      FIB_LAUNCH(256, false);
    } else {
      FIB_LAUNCH(128, false);
    }
  }
#undef FIB_LAUNCH
  if (splits > 1) {
    const int64_t total = int64_t(M) * N;
    int blocks = int((total + 255) / 256);
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    LaunchCfg lc(dim3(blocks), dim3(256), 0, stream, pdl);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, splitk_reduce_kernel<OutT>, (const float*)workspace, C, bias, splits, M, N,
                                      ldc, swap ? 1 : 0));
  }
  return 0;
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
