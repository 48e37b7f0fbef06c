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
  static constexpr int kTotal = kBarOffset + 320 + 1024;  // + barriers + flags + alignment slack
};

// Stream-K work decomposition: the flattened (tile, k-block) space is cut into equal contiguous
// ranges, one per CTA.  A tile that straddles CTA boundaries is finished by a deterministic in-kernel
// fix-up: every part writes its fp32 partial to an L2-resident workspace slot and bumps the tile's
// counter; the last arriver sums the slots in slot order (bitwise reproducible) and writes the output.
struct StreamK {
  int tiles_a, tiles_b, kblocks, grid;
  int64_t units;
  int max_parts;
  __device__ __forceinline__ int64_t begin(int c) const { return (int64_t(c) * units) / grid; }
  __device__ __forceinline__ int cta_of(int64_t u) const { return int(((u + 1) * grid + units - 1) / units) - 1; }
};

template <int BN, bool kSwap, typename OutT>
__global__ void __launch_bounds__(256, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, OutT* __restrict__ C,
               float* __restrict__ partial, int* __restrict__ counters, const OutT* __restrict__ bias, int rowsA,
               int rowsB, int K, int64_t ldc, const StreamK sk, uint32_t idesc) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  int* s_flag = reinterpret_cast<int*>(tmem_ptr + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t u_begin = sk.begin(blockIdx.x);
  const int64_t u_end = sk.begin(blockIdx.x + 1);

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
      for (int64_t u = u_begin; u < u_end;) {
        const int t = int(u / sk.kblocks);
        const int kb0 = int(u % sk.kblocks);
        const int kb1 = int((kb0 + (u_end - u)) < int64_t(sk.kblocks) ? (kb0 + (u_end - u)) : int64_t(sk.kblocks));
        const int ta = t % sk.tiles_a, tb = t / sk.tiles_a;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          ptx::tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, ta * BM, kSwap ? ptx::kEvictFirst : ptx::kEvictNormal);
          ptx::tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, tb * BN, kSwap ? ptx::kEvictLast : ptx::kEvictFirst);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        u += kb1 - kb0;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t u = u_begin; u < u_end;) {
      const int kb0 = int(u % sk.kblocks);
      const int kb1 = int((kb0 + (u_end - u)) < int64_t(sk.kblocks) ? (kb0 + (u_end - u)) : int64_t(sk.kblocks));
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
      u += kb1 - kb0;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    ptx::grid_dep_launch();
  } else if (warp >= 4) {
    // ===================== epilogue (+ stream-K fix-up) =====================
    const int q = warp - 4;  // TMEM lane quadrant == warp % 4
    const int etid = threadIdx.x - 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr int CH = (BN >= 32) ? 32 : 16;
    for (int64_t u = u_begin; u < u_end;) {
      const int t = int(u / sk.kblocks);
      const int kb0 = int(u % sk.kblocks);
      const int kb1 = int((kb0 + (u_end - u)) < int64_t(sk.kblocks) ? (kb0 + (u_end - u)) : int64_t(sk.kblocks));
      u += kb1 - kb0;
      const int ta = t % sk.tiles_a, tb = t / sk.tiles_a;
      const bool full_tile = (kb0 == 0 && kb1 == sk.kblocks);
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int r_in_tile = q * 32 + lane;
      const int a_row = ta * BM + r_in_tile;  // row of the A-side operand owned by this thread
      const uint32_t taddr = tmem_base + acc * BN + (uint32_t(q * 32) << 16);

      int parts = 1, c_first = 0;
      bool i_am_last = true;
      float* my_slot = nullptr;
      if (!full_tile) {
        c_first = sk.cta_of(int64_t(t) * sk.kblocks);
        const int c_last = sk.cta_of(int64_t(t + 1) * sk.kblocks - 1);
        parts = c_last - c_first + 1;
        float* tile_ws = partial + int64_t(c_first) * sk.max_parts * (BM * BN);
        my_slot = tile_ws + int64_t(blockIdx.x - c_first) * (BM * BN);
        // 1) publish my partial
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CH) {
          uint32_t r[CH];
          if constexpr (CH == 32) ptx::tmem_ld_x32(taddr + c0, r); else ptx::tmem_ld_x16(taddr + c0, r);
          ptx::tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(my_slot + r_in_tile * BN + c0);
#pragma unroll
          for (int j = 0; j < CH; j += 4)
            __stcg(dst + j / 4, make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                            __uint_as_float(r[j + 3])));
        }
        __threadfence();
        ptx::named_bar_sync(1, 128);
        if (etid == 0) {
          const int old = atomicAdd(&counters[c_first], 1);
          const int last = (old == parts - 1);
          if (last) counters[c_first] = 0;  // self-reset for the next launch / graph replay
          *s_flag = last;
        }
        ptx::named_bar_sync(1, 128);
        i_am_last = (*s_flag != 0);
        if (i_am_last) __threadfence();
      }

      if (i_am_last) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CH) {
          float v[CH];
          if (full_tile) {
            uint32_t r[CH];
            if constexpr (CH == 32) ptx::tmem_ld_x32(taddr + c0, r); else ptx::tmem_ld_x16(taddr + c0, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]);
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = 0.f;
            const float* tile_ws = partial + int64_t(c_first) * sk.max_parts * (BM * BN);
            for (int pth = 0; pth < parts; ++pth) {
              const float4* src = reinterpret_cast<const float4*>(tile_ws + int64_t(pth) * (BM * BN) + r_in_tile * BN + c0);
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                const float4 x = __ldcg(src + j / 4);
                v[j] += x.x;
                v[j + 1] += x.y;
                v[j + 2] += x.z;
                v[j + 3] += x.w;
              }
            }
          }
          const int b_row0 = tb * BN + c0;
          if constexpr (kSwap) {
            // C[b_row][a_row]: consecutive lanes -> consecutive a_row -> coalesced 2B stores
            if (a_row < rowsA) {
              const float bv = bias ? to_f32(bias[a_row]) : 0.f;
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (b_row0 + j < rowsB) C[int64_t(b_row0 + j) * ldc + a_row] = from_f32<OutT>(v[j] + bv);
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
                    float x = v[j + e];
                    if (bias) x += to_f32(bias[b_row0 + j + e]);
                    o.v[e] = from_f32<OutT>(x);
                  }
                  st16(dst + j, o);
                }
              } else {
#pragma unroll
                for (int j = 0; j < CH; ++j)
                  if (b_row0 + j < rowsB) {
                    float x = v[j];
                    if (bias) x += to_f32(bias[b_row0 + j]);
                    dst[j] = from_f32<OutT>(x);
                  }
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

template <int BN, bool kSwap, typename OutT>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, OutT* C, float* workspace, int64_t workspace_bytes,
                const OutT* bias, int rowsA, int rowsB, int K, int64_t ldc, bool f16, bool pdl, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  auto kern = gemm_nt_kernel<BN, kSwap, OutT>;
  static bool attr_set = false;
  if (!attr_set) {
    FIB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set = true;
  }
  StreamK sk;
  sk.tiles_a = (rowsA + BM - 1) / BM;
  sk.tiles_b = (rowsB + BN - 1) / BN;
  sk.kblocks = (K + BK - 1) / BK;
  const int tiles = sk.tiles_a * sk.tiles_b;
  sk.units = int64_t(tiles) * sk.kblocks;
  // workspace layout: [counters: 1024 ints][partials]
  const int64_t slot_bytes = int64_t(BM) * BN * 4;
  const int64_t ws_partial = workspace ? workspace_bytes - 4096 : 0;
  int grid = num_sms();
  if (sk.units < grid) grid = (int)sk.units;
  const bool streamk = workspace != nullptr && (tiles % grid != 0);
  auto parts_for = [&](int g) {
    const int64_t per = sk.units / g > 0 ? sk.units / g : 1;
    return int((sk.kblocks + per - 1) / per) + 1;
  };
  if (streamk) {
    // at least 4 k-blocks per part, and the fix-up slots must fit the workspace
    while (grid > 1 && (sk.units / grid < 4 || int64_t(grid) * parts_for(grid) * slot_bytes > ws_partial)) --grid;
    sk.max_parts = parts_for(grid);
  } else {
    sk.max_parts = 1;
    if (tiles % grid != 0) {
      // no workspace: keep CTA ranges tile-aligned -> largest grid <= #SM that divides the tile count
      for (int gtry = grid; gtry >= 1; --gtry)
        if (tiles % gtry == 0) {
          grid = gtry;
          break;
        }
    }
  }
  sk.grid = grid;
  const uint32_t idesc = ptx::make_idesc_f16(f16 ? ptx::kFmtF16 : ptx::kFmtBF16, BM, BN, 0, 0);
  LaunchCfg lc(dim3(grid), dim3(256), S::kTotal, stream, pdl);
  int* counters = reinterpret_cast<int*>(workspace);
  float* partial = workspace ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 4096) : nullptr;
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, tmA, tmB, C, partial, counters, bias, rowsA, rowsB, K, ldc, sk, idesc));
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
#define FIB_LAUNCH(BN_, SWAP_)                                                                                 \
  if (launch_gemm<BN_, SWAP_, OutT>(tmA, tmB, C, workspace, workspace_bytes, bias, rowsA, rowsB, K, ldc, f16, pdl, \
                                    stream))                                                                   \
    return 1;
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
