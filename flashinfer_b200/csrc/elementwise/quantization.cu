// Block-scaled quantisation kernels for the Blackwell low-precision formats + bit packing.
//
// Parity: reference flashinfer/quantization/fp4_quantization.py:790-1683 (fp4_quantize / nvfp4_quantize /
// mxfp4_quantize / nvfp4_batched_quantize / block_scale_interleave / e2m1_and_ufp8sf_scale_to_float),
// fp8_quantization.py:163-276 (mxfp8_quantize), packbits.py:47-139 and the TRT-LLM kernels under
// csrc/nv_internal/.  Formats: NVFP4 = e2m1 pairs packed in uint8, block 16, UE4M3 block
// scale + fp32 global scale; MXFP4/MXFP8 = block 32, UE8M0 scale.  Scale-factor layouts: linear [M, K/vec]
// or the 128x4 tile-swizzled layout consumed by tcgen05 block-scaled MMA (via tcgen05.cp 32x128b).
//
// One thread quantises one scale block (16 or 32 contiguous elements: 2 or 4 x 16 B loads), so global
// traffic is fully vectorised; e2m1 conversion uses cvt.rn.satfinite.e2m1x2.f32 (sm_100a).
#include <cuda_fp4.h>
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

// offset of scale (m, kc) in the 128x4 swizzled layout; kc_pad = round_up(Kc, 4)
__host__ __device__ __forceinline__ int64_t sf_swizzled_offset(int64_t m, int64_t kc, int64_t kc_pad) {
  const int64_t tile = (m / 128) * (kc_pad / 4) + kc / 4;
  return tile * 512 + (m % 32) * 16 + ((m % 128) / 32) * 4 + (kc % 4);
}

__device__ __forceinline__ uint8_t f32_to_ue8m0_ceil(float x) {
  // smallest power of two >= x (x > 0), biased exponent; 0 -> exponent for 2^-127
  if (!(x > 0.f)) return 0;
  const uint32_t b = __float_as_uint(x);
  uint32_t e = (b >> 23) & 0xff;
  if (b & 0x7fffff) e += 1;
  if (e > 254) e = 254;
  return (uint8_t)e;
}
__device__ __forceinline__ float ue8m0_to_f32(uint8_t e) { return __uint_as_float(uint32_t(e == 0 ? 0 : e) << 23); }

// ---------------------------------------------------------------- fp4 (nvfp4 / mxfp4)
// Quantise one block of VEC values: returns the scale byte (UE4M3 or UE8M0) and stores VEC / 2 packed e2m1 bytes at dst.
template <int VEC, bool kUE8M0>
__device__ __forceinline__ uint8_t fp4_quantize_block(const float (&v)[VEC], float gs, uint8_t* __restrict__ dst) {
  float amax = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) amax = fmaxf(amax, fabsf(v[j]));
  uint8_t sf_byte;
  float out_scale;
  if constexpr (kUE8M0) {
    sf_byte = f32_to_ue8m0_ceil(amax * (1.f / 6.f) * gs);
    const float sfv = ue8m0_to_f32(sf_byte);
    out_scale = sfv > 0.f ? gs / sfv : 0.f;
  } else {
    const float sfv_f = gs * (amax * (1.f / 6.f));
    const __nv_fp8_e4m3 s8(sfv_f);
    sf_byte = *reinterpret_cast<const uint8_t*>(&s8);
    const float sfv = float(s8);
    out_scale = sfv != 0.f ? gs / sfv : 0.f;
  }
  uint8_t packed[VEC / 2];
#pragma unroll
  for (int j = 0; j < VEC; j += 2)
    packed[j / 2] = (uint8_t)__nv_cvt_float2_to_fp4x2(make_float2(v[j] * out_scale, v[j + 1] * out_scale), __NV_E2M1,
                                                      cudaRoundNearest);
  if constexpr (VEC == 16) {
    *reinterpret_cast<int2*>(dst) = *reinterpret_cast<const int2*>(packed);
  } else {
    *reinterpret_cast<int4*>(dst) = *reinterpret_cast<const int4*>(packed);
  }
  return sf_byte;
}

// x [B, M, K] (row stride ldx) -> q [B, M, K/2] uint8, sf [B][...] uint8
template <typename T, int VEC, bool kUE8M0>
__global__ void __launch_bounds__(256)
fp4_quantize_kernel(const T* __restrict__ x, uint8_t* __restrict__ q, uint8_t* __restrict__ sf,
                    const float* __restrict__ global_scale, int64_t batch, int64_t M, int64_t K, int64_t ldx,
                    int64_t x_batch_stride, int swizzled, int64_t sf_batch_stride, const int32_t* __restrict__ row_map,
                    int gather, int gated, const int32_t* __restrict__ row_list, int64_t n_list, int list_div) {
  const int64_t kc_total = K / VEC;
  const int64_t kc_pad = (kc_total + 3) / 4 * 4;
  // row_list (MoE): only the n_list live destination rows are visited (row_list[j] = permuted row of expanded entry j, -1 =
  // not local); with `gather` the source row is token j / list_div.  Otherwise all batch * M rows are walked.
  const int64_t total = row_list ? n_list * kc_total : batch * M * kc_total;
  ptx::grid_dep_wait();
  const float gs = global_scale ? __ldg(global_scale) : 1.f;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int64_t kc, rowi, m, b;
    fast_divmod(i, kc_total, rowi, kc);
    fast_divmod(rowi, M, b, m);
    if (row_list) b = 0;
    int64_t src_row = m;
    if (row_list) {
      const int64_t j = rowi;
      m = row_list[j];
      if (m < 0) continue;
      src_row = gather ? int64_t(uint32_t(j) / uint32_t(list_div)) : m;
    } else if (row_map) {
      const int rm = row_map[m];
      if (rm < 0) continue;     // MoE padding row (scale bytes stay as initialised: finite)
      if (gather) src_row = rm; // fused MoE gather: quantise x[token(m)] straight into the permuted row m
    }
    const T* src = x + b * x_batch_stride + src_row * ldx + kc * VEC;
    float v[VEC];
#pragma unroll
    for (int j = 0; j < VEC; j += 8) {
      const Vec16<T> raw = ld16(src + j);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[j + e] = to_f32(raw.v[e]);
    }
    if (gated) {
      // fused SwiGLU: the row holds [linear | gate] halves of width K; quantise silu(gate) * linear
#pragma unroll
      for (int j = 0; j < VEC; j += 8) {
        const Vec16<T> raw = ld16(src + K + j);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float g = to_f32(raw.v[e]);
          v[j + e] *= g / (1.f + __expf(-g));
        }
      }
    }
    uint8_t* dst = q + (b * M + m) * (K / 2) + kc * (VEC / 2);
    const uint8_t sf_byte = fp4_quantize_block<VEC, kUE8M0>(v, gs, dst);
    // swizzled: 0 linear [M, K / vec] | 1 the 128x4 tile layout of tcgen05 block-scaled MMA | 2 the 8x4 tile layout
    // (reference SfLayout.layout_8x4: tiles of 8 rows x 4 scale columns, 32 bytes each, row-major inside the tile)
    const int64_t sf_off = swizzled == 2 ? ((m / 8) * (kc_pad / 4) + kc / 4) * 32 + (m % 8) * 4 + (kc % 4)
                                         : (swizzled ? sf_swizzled_offset(m, kc, kc_pad) : m * kc_total + kc);
    sf[b * sf_batch_stride + sf_off] = sf_byte;
  }
  ptx::grid_dep_launch();
}

// ---------------------------------------------------------------- (add +) RMSNorm + FP4 quantisation in one kernel
// One CTA per row: pass 1 loads the row (adds and writes back the residual when given), keeps it in shared memory as fp32
// and reduces the sum of squares; pass 2 normalises, applies the weight and quantises 16- / 32-element blocks straight to
// packed e2m1 + scale bytes.  The normalised activations never touch HBM (reference: rmsnorm_fp4quant /
// add_rmsnorm_fp4quant, flashinfer/norm/__init__.py, a CuTe-DSL kernel there).
template <typename T, int VEC, bool kUE8M0>
__global__ void __launch_bounds__(256)
rmsnorm_fp4quant_kernel(const T* __restrict__ x, T* __restrict__ residual, const T* __restrict__ weight, uint8_t* __restrict__ q,
                        uint8_t* __restrict__ sf, uint8_t* __restrict__ sf2, const float* __restrict__ global_scale, int64_t M,
                        int64_t K, int64_t ldx, int64_t ldr, float eps, int swizzled, int weight_bias) {
  extern __shared__ float rowbuf[];
  __shared__ float red[8];
  const int64_t m = blockIdx.x;
  ptx::grid_dep_wait();
  float ss = 0.f;
  for (int64_t c = threadIdx.x * 8; c < K; c += blockDim.x * 8) {
    const Vec16<T> xv = ld16(x + m * ldx + c);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = to_f32(xv.v[e]);
    if (residual) {
      const Vec16<T> rv = ld16(residual + m * ldr + c);
      Vec16<T> ro;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        f[e] += to_f32(rv.v[e]);
        ro.v[e] = from_f32<T>(f[e]);
        f[e] = to_f32(ro.v[e]);  // the norm sees the stored (rounded) residual, like the two-kernel composition
      }
      st16(residual + m * ldr + c, ro);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      rowbuf[c + e] = f[e];
      ss += f[e] * f[e];
    }
  }
  ss = warp_reduce_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float rstd = rsqrtf(tot / float(K) + eps);
  const float gs = global_scale ? __ldg(global_scale) : 1.f;
  const int64_t kc_total = K / VEC, kc_pad = (kc_total + 3) / 4 * 4;
  ptx::grid_dep_launch();
  for (int64_t kc = threadIdx.x; kc < kc_total; kc += blockDim.x) {
    float v[VEC];
#pragma unroll
    for (int j = 0; j < VEC; j += 8) {
      const Vec16<T> wv = ld16(weight + kc * VEC + j);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float wf = to_f32(wv.v[e]) + (weight_bias ? 1.f : 0.f);
        // round through T like the composed path (rmsnorm writes T, the quantiser reads T)
        v[j + e] = to_f32(from_f32<T>(rowbuf[kc * VEC + j + e] * rstd * wf));
      }
    }
    const uint8_t sf_byte = fp4_quantize_block<VEC, kUE8M0>(v, gs, q + m * (K / 2) + kc * (VEC / 2));
    sf[swizzled ? sf_swizzled_offset(m, kc, kc_pad) : m * kc_total + kc] = sf_byte;
    if (sf2) sf2[swizzled ? m * kc_total + kc : sf_swizzled_offset(m, kc, kc_pad)] = sf_byte;  // the other layout as well
  }
}

// ---------------------------------------------------------------- fp8 e4m3 with fp32 scales per 1 x 128 group (DeepSeek)
// x [M, K] -> q [M, K] e4m3, scale [M, K/128] fp32 (scale = amax / 448).  8 lanes x 16 elements cover one group.
// MoE modes as in fp4_quantize_kernel: row_list (visit only live rows; `gather`: source row = j / list_div) and `gated`
// (row = [linear | gate] halves of width K, quantise silu(gate) * linear).
template <typename T>
__global__ void __launch_bounds__(256)
fp8_group_quantize_kernel(const T* __restrict__ x, __nv_fp8_e4m3* __restrict__ q, float* __restrict__ scale, int64_t M,
                          int64_t K, int64_t ldx, int gated, const int32_t* __restrict__ row_list, int64_t n_list,
                          int gather, int list_div) {
  const int64_t per_row = K / 16;  // threads per row
  const int64_t total = (row_list ? n_list : M) * per_row;
  ptx::grid_dep_wait();
  // the grid-stride loop keeps whole warps together (total is a multiple of 8 lanes per group; per_row % 8 == 0)
  for (int64_t i0 = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) & ~int64_t(31); i0 < total;
       i0 += int64_t(gridDim.x) * blockDim.x) {
    const int64_t i = i0 + (threadIdx.x & 31);
    const bool live = i < total;
    const int64_t ii = live ? i : total - 1;
    int64_t c, m;
    fast_divmod(ii, per_row, m, c);
    int64_t src_row = m;
    bool skip = !live;
    if (row_list) {
      const int64_t j = m;
      m = row_list[j];
      if (m < 0) {
        skip = true;
        m = 0;
      }
      src_row = gather ? int64_t(uint32_t(j) / uint32_t(list_div)) : m;
    }
    float v[16];
    if (!skip) {
      const T* src = x + src_row * ldx + c * 16;
#pragma unroll
      for (int j = 0; j < 16; j += 8) {
        const Vec16<T> raw = ld16(src + j);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[j + e] = to_f32(raw.v[e]);
      }
      if (gated) {
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          const Vec16<T> raw = ld16(src + K + j);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float g = to_f32(raw.v[e]);
            v[j + e] *= g / (1.f + __expf(-g));
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) amax = fmaxf(amax, fabsf(v[j]));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (skip) continue;
    const float sc = fmaxf(amax, 1e-10f) * (1.f / 448.f);
    const float inv = 1.f / sc;
    __nv_fp8_e4m3 o8[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) o8[j] = __nv_fp8_e4m3(v[j] * inv);
    *reinterpret_cast<int4*>(q + m * K + c * 16) = *reinterpret_cast<const int4*>(o8);
    if ((c & 7) == 0) scale[m * (K / 128) + c / 8] = sc;
  }
  ptx::grid_dep_launch();
}

// ---------------------------------------------------------------- mxfp8 (e4m3 + ue8m0 / 32)
template <typename T>
__global__ void __launch_bounds__(256)
mxfp8_quantize_kernel(const T* __restrict__ x, __nv_fp8_e4m3* __restrict__ q, uint8_t* __restrict__ sf, int64_t M,
                      int64_t K, int64_t ldx, int swizzled) {
  constexpr int VEC = 32;
  const int64_t kc_total = K / VEC;
  const int64_t kc_pad = (kc_total + 3) / 4 * 4;
  const int64_t total = M * kc_total;
  ptx::grid_dep_wait();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t kc = i % kc_total, m = i / kc_total;
    const T* src = x + m * ldx + kc * VEC;
    float v[VEC];
#pragma unroll
    for (int j = 0; j < VEC; j += 8) {
      const Vec16<T> raw = ld16(src + j);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[j + e] = to_f32(raw.v[e]);
    }
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) amax = fmaxf(amax, fabsf(v[j]));
    const uint8_t sf_byte = f32_to_ue8m0_ceil(amax * (1.f / 448.f));
    const float sfv = ue8m0_to_f32(sf_byte);
    const float inv = sfv > 0.f ? 1.f / sfv : 0.f;
    __nv_fp8_e4m3 out[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) out[j] = __nv_fp8_e4m3(v[j] * inv);
    int4* dst = reinterpret_cast<int4*>(q + m * K + kc * VEC);
    dst[0] = reinterpret_cast<const int4*>(out)[0];
    dst[1] = reinterpret_cast<const int4*>(out)[1];
    sf[swizzled ? sf_swizzled_offset(m, kc, kc_pad) : m * kc_total + kc] = sf_byte;
  }
  ptx::grid_dep_launch();
}

// ---------------------------------------------------------------- scale-factor re-layout + dequant
__global__ void sf_interleave_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int64_t batch, int64_t M,
                                     int64_t Kc, int64_t out_batch_stride, int to_swizzled) {
  const int64_t kc_pad = (Kc + 3) / 4 * 4;
  const int64_t total = batch * M * Kc;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t kc = i % Kc, m = (i / Kc) % M, b = i / (Kc * M);
    if (to_swizzled)
      out[b * out_batch_stride + sf_swizzled_offset(m, kc, kc_pad)] = in[i];
    else
      out[i] = in[b * out_batch_stride + sf_swizzled_offset(m, kc, kc_pad)];
  }
}

__device__ __forceinline__ float e2m1_to_f32(uint8_t nib) {
  const float mag[8] = {0.f, 0.5f, 1.f, 1.5f, 2.f, 3.f, 4.f, 6.f};
  const float v = mag[nib & 7];
  return (nib & 8) ? -v : v;
}

__global__ void fp4_dequant_kernel(const uint8_t* __restrict__ q, const uint8_t* __restrict__ sf,
                                   const float* __restrict__ global_scale, float* __restrict__ out, int64_t M, int64_t K,
                                   int vec, int ue8m0, int swizzled) {
  const int64_t kc_total = K / vec;
  const int64_t kc_pad = (kc_total + 3) / 4 * 4;
  const float gs = global_scale ? __ldg(global_scale) : 1.f;
  const int64_t total = M * K / 2;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t m = i / (K / 2), k2 = i % (K / 2);
    const int64_t kc = (k2 * 2) / vec;
    const uint8_t sb = sf[swizzled ? sf_swizzled_offset(m, kc, kc_pad) : m * kc_total + kc];
    float s;
    if (ue8m0) {
      s = ue8m0_to_f32(sb);
    } else {
      __nv_fp8_e4m3 t;
      *reinterpret_cast<uint8_t*>(&t) = sb;
      s = float(t);
    }
    s = s / gs;
    const uint8_t b = q[i];
    out[m * K + k2 * 2] = e2m1_to_f32(b & 0xf) * s;
    out[m * K + k2 * 2 + 1] = e2m1_to_f32(b >> 4) * s;
  }
}

// ---------------------------------------------------------------- packbits
// x: uint8 (0 / non-zero) -> packed bits.  Segments: output segment b starts at out_indptr[b].
__global__ void packbits_kernel(const uint8_t* __restrict__ x, uint8_t* __restrict__ y, const int32_t* __restrict__ in_indptr,
                                const int32_t* __restrict__ out_indptr, int64_t n, int num_segments, int big_endian) {
  if (num_segments == 0) {
    const int64_t nbytes = (n + 7) / 8;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < nbytes; i += int64_t(gridDim.x) * blockDim.x) {
      uint8_t v = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int64_t j = i * 8 + b;
        const int bit = (j < n && x[j]) ? 1 : 0;
        v |= bit << (big_endian ? 7 - b : b);
      }
      y[i] = v;
    }
  } else {
    for (int s = blockIdx.y; s < num_segments; s += gridDim.y) {
      const int64_t i0 = in_indptr[s], len = in_indptr[s + 1] - i0, o0 = out_indptr[s];
      const int64_t nbytes = (len + 7) / 8;
      for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < nbytes; i += int64_t(gridDim.x) * blockDim.x) {
        uint8_t v = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int64_t j = i * 8 + b;
          const int bit = (j < len && x[i0 + j]) ? 1 : 0;
          v |= bit << (big_endian ? 7 - b : b);
        }
        y[o0 + i] = v;
      }
    }
  }
}

inline int grid_for(int64_t total, int threads = 256) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = int64_t(num_sms()) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

// vec = 16 (NVFP4, UE4M3 scale unless ue8m0) or 32 (MXFP4, UE8M0 scale)
extern "C" int fp4_quantize(void* x, void* q, void* sf, void* global_scale, int64_t batch, int64_t M, int64_t K,
                            int64_t ldx, int64_t x_batch_stride, int64_t vec, int64_t ue8m0, int64_t swizzled,
                            int64_t sf_batch_stride, void* row_map, int64_t gather, int64_t gated, void* row_list,
                            int64_t n_list, int64_t list_div, int64_t dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(vec == 16 || vec == 32, "fp4_quantize: sf_vec_size must be 16 or 32");
  FIB_CHECK(K % vec == 0 && ldx % 8 == 0, "fp4_quantize: K must be a multiple of sf_vec_size and rows 16B aligned");
  if (batch * M * K == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FIB_CHECK(!row_list || (batch == 1 && list_div >= 1), "fp4_quantize: row_list needs batch == 1");
  const int64_t total = (row_list ? n_list : batch * M) * (K / vec);
  if (total == 0) return 0;
  LaunchCfg lc(dim3(grid_for(total)), dim3(256), 0, stream, pdl != 0);
  return FIB_DISPATCH_HALF(dtype, T, [&]() -> int {
    auto launch = [&](auto kern) -> int {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, (const T*)x, (uint8_t*)q, (uint8_t*)sf, (const float*)global_scale,
                                        batch, M, K, ldx, x_batch_stride, (int)swizzled, sf_batch_stride, (const int32_t*)row_map,
                                        (int)gather, (int)gated, (const int32_t*)row_list, n_list, (int)list_div));
      return 0;
    };
    if (vec == 16 && !ue8m0) return launch(fp4_quantize_kernel<T, 16, false>);
    if (vec == 16) return launch(fp4_quantize_kernel<T, 16, true>);
    if (ue8m0) return launch(fp4_quantize_kernel<T, 32, true>);
    return launch(fp4_quantize_kernel<T, 32, false>);
  });
}

// (add +) RMSNorm + FP4 quantisation, one kernel.  sf2: optional second scale tensor in the other layout.
extern "C" int rmsnorm_fp4quant(void* x, void* residual, void* weight, void* q, void* sf, void* sf2, void* global_scale, int64_t M,
                                int64_t K, int64_t ldx, int64_t ldr, double eps, int64_t vec, int64_t ue8m0, int64_t swizzled,
                                int64_t weight_bias, int64_t dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(vec == 16 || vec == 32, "rmsnorm_fp4quant: block size must be 16 or 32");
  FIB_CHECK(K % vec == 0 && K % 8 == 0 && ldx % 8 == 0 && ldr % 8 == 0, "rmsnorm_fp4quant: hidden must be a multiple of the block size");
  FIB_CHECK(K * 4 <= 200 * 1024, "rmsnorm_fp4quant: hidden size too large for the row buffer");
  if (M == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const size_t smem = size_t(K) * 4;
  LaunchCfg lc(dim3((unsigned)M), dim3(256), smem, stream, pdl != 0);
  return FIB_DISPATCH_HALF(dtype, T, [&]() -> int {
    auto launch = [&](auto kern) -> int {
      if (smem > 48 * 1024) FIB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, (const T*)x, (T*)residual, (const T*)weight, (uint8_t*)q, (uint8_t*)sf,
                                        (uint8_t*)sf2, (const float*)global_scale, M, K, ldx, ldr, (float)eps, (int)swizzled,
                                        (int)weight_bias));
      return 0;
    };
    if (vec == 16 && !ue8m0) return launch(rmsnorm_fp4quant_kernel<T, 16, false>);
    if (vec == 16) return launch(rmsnorm_fp4quant_kernel<T, 16, true>);
    if (ue8m0) return launch(rmsnorm_fp4quant_kernel<T, 32, true>);
    return launch(rmsnorm_fp4quant_kernel<T, 32, false>);
  });
}

// 1 x 128 group fp8 quantisation (DeepSeek activations); see fp8_group_quantize_kernel for the MoE modes.
extern "C" int fp8_group_quantize(void* x, void* q, void* scale, int64_t M, int64_t K, int64_t ldx, int64_t gated, void* row_list,
                                  int64_t n_list, int64_t gather, int64_t list_div, int64_t dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(K % 128 == 0 && ldx % 8 == 0, "fp8_group_quantize: K must be a multiple of 128");
  const int64_t total = (row_list ? n_list : M) * (K / 16);
  if (total == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LaunchCfg lc(dim3(grid_for(total)), dim3(256), 0, stream, pdl != 0);
  return FIB_DISPATCH_HALF(dtype, T, [&]() -> int {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, fp8_group_quantize_kernel<T>, (const T*)x, (__nv_fp8_e4m3*)q, (float*)scale, M, K,
                                      ldx, (int)gated, (const int32_t*)row_list, n_list, (int)gather,
                                      (int)(list_div > 0 ? list_div : 1)));
    return 0;
  });
}

extern "C" int mxfp8_quantize(void* x, void* q, void* sf, int64_t M, int64_t K, int64_t ldx, int64_t swizzled,
                              int64_t dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(K % 32 == 0 && ldx % 8 == 0, "mxfp8_quantize: K must be a multiple of 32");
  if (M * K == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LaunchCfg lc(dim3(grid_for(M * (K / 32))), dim3(256), 0, stream, pdl != 0);
  return FIB_DISPATCH_HALF(dtype, T, [&]() -> int {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, mxfp8_quantize_kernel<T>, (const T*)x, (__nv_fp8_e4m3*)q, (uint8_t*)sf, M,
                                      K, ldx, (int)swizzled));
    return 0;
  });
}

extern "C" int sf_interleave(void* in, void* out, int64_t batch, int64_t M, int64_t Kc, int64_t out_batch_stride,
                             int64_t to_swizzled, int64_t stream_) {
  if (batch * M * Kc == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ++launch_counter();
  sf_interleave_kernel<<<grid_for(batch * M * Kc), 256, 0, stream>>>((const uint8_t*)in, (uint8_t*)out, batch, M, Kc,
                                                                     out_batch_stride, (int)to_swizzled);
  FIB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int fp4_dequantize(void* q, void* sf, void* global_scale, void* out, int64_t M, int64_t K, int64_t vec,
                              int64_t ue8m0, int64_t swizzled, int64_t stream_) {
  if (M * K == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ++launch_counter();
  fp4_dequant_kernel<<<grid_for(M * K / 2), 256, 0, stream>>>((const uint8_t*)q, (const uint8_t*)sf,
                                                              (const float*)global_scale, (float*)out, M, K, (int)vec,
                                                              (int)ue8m0, (int)swizzled);
  FIB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int packbits(void* x, void* y, void* in_indptr, void* out_indptr, int64_t n, int64_t num_segments,
                        int64_t big_endian, int64_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ++launch_counter();
  if (num_segments == 0) {
    if (n == 0) return 0;
    packbits_kernel<<<grid_for((n + 7) / 8), 256, 0, stream>>>((const uint8_t*)x, (uint8_t*)y, nullptr, nullptr, n, 0,
                                                               (int)big_endian);
  } else {
    dim3 grid(64, (unsigned)(num_segments < 1024 ? num_segments : 1024));
    packbits_kernel<<<grid, 256, 0, stream>>>((const uint8_t*)x, (uint8_t*)y, (const int32_t*)in_indptr,
                                              (const int32_t*)out_indptr, n, (int)num_segments, (int)big_endian);
  }
  FIB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
