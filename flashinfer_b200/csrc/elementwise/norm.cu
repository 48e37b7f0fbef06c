// RMSNorm / LayerNorm family for sm_100a (bandwidth-bound SIMT, 16 B vectors, PDL at both ends).
//
// Parity: reference flashinfer/norm/__init__.py:112-551 (rmsnorm, rmsnorm_quant, fused_add_rmsnorm,
// fused_add_rmsnorm_quant, gemma_*, layernorm, fused_rmsnorm_silu) and kernels
// include/flashinfer/norm.cuh:37-745.  One template covers every RMSNorm variant: optional residual
// add (in place), weight bias (Gemma's 1+w), SiLU epilogue, fp8 (e4m3/e5m2) quantised output.
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

template <int kThreadsMax = 1024>
__device__ __forceinline__ float block_sum(float v, float* smem) {
  v = warp_reduce_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = (blockDim.x + 31) >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = (lane < nwarps) ? smem[lane] : 0.f;
  r = warp_reduce_sum(r);
  __syncthreads();
  return r;
}

template <typename OutT>
__device__ __forceinline__ OutT cvt_out(float v);
template <>
__device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_fp8_e4m3 cvt_out<__nv_fp8_e4m3>(float v) {
  return __nv_fp8_e4m3(fminf(fmaxf(v, -448.f), 448.f));
}
template <>
__device__ __forceinline__ __nv_fp8_e5m2 cvt_out<__nv_fp8_e5m2>(float v) {
  return __nv_fp8_e5m2(fminf(fmaxf(v, -57344.f), 57344.f));
}

struct NormParams {
  const void* x;        // input  (also normed output when out == x)
  void* out;
  void* residual;       // in/out (kAdd)
  const void* w;
  const float* scale_ptr;  // optional device scalar (quant): out = y / scale
  int64_t rows, hidden, heads;
  int64_t x_s0, x_s1, o_s0, o_s1, r_s0, r_s1;  // row = (i / heads) * s0 + (i % heads) * s1
  float eps, weight_bias, scale;
};

template <typename T, typename OutT, bool kAdd, bool kSilu>
__global__ void __launch_bounds__(1024) rmsnorm_kernel(const NormParams p) {
  constexpr int VN = 16 / sizeof(T);
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const int64_t r0 = row / p.heads, r1 = row % p.heads;
  const T* x = reinterpret_cast<const T*>(p.x) + r0 * p.x_s0 + r1 * p.x_s1;
  OutT* o = reinterpret_cast<OutT*>(p.out) + r0 * p.o_s0 + r1 * p.o_s1;
  T* res = kAdd ? reinterpret_cast<T*>(p.residual) + r0 * p.r_s0 + r1 * p.r_s1 : nullptr;
  const T* w = reinterpret_cast<const T*>(p.w);
  const int nvec = int(p.hidden / VN);

  ptx::grid_dep_wait();
  ptx::grid_dep_launch();  // early trigger: dependents overlap their prologue, they still wait for our completion

  float cache[VN];
  float ss = 0.f;
  bool cached = false;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    Vec16<T> xv = ld16(x + v * VN);
    float f[VN];
    if constexpr (kAdd) {
      Vec16<T> rv = ld16(res + v * VN);
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const float s = to_f32(xv.v[e]) + to_f32(rv.v[e]);
        rv.v[e] = from_f32<T>(s);
        f[e] = to_f32(rv.v[e]);  // normalise the rounded residual, like the reference
      }
      st16(res + v * VN, rv);
    } else {
#pragma unroll
      for (int e = 0; e < VN; ++e) f[e] = to_f32(xv.v[e]);
    }
#pragma unroll
    for (int e = 0; e < VN; ++e) ss += f[e] * f[e];
    if (v == threadIdx.x) {
#pragma unroll
      for (int e = 0; e < VN; ++e) cache[e] = f[e];
      cached = true;
    }
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / float(p.hidden) + p.eps);
  float qs = p.scale;
  if (p.scale_ptr) qs = 1.f / __ldg(p.scale_ptr);

  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float f[VN];
    if (v == threadIdx.x && cached) {
#pragma unroll
      for (int e = 0; e < VN; ++e) f[e] = cache[e];
    } else {
      Vec16<T> xv = ld16((kAdd ? (const T*)res : x) + v * VN);
#pragma unroll
      for (int e = 0; e < VN; ++e) f[e] = to_f32(xv.v[e]);
    }
    Vec16<T> wv = ldg16(w + v * VN);
    OutT ov[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      float y = f[e] * rstd * (to_f32(wv.v[e]) + p.weight_bias);
      if constexpr (kSilu) y = y / (1.f + __expf(-y));
      ov[e] = cvt_out<OutT>(y * qs);
    }
    if constexpr (sizeof(OutT) == sizeof(T)) {
      *reinterpret_cast<int4*>(o + v * VN) = *reinterpret_cast<const int4*>(ov);
    } else if constexpr (sizeof(OutT) * 2 == sizeof(T)) {
      *reinterpret_cast<int2*>(o + v * VN) = *reinterpret_cast<const int2*>(ov);
    } else {
#pragma unroll
      for (int e = 0; e < VN; ++e) o[v * VN + e] = ov[e];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(1024)
layernorm_kernel(const T* __restrict__ xin, T* __restrict__ out, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int64_t hidden, int64_t x_stride, int64_t o_stride, float eps) {
  __shared__ float red[32];
  const T* x = xin + blockIdx.x * x_stride;
  T* o = out + blockIdx.x * o_stride;
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();  // early trigger: dependents overlap their prologue, they still wait for our completion
  float s = 0.f;
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) s += to_f32(x[i]);
  const float mean = block_sum(s, red) / float(hidden);
  float ss = 0.f;
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
    const float d = to_f32(x[i]) - mean;
    ss += d * d;
  }
  const float rstd = rsqrtf(block_sum(ss, red) / float(hidden) + eps);
  for (int i = threadIdx.x; i < hidden; i += blockDim.x)
    o[i] = from_f32<T>((to_f32(x[i]) - mean) * rstd * gamma[i] + beta[i]);
}

template <typename T, typename OutT>
int launch_rms(const NormParams& p, bool add, bool silu, bool pdl, cudaStream_t stream) {
  constexpr int VN = 16 / sizeof(T);
  int threads = int((p.hidden / VN + 31) / 32 * 32);
  if (threads > 1024) threads = 1024;
  if (threads < 32) threads = 32;
  LaunchCfg lc(dim3((unsigned)p.rows), dim3(threads), 0, stream, pdl);
  if (add && silu) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rmsnorm_kernel<T, OutT, true, true>, p));
  } else if (add) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rmsnorm_kernel<T, OutT, true, false>, p));
  } else if (silu) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rmsnorm_kernel<T, OutT, false, true>, p));
  } else {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rmsnorm_kernel<T, OutT, false, false>, p));
  }
  return 0;
}

}  // namespace

// Generic RMSNorm entry. out_dtype may equal dtype or be e4m3/e5m2 (quantised output, out = y/scale).
extern "C" int rmsnorm_run(void* x, void* out, void* residual, void* w, void* scale_ptr, int64_t rows, int64_t hidden,
                           int64_t heads, int64_t x_s0, int64_t x_s1, int64_t o_s0, int64_t o_s1, int64_t r_s0,
                           int64_t r_s1, double eps, double weight_bias, double scale, int64_t silu, int64_t dtype,
                           int64_t out_dtype, int64_t pdl, int64_t stream_) {
  if (rows == 0) return 0;
  FIB_CHECK(hidden % (16 / dtype_size(dtype)) == 0, "hidden must be a multiple of the 16B vector width");
  NormParams p;
  p.x = x;
  p.out = out;
  p.residual = residual;
  p.w = w;
  p.scale_ptr = (const float*)scale_ptr;
  p.rows = rows;
  p.hidden = hidden;
  p.heads = heads;
  p.x_s0 = x_s0;
  p.x_s1 = x_s1;
  p.o_s0 = o_s0;
  p.o_s1 = o_s1;
  p.r_s0 = r_s0;
  p.r_s1 = r_s1;
  p.eps = (float)eps;
  p.weight_bias = (float)weight_bias;
  p.scale = scale != 0.0 ? (float)(1.0 / scale) : 1.f;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const bool add = residual != nullptr;
  return FIB_DISPATCH_HALF(dtype, T, [&]() -> int {
    if (out_dtype == dtype) return launch_rms<T, T>(p, add, silu != 0, pdl != 0, stream);
    if (out_dtype == kE4M3) return launch_rms<T, __nv_fp8_e4m3>(p, add, silu != 0, pdl != 0, stream);
    if (out_dtype == kE5M2) return launch_rms<T, __nv_fp8_e5m2>(p, add, silu != 0, pdl != 0, stream);
    return set_error("rmsnorm: unsupported output dtype");
  });
}

extern "C" int layernorm_run(void* x, void* out, void* gamma, void* beta, int64_t rows, int64_t hidden, int64_t x_stride,
                             int64_t o_stride, double eps, int64_t dtype, int64_t pdl, int64_t stream_) {
  if (rows == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int threads = int((hidden + 31) / 32 * 32);
  if (threads > 1024) threads = 1024;
  return FIB_DISPATCH_HALF(dtype, T, [&]() -> int {
    LaunchCfg lc(dim3((unsigned)rows), dim3(threads), 0, stream, pdl != 0);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, layernorm_kernel<T>, (const T*)x, (T*)out, (const float*)gamma,
                                      (const float*)beta, hidden, x_stride, o_stride, (float)eps));
    return 0;
  });
}
