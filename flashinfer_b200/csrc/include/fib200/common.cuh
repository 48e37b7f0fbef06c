// fib200/common.cuh — host/device utilities shared by every native module.
//
// * uniform C ABI error channel (thread-local message + int return code)
// * dtype codes shared with python (flashinfer_b200/utils/dtypes.py)
// * PDL-aware launch helper (cudaLaunchKernelEx + programmatic stream serialization)
// * host-side TMA tensor-map builder through the driver entry point (no -lcuda link)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

namespace fib200 {

// ------------------------------------------------------------------ errors
inline std::string& last_error_storage() {
  static thread_local std::string s;
  return s;
}
inline int set_error(const std::string& msg) {
  last_error_storage() = msg;
  return 1;
}

#define FIB_CUDA_CHECK(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      return ::fib200::set_error(std::string(#expr) + " -> " + cudaGetErrorString(_e) + " at " +   \
                                 __FILE__ + ":" + std::to_string(__LINE__));                       \
    }                                                                                              \
  } while (0)

#define FIB_CHECK(cond, msg)                                                                       \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      return ::fib200::set_error(std::string("check failed: ") + #cond + " : " + (msg) + " at " +  \
                                 __FILE__ + ":" + std::to_string(__LINE__));                       \
    }                                                                                              \
  } while (0)

// per-translation-unit (internal linkage): every native module counts only its own launches
static long long g_launch_counter = 0;
static inline long long& launch_counter() { return g_launch_counter; }

#define FIB_EXPORT_LAST_ERROR()                                                                    \
  extern "C" const char* fib200_last_error() { return ::fib200::last_error_storage().c_str(); }    \
  extern "C" long long fib200_launch_count() { return ::fib200::launch_counter(); }

// 64-bit integer division is ~100 instructions on the GPU; flat-index kernels decompose (row, column) with 32-bit math
// whenever both operands fit (always, except for > 4G-element tensors).
__host__ __device__ __forceinline__ void fast_divmod(int64_t i, int64_t d, int64_t& q, int64_t& r) {
  if ((uint64_t(i) | uint64_t(d)) <= 0xffffffffull) {
    const uint32_t qq = uint32_t(i) / uint32_t(d);
    q = qq;
    r = uint32_t(i) - qq * uint32_t(d);
  } else {
    q = i / d;
    r = i - q * d;
  }
}

// ------------------------------------------------------------------ dtypes
enum DType : int64_t {
  kF16 = 0,
  kBF16 = 1,
  kF32 = 2,
  kE4M3 = 3,
  kE5M2 = 4,
  kU8 = 5,
  kI32 = 6,
  kI64 = 7,
};

__host__ __device__ inline int dtype_size(int64_t dt) {
  switch (dt) {
    case kF16:
    case kBF16:
      return 2;
    case kF32:
    case kI32:
      return 4;
    case kI64:
      return 8;
    default:
      return 1;
  }
}

#define FIB_DISPATCH_HALF(dt, T, ...)                                           \
  [&]() -> int {                                                                \
    if ((dt) == ::fib200::kF16) {                                               \
      using T = __half;                                                         \
      return __VA_ARGS__();                                                     \
    } else if ((dt) == ::fib200::kBF16) {                                       \
      using T = __nv_bfloat16;                                                  \
      return __VA_ARGS__();                                                     \
    }                                                                           \
    return ::fib200::set_error("unsupported dtype (expected f16/bf16)");        \
  }()

#define FIB_DISPATCH_FLOAT(dt, T, ...)                                          \
  [&]() -> int {                                                                \
    if ((dt) == ::fib200::kF16) {                                               \
      using T = __half;                                                         \
      return __VA_ARGS__();                                                     \
    } else if ((dt) == ::fib200::kBF16) {                                       \
      using T = __nv_bfloat16;                                                  \
      return __VA_ARGS__();                                                     \
    } else if ((dt) == ::fib200::kF32) {                                        \
      using T = float;                                                          \
      return __VA_ARGS__();                                                     \
    }                                                                           \
    return ::fib200::set_error("unsupported dtype (expected f16/bf16/f32)");    \
  }()

// ------------------------------------------------------------------ conversions
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector of T.
template <typename T>
struct alignas(16) Vec16 {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

template <typename T>
__device__ __forceinline__ Vec16<T> ldg16(const T* p) {
  Vec16<T> r;
  *reinterpret_cast<int4*>(&r) = __ldg(reinterpret_cast<const int4*>(p));
  return r;
}
template <typename T>
__device__ __forceinline__ Vec16<T> ld16(const T* p) {
  Vec16<T> r;
  *reinterpret_cast<int4*>(&r) = *reinterpret_cast<const int4*>(p);
  return r;
}
template <typename T>
__device__ __forceinline__ void st16(T* p, const Vec16<T>& r) {
  *reinterpret_cast<int4*>(p) = *reinterpret_cast<const int4*>(&r);
}

__device__ __forceinline__ float warp_reduce_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_reduce_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------ launch helper
inline int num_sms() {
  static int n = [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

struct LaunchCfg {
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attrs[3];
  LaunchCfg(dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, int cluster_x = 1) {
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    int n = 0;
    if (pdl) {
      attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attrs[n].val.programmaticStreamSerializationAllowed = 1;
      ++n;
    }
    if (cluster_x > 1) {
      attrs[n].id = cudaLaunchAttributeClusterDimension;
      attrs[n].val.clusterDim.x = cluster_x;
      attrs[n].val.clusterDim.y = 1;
      attrs[n].val.clusterDim.z = 1;
      ++n;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = n;
    ++launch_counter();
  }
};

// ------------------------------------------------------------------ TMA descriptor (host)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
#if CUDART_VERSION >= 12050
    cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
#else
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
#endif
    return reinterpret_cast<PFN_encodeTiled>(p);
  }();
  return fn;
}

// Build a tiled tensor map. dims/strides are innermost-first; strides[i] is the byte stride of
// dim i+1 (dim 0 is contiguous). Returns 0 on success.
inline int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                     CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return set_error("cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]",
             (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
             (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0), bx[0],
             rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
    return set_error(buf);
  }
  return 0;
}

inline CUtensorMapDataType tmap_dtype(int64_t dt) {
  switch (dt) {
    case kF16:
      return CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    case kBF16:
      return CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    case kF32:
      return CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    default:
      return CU_TENSOR_MAP_DATA_TYPE_UINT8;
  }
}

}  // namespace fib200
