// Small runtime services: tensor statistics kernel (API logging level 5, cf. reference
// csrc/api_log_stats.cu:64), L2 flush for benchmarking, device facts.
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

// out[0]=min out[1]=max out[2]=sum out[3]=nan_count out[4]=inf_count
template <typename T>
__global__ void tensor_stats_kernel(const T* __restrict__ x, int64_t n, float* __restrict__ out) {
  float mn = INFINITY, mx = -INFINITY, sum = 0.f;
  int nan = 0, inf = 0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float v = to_f32(x[i]);
    if (isnan(v)) {
      ++nan;
    } else if (isinf(v)) {
      ++inf;
    } else {
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
      sum += v;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    nan += __shfl_xor_sync(0xffffffffu, nan, o);
    inf += __shfl_xor_sync(0xffffffffu, inf, o);
  }
  if ((threadIdx.x & 31) == 0) {
    // float atomics via CAS on the ordered-int trick are overkill: use atomicMin/Max on int views
    atomicMin(reinterpret_cast<int*>(out + 0), mn >= 0 ? __float_as_int(mn) : INT_MAX);  // fixed up below
    atomicAdd(out + 2, sum);
    atomicAdd(out + 3, float(nan));
    atomicAdd(out + 4, float(inf));
    // generic float min/max via CAS loop
    float old = out[5];
    while (mn < old) {
      const int assumed = __float_as_int(old);
      const int prev = atomicCAS(reinterpret_cast<int*>(out + 5), assumed, __float_as_int(mn));
      if (prev == assumed) break;
      old = __int_as_float(prev);
    }
    old = out[1];
    while (mx > old) {
      const int assumed = __float_as_int(old);
      const int prev = atomicCAS(reinterpret_cast<int*>(out + 1), assumed, __float_as_int(mx));
      if (prev == assumed) break;
      old = __int_as_float(prev);
    }
  }
}

__global__ void l2_flush_kernel(int4* __restrict__ buf, int64_t n16) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n16; i += int64_t(gridDim.x) * blockDim.x)
    buf[i] = make_int4(int(i), 0, 0, 0);
}

}  // namespace

// out must be 6 floats initialised by the caller to {+inf(unused), -inf, 0, 0, 0, +inf}; min is returned in out[5].
extern "C" int tensor_stats(void* x, int64_t n, int64_t dtype, void* out, int64_t stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  return FIB_DISPATCH_FLOAT(dtype, T, [&]() -> int {
    tensor_stats_kernel<T><<<blocks, 256, 0, s>>>((const T*)x, n, (float*)out);
    FIB_CUDA_CHECK(cudaGetLastError());
    return 0;
  });
}

extern "C" int l2_flush(void* buf, int64_t bytes, int64_t stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  l2_flush_kernel<<<1184, 256, 0, s>>>((int4*)buf, bytes / 16);
  FIB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int device_facts(int64_t* out) {
  int dev = 0;
  FIB_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  FIB_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  out[0] = prop.multiProcessorCount;
  out[1] = prop.major * 10 + prop.minor;
  out[2] = (int64_t)prop.sharedMemPerBlockOptin;
  out[3] = (int64_t)prop.l2CacheSize;
  out[4] = (int64_t)prop.totalGlobalMem;
  return 0;
}
