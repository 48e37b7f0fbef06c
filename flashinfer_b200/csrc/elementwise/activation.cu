// Gated activations: out[..., d] = act(x[..., :d]) * x[..., d:]   (silu / gelu / gelu_tanh)
// Parity: reference flashinfer/activation.py:77-202, include/flashinfer/activation.cuh:29.
// Bandwidth-bound: 16 B vector ld/st, flat grid over (rows x d/8) vectors, PDL.
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

enum Act { kSilu = 0, kGelu = 1, kGeluTanh = 2 };

template <int ACT>
__device__ __forceinline__ float act_fn(float x) {
  if constexpr (ACT == kSilu) {
    return x / (1.f + __expf(-x));
  } else if constexpr (ACT == kGelu) {
    return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  } else {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    return 0.5f * x * (1.f + tanhf(k0 * (x + k1 * x * x * x)));
  }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(256)
act_and_mul_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t rows, int64_t d, int64_t in_stride,
                   int64_t out_stride, int gate_second, const int32_t* __restrict__ row_map, int row_list) {
  constexpr int VN = 16 / sizeof(T);
  const int64_t vec_per_row = d / VN;
  const int64_t total = rows * vec_per_row;
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();  // early trigger: dependents overlap their prologue, they still wait for our completion
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int64_t r, c;
    fast_divmod(i, vec_per_row, r, c);
    c *= VN;
    if (row_list) {
      // row_map lists the live rows (expanded -> permuted row, -1 = not local): only those are visited
      r = row_map[r];
      if (r < 0) continue;
    } else if (row_map && row_map[r] < 0) {
      continue;  // MoE padding row: nothing reads its result
    }
    const Vec16<T> a = ld16(in + r * in_stride + (gate_second ? d : 0) + c);   // activated half
    const Vec16<T> b = ld16(in + r * in_stride + (gate_second ? 0 : d) + c);   // linear half
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < VN; ++e) o.v[e] = from_f32<T>(act_fn<ACT>(to_f32(a.v[e])) * to_f32(b.v[e]));
    st16(out + r * out_stride + c, o);
  }
}

}  // namespace

extern "C" int act_and_mul(void* in, void* out, int64_t rows, int64_t d, int64_t in_stride, int64_t out_stride,
                           int64_t act, int64_t gate_second, void* row_map, int64_t row_list, int64_t dtype, int64_t pdl,
                           int64_t stream_) {
  if (rows == 0 || d == 0) return 0;
  FIB_CHECK(d % (16 / dtype_size(dtype)) == 0, "act_and_mul: d must be a multiple of the 16B vector width");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  return FIB_DISPATCH_FLOAT(dtype, T, [&]() -> int {
    constexpr int VN = 16 / sizeof(T);
    const int64_t total = rows * (d / VN);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = int64_t(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    LaunchCfg lc(dim3((unsigned)blocks), dim3(256), 0, stream, pdl != 0);
    if (act == kSilu) {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, act_and_mul_kernel<T, kSilu>, (const T*)in, (T*)out, rows, d, in_stride,
                                        out_stride, (int)gate_second, (const int32_t*)row_map, (int)row_list));
    } else if (act == kGelu) {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, act_and_mul_kernel<T, kGelu>, (const T*)in, (T*)out, rows, d, in_stride,
                                        out_stride, (int)gate_second, (const int32_t*)row_map, (int)row_list));
    } else {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, act_and_mul_kernel<T, kGeluTanh>, (const T*)in, (T*)out, rows, d,
                                        in_stride, out_stride, (int)gate_second, (const int32_t*)row_map, (int)row_list));
    }
    return 0;
  });
}
