// Attention-state merge operators (o, lse base-2).  Parity: reference flashinfer/cascade.py:42-224,
// include/flashinfer/attention/cascade.cuh:45-466 (MergeState, MergeStateInPlace, MergeStates and the
// variable-length merge used by split-KV).  SIMT, 16 B vectors over head_dim, PDL.
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

// v_a/v_b/v_out: [n, H, D] (contiguous), s_*: [n, H] fp32.  in_place: v_out==v_a, s_out==s_a.
template <typename T>
__global__ void __launch_bounds__(256)
merge_state_kernel(const T* __restrict__ v_a, const float* __restrict__ s_a, const T* __restrict__ v_b,
                   const float* __restrict__ s_b, T* __restrict__ v_out, float* __restrict__ s_out,
                   const uint8_t* __restrict__ mask, int64_t n, int H, int D) {
  constexpr int VN = 16 / sizeof(T);
  const int vec = D / VN;
  const int64_t total = n * H * vec;
  ptx::grid_dep_wait();
  for (int64_t w = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; w < total; w += int64_t(gridDim.x) * blockDim.x) {
    const int v = int(w % vec);
    const int64_t nh = w / vec;
    if (mask && !mask[nh / H]) continue;
    const float sa = s_a[nh], sb = s_b[nh];
    float m = fmaxf(sa, sb);
    if (m == -INFINITY) m = 0.f;
    const float wa = exp2f(sa - m), wb = exp2f(sb - m);
    const float den = wa + wb;
    const float inv = den > 0.f ? 1.f / den : 0.f;
    const Vec16<T> a = ld16(v_a + nh * D + v * VN);
    const Vec16<T> b = ld16(v_b + nh * D + v * VN);
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < VN; ++e) o.v[e] = from_f32<T>((to_f32(a.v[e]) * wa + to_f32(b.v[e]) * wb) * inv);
    st16(v_out + nh * D + v * VN, o);
    if (v == 0) s_out[nh] = den > 0.f ? log2f(den) + m : -INFINITY;
  }
  ptx::grid_dep_launch();
}

// v: [n, K, H, D], s: [n, K, H] -> v_out [n, H, D], s_out [n, H]
template <typename T>
__global__ void __launch_bounds__(256)
merge_states_kernel(const T* __restrict__ v, const float* __restrict__ s, T* __restrict__ v_out,
                    float* __restrict__ s_out, int64_t n, int K, int H, int D) {
  constexpr int VN = 16 / sizeof(T);
  const int vec = D / VN;
  const int64_t total = n * H * vec;
  ptx::grid_dep_wait();
  for (int64_t w = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; w < total; w += int64_t(gridDim.x) * blockDim.x) {
    const int vi = int(w % vec);
    const int64_t nh = w / vec;
    const int64_t i = nh / H;
    const int h = int(nh % H);
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, s[(i * K + k) * H + h]);
    if (m == -INFINITY) m = 0.f;
    float acc[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[e] = 0.f;
    float den = 0.f;
    for (int k = 0; k < K; ++k) {
      const float sk = s[(i * K + k) * H + h];
      if (sk == -INFINITY) continue;  // unused split slot: its v may be uninitialised
      const float wk = exp2f(sk - m);
      den += wk;
      const Vec16<T> x = ld16(v + ((i * K + k) * H + h) * int64_t(D) + vi * VN);
#pragma unroll
      for (int e = 0; e < VN; ++e) acc[e] += wk * to_f32(x.v[e]);
    }
    const float inv = den > 0.f ? 1.f / den : 0.f;
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < VN; ++e) o.v[e] = from_f32<T>(acc[e] * inv);
    st16(v_out + nh * D + vi * VN, o);
    if (vi == 0) s_out[nh] = den > 0.f ? log2f(den) + m : -INFINITY;
  }
  ptx::grid_dep_launch();
}

}  // namespace

extern "C" int merge_state(void* v_a, void* s_a, void* v_b, void* s_b, void* v_out, void* s_out, void* mask, int64_t n,
                           int64_t H, int64_t D, int64_t dtype, int64_t pdl, int64_t stream_) {
  if (n == 0) return 0;
  FIB_CHECK(D % (16 / dtype_size(dtype)) == 0, "merge_state: head_dim must be a multiple of the 16B vector width");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  return FIB_DISPATCH_FLOAT(dtype, T, [&]() -> int {
    constexpr int VN = 16 / sizeof(T);
    const int64_t total = n * H * (D / VN);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = int64_t(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    LaunchCfg lc(dim3((unsigned)blocks), dim3(256), 0, stream, pdl != 0);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, merge_state_kernel<T>, (const T*)v_a, (const float*)s_a, (const T*)v_b,
                                      (const float*)s_b, (T*)v_out, (float*)s_out, (const uint8_t*)mask, n, (int)H,
                                      (int)D));
    return 0;
  });
}

extern "C" int merge_states(void* v, void* s, void* v_out, void* s_out, int64_t n, int64_t K, int64_t H, int64_t D,
                            int64_t dtype, int64_t pdl, int64_t stream_) {
  if (n == 0) return 0;
  FIB_CHECK(D % (16 / dtype_size(dtype)) == 0, "merge_states: head_dim must be a multiple of the 16B vector width");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  return FIB_DISPATCH_FLOAT(dtype, T, [&]() -> int {
    constexpr int VN = 16 / sizeof(T);
    const int64_t total = n * H * (D / VN);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = int64_t(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    LaunchCfg lc(dim3((unsigned)blocks), dim3(256), 0, stream, pdl != 0);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, merge_states_kernel<T>, (const T*)v, (const float*)s, (T*)v_out,
                                      (float*)s_out, n, (int)K, (int)H, (int)D));
    return 0;
  });
}
