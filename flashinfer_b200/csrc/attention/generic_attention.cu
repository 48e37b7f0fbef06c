// Generic attention kernel (CUDA cores): the catch-all behind the tcgen05 kernels.
//
// Covers what the specialised decode / prefill kernels do not: head_dim 32..256 (qk and vo may differ), fp8 (e4m3 /
// e5m2) or 16-bit KV with de-quantisation scales, packed custom masks, ALiBi, sliding window, soft-cap, ragged or
// paged KV in NHD / HND, any q_len.  Parity: reference AttentionVariant hooks (include/flashinfer/attention/variants.cuh:
// 31-92: custom mask, sliding window, soft-cap, ALiBi), fp8 KV dequant-on-load (prefill.cuh), head_dim 64 / 256
// instantiations of the fa2 kernels (flashinfer/aot.py:113-211).
//
// One warp per (query row, head): the 32 lanes score 32 keys at a time (each lane one full q.k dot product with
// 16-byte loads), online softmax in the exp2 domain, then the P.V accumulation broadcasts p_j with shuffles while each
// lane owns D/32 output dims (coalesced V reads).  Bandwidth-bound shapes reach a fraction of the TMA kernels; the
// point of this kernel is coverage with native code instead of a framework fallback.
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>
#include <cuda_fp8.h>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

struct GParams {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  float* lse;
  const int32_t* qo_indptr;       // [B+1]
  const int32_t* kv_indptr;       // ragged: [B+1] token offsets; paged: [B+1] page offsets
  const int32_t* kv_indices;      // paged: page ids (null = ragged)
  const int32_t* kv_last_page_len;
  const uint8_t* packed_mask;     // optional, little-endian bits, per request row-major [qo_len, kv_len]
  const int32_t* mask_indptr;     // [B+1] bit offsets
  const float* alibi_slopes;      // [Hq] or null
  int64_t q_sn, q_sh, o_sn, o_sh;
  int64_t k_sp, k_sn, k_sh, v_sp, v_sn, v_sh;  // page / token / head strides (elements); ragged uses sn / sh only
  int B, Hq, Hkv, Dqk, Dvo, page_size, total_q;
  int causal, window_left;
  float sm_scale, soft_cap, k_scale, v_scale;
};

template <typename T>
__device__ __forceinline__ float cvt(T x) { return to_f32(x); }
template <>
__device__ __forceinline__ float cvt<__nv_fp8_e4m3>(__nv_fp8_e4m3 x) { return float(x); }
template <>
__device__ __forceinline__ float cvt<__nv_fp8_e5m2>(__nv_fp8_e5m2 x) { return float(x); }

template <typename TQ, typename TKV>
__global__ void __launch_bounds__(128) generic_attention_kernel(const GParams p) {
  constexpr int kMaxD = 256;
  constexpr int KV_VEC = 16 / sizeof(TKV);
  __shared__ float sq[4][kMaxD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  const int h = blockIdx.y;
  if (row >= p.total_q) return;
  ptx::grid_dep_wait();
  // request of this row
  int lo = 0, hi = p.B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (p.qo_indptr[mid] <= row) lo = mid; else hi = mid;
  }
  const int b = lo;
  const int q0 = p.qo_indptr[b], qo_len = p.qo_indptr[b + 1] - q0, qi = row - q0;
  int kv_len;
  if (p.kv_indices) {
    const int np = p.kv_indptr[b + 1] - p.kv_indptr[b];
    kv_len = np > 0 ? (np - 1) * p.page_size + p.kv_last_page_len[b] : 0;
  } else {
    kv_len = p.kv_indptr[b + 1] - p.kv_indptr[b];
  }
  const int hk = h / (p.Hq / p.Hkv);
  const int q_abs = qi + kv_len - qo_len;  // absolute position of this query
  const TQ* q = reinterpret_cast<const TQ*>(p.q) + int64_t(row) * p.q_sn + int64_t(h) * p.q_sh;
  const float qs = p.soft_cap > 0.f ? p.sm_scale / p.soft_cap : p.sm_scale * 1.4426950408889634f;
  for (int d = lane; d < p.Dqk; d += 32) sq[warp][d] = to_f32(q[d]) * qs * p.k_scale;
  __syncwarp();
  const TKV* kbase = reinterpret_cast<const TKV*>(p.k);
  const TKV* vbase = reinterpret_cast<const TKV*>(p.v);
  auto kv_off = [&](int pos, int64_t sp, int64_t sn, int64_t sh) -> int64_t {
    if (p.kv_indices) {
      const int page = p.kv_indices[p.kv_indptr[b] + pos / p.page_size];
      return int64_t(page) * sp + int64_t(pos % p.page_size) * sn + int64_t(hk) * sh;
    }
    return int64_t(p.kv_indptr[b] + pos) * sn + int64_t(hk) * sh;
  };
  int j_end = kv_len;
  if (p.causal) j_end = min(kv_len, q_abs + 1);
  int j_begin = 0;
  if (p.window_left >= 0) j_begin = max(0, q_abs - p.window_left);
  const float slope = p.alibi_slopes ? p.alibi_slopes[h] * 1.4426950408889634f : 0.f;
  const int64_t mbase = p.packed_mask ? int64_t(p.mask_indptr[b]) + int64_t(qi) * kv_len : 0;
  constexpr int kPer = kMaxD / 32;
  float acc[kPer];
#pragma unroll
  for (int i = 0; i < kPer; ++i) acc[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int j0 = j_begin; j0 < j_end; j0 += 32) {
    const int j = j0 + lane;
    float s = -INFINITY;
    bool ok = j < j_end;
    if (ok && p.packed_mask) {
      const int64_t bit = mbase + j;
      ok = (p.packed_mask[bit >> 3] >> (bit & 7)) & 1;
    }
    if (ok) {
      const TKV* kr = kbase + kv_off(j, p.k_sp, p.k_sn, p.k_sh);
      float dot = 0.f;
      for (int d = 0; d < p.Dqk; d += KV_VEC) {
        const int4 raw = *reinterpret_cast<const int4*>(kr + d);
        const TKV* kk = reinterpret_cast<const TKV*>(&raw);
#pragma unroll
        for (int e = 0; e < KV_VEC; ++e) dot += sq[warp][d + e] * cvt<TKV>(kk[e]);
      }
      if (p.soft_cap > 0.f) dot = p.soft_cap * 1.4426950408889634f * tanhf(dot);
      s = dot + slope * float(j - q_abs);
    }
    float mx = s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float m_new = fmaxf(m, mx);
    if (m_new == -INFINITY) continue;  // whole chunk masked so far
    const float corr = exp2f(m - m_new);
    const float pj = ok ? exp2f(s - m_new) : 0.f;
    float ps = pj;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
    l = l * corr + ps;
#pragma unroll
    for (int i = 0; i < kPer; ++i) acc[i] *= corr;
    m = m_new;
    const int cnt = min(32, j_end - j0);
    for (int t = 0; t < cnt; ++t) {
      const float pt = __shfl_sync(0xffffffffu, pj, t);
      if (pt == 0.f) continue;
      const TKV* vr = vbase + kv_off(j0 + t, p.v_sp, p.v_sn, p.v_sh);
#pragma unroll
      for (int i = 0; i < kPer; ++i) {
        const int d = lane + 32 * i;
        if (d < p.Dvo) acc[i] += pt * cvt<TKV>(vr[d]);
      }
    }
  }
  TQ* o = reinterpret_cast<TQ*>(p.o) + int64_t(row) * p.o_sn + int64_t(h) * p.o_sh;
  const float inv = l > 0.f ? p.v_scale / l : 0.f;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int d = lane + 32 * i;
    if (d < p.Dvo) o[d] = from_f32<TQ>(acc[i] * inv);
  }
  if (p.lse && lane == 0) p.lse[int64_t(row) * p.Hq + h] = l > 0.f ? m + log2f(l) : -INFINITY;
  ptx::grid_dep_launch();
}

template <typename TQ>
int launch_kv(const GParams& p, int64_t kv_dtype, cudaStream_t stream, bool pdl) {
  LaunchCfg lc(dim3((unsigned)((p.total_q + 3) / 4), (unsigned)p.Hq), dim3(128), 0, stream, pdl);
  if (kv_dtype == kF16) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, generic_attention_kernel<TQ, __half>, p));
  } else if (kv_dtype == kBF16) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, generic_attention_kernel<TQ, __nv_bfloat16>, p));
  } else if (kv_dtype == kE4M3) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, generic_attention_kernel<TQ, __nv_fp8_e4m3>, p));
  } else if (kv_dtype == kE5M2) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, generic_attention_kernel<TQ, __nv_fp8_e5m2>, p));
  } else {
    FIB_CHECK(false, "generic_attention: kv dtype must be f16 / bf16 / e4m3 / e5m2");
  }
  return 0;
}

}  // namespace

// ints: int64[16] = {q_sn,q_sh,o_sn,o_sh,k_sp,k_sn,k_sh,v_sp,v_sn,v_sh, 0...}
extern "C" int generic_attention(void* q, void* k, void* v, void* o, void* lse, void* qo_indptr, void* kv_indptr, void* kv_indices,
                                 void* kv_last_page_len, void* packed_mask, void* mask_indptr, void* alibi_slopes,
                                 void* strides_host, int64_t B, int64_t total_q, int64_t Hq, int64_t Hkv, int64_t Dqk, int64_t Dvo,
                                 int64_t page_size, int64_t causal, int64_t window_left, double sm_scale, double soft_cap,
                                 double k_scale, double v_scale, int64_t q_dtype, int64_t kv_dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(Dqk <= 256 && Dvo <= 256 && Dqk % 16 == 0, "generic_attention: head dims must be <= 256 and qk dim a multiple of 16");
  FIB_CHECK(Hq % Hkv == 0, "generic_attention: num_qo_heads must be a multiple of num_kv_heads");
  if (total_q == 0 || B == 0) return 0;
  const int64_t* s = (const int64_t*)strides_host;
  GParams p;
  p.q = q; p.k = k; p.v = v; p.o = o; p.lse = (float*)lse;
  p.qo_indptr = (const int32_t*)qo_indptr; p.kv_indptr = (const int32_t*)kv_indptr; p.kv_indices = (const int32_t*)kv_indices;
  p.kv_last_page_len = (const int32_t*)kv_last_page_len; p.packed_mask = (const uint8_t*)packed_mask;
  p.mask_indptr = (const int32_t*)mask_indptr; p.alibi_slopes = (const float*)alibi_slopes;
  p.q_sn = s[0]; p.q_sh = s[1]; p.o_sn = s[2]; p.o_sh = s[3];
  p.k_sp = s[4]; p.k_sn = s[5]; p.k_sh = s[6]; p.v_sp = s[7]; p.v_sn = s[8]; p.v_sh = s[9];
  p.B = (int)B; p.Hq = (int)Hq; p.Hkv = (int)Hkv; p.Dqk = (int)Dqk; p.Dvo = (int)Dvo; p.page_size = (int)page_size;
  p.total_q = (int)total_q; p.causal = (int)causal; p.window_left = (int)window_left;
  p.sm_scale = (float)sm_scale; p.soft_cap = (float)soft_cap; p.k_scale = (float)k_scale; p.v_scale = (float)v_scale;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (q_dtype == kF16) return launch_kv<__half>(p, kv_dtype, stream, pdl != 0);
  if (q_dtype == kBF16) return launch_kv<__nv_bfloat16>(p, kv_dtype, stream, pdl != 0);
  FIB_CHECK(false, "generic_attention: q dtype must be f16 / bf16");
  return 1;
}
