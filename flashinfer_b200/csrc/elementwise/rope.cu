// Rotary position embedding family (bandwidth-bound SIMT, flat work-item grid, PDL).
//
// Parity: reference flashinfer/rope.py:433-1691 and include/flashinfer/pos_enc.cuh:294-808:
// apply_rope(_inplace) (indptr+offsets), apply_rope_pos_ids, llama3.1 frequency scaling,
// cos/sin-cache variant (vLLM/SGL compatible), rope + fp8 quantisation (rope_quantize_fp8 /
// mla_rope_quantize_fp8) and the fused RoPE + fp8 quant + paged-KV append.
// One kernel template covers all of them: position source (pos_ids | indptr+offsets), frequency
// source (on-the-fly | cos_sin_cache), layout (NeoX halves | interleaved pairs), output dtype.
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

struct RopeParams {
  const void* q;
  const void* k;
  void* q_out;
  void* k_out;
  const int32_t* pos_ids;   // [nnz] or null
  const int64_t* pos_ids64; // alternative int64 positions
  const int32_t* indptr;    // [B+1] (with offsets) or null
  const int32_t* offsets;   // [B]
  const float* cos_sin_cache;  // [max_pos, rotary_dim] = cos | sin halves, or null
  int64_t nnz, batch;
  int num_q_heads, num_k_heads, head_dim, rotary_dim, interleave;
  int64_t q_sn, q_sh, k_sn, k_sh, qo_sn, qo_sh, ko_sn, ko_sh;
  float rope_rcp_scale, rope_theta_log2;  // freq_i = rcp_scale * 2^(-log2(theta) * 2i/rd)
  int llama31;
  float smooth_a, smooth_b;
  float q_out_scale, k_out_scale;  // quantisation multipliers (1 for plain)
  // fused paged-KV append (optional): roped K and raw V go straight into the cache pages
  const void* v;                  // [nnz, Hk, D] (strides v_sn, v_sh) or null
  void* k_cache;                  // paged destination of K (null = k_out)
  void* v_cache;
  const int32_t* batch_indices;   // [nnz] request of every token
  const int32_t* kv_indices;      // page ids
  const int32_t* kv_indptr;       // [B+1]
  int64_t v_sn, v_sh, c_sp, c_sn, c_sh;  // cache strides: page / in-page token / head (elements)
  int page_size;
  float v_out_scale;
};

__device__ __forceinline__ float rope_freq(const RopeParams& p, int i /* pair index */) {
  float inv = exp2f(-p.rope_theta_log2 * float(2 * i) / float(p.rotary_dim));
  if (p.llama31) {
    float smooth = inv * p.smooth_a + p.smooth_b;
    smooth = fminf(fmaxf(smooth, 0.f), 1.f);
    inv = (1.f - smooth) * (inv * p.rope_rcp_scale) + smooth * inv;
  } else {
    inv *= p.rope_rcp_scale;
  }
  return inv;
}

template <typename OutT>
__device__ __forceinline__ OutT rope_cvt(float v);
template <>
__device__ __forceinline__ __half rope_cvt<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 rope_cvt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __nv_fp8_e4m3 rope_cvt<__nv_fp8_e4m3>(float v) {
  return __nv_fp8_e4m3(fminf(fmaxf(v, -448.f), 448.f));
}
template <>
__device__ __forceinline__ __nv_fp8_e5m2 rope_cvt<__nv_fp8_e5m2>(float v) {
  return __nv_fp8_e5m2(fminf(fmaxf(v, -57344.f), 57344.f));
}

// work item = (token, head, 8-wide chunk index c); chunks [0, rd/16) are rotary pair-chunks for the
// NeoX layout (elements c*8.. and rd/2 + c*8..), [0, rd/8) for interleaved; the rest are pass-through.
template <typename T, typename OutT>
__global__ void __launch_bounds__(256) rope_kernel(const RopeParams p) {
  constexpr int VN = 8;
  const int rd = p.rotary_dim, D = p.head_dim;
  const int rot_items = p.interleave ? rd / VN : rd / (2 * VN);
  const int pass_items = (D - rd) / VN;
  const int items_per_row = rot_items + pass_items;
  const int heads = p.num_q_heads + p.num_k_heads + (p.v ? p.num_k_heads : 0);
  const int64_t total = p.nnz * heads * items_per_row;
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();  // early trigger: dependents overlap their prologue, they still wait for our completion
  for (int64_t w = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; w < total; w += int64_t(gridDim.x) * blockDim.x) {
    const int item = int(w % items_per_row);
    const int64_t rest = w / items_per_row;
    const int head = int(rest % heads);
    const int64_t tok = rest / heads;
    const bool is_q = head < p.num_q_heads;
    const bool is_v = head >= p.num_q_heads + p.num_k_heads;
    const int h = is_q ? head : (is_v ? head - p.num_q_heads - p.num_k_heads : head - p.num_q_heads);
    const T* src = is_v ? reinterpret_cast<const T*>(p.v) + tok * p.v_sn + h * p.v_sh
                        : reinterpret_cast<const T*>(is_q ? p.q : p.k) + tok * (is_q ? p.q_sn : p.k_sn) + h * (is_q ? p.q_sh : p.k_sh);
    OutT* dst;
    if (!is_q && p.k_cache) {
      // paged destination: the position of the token inside its request selects page + slot
      const int b = p.batch_indices[tok];
      const int ppos = p.pos_ids ? p.pos_ids[tok] : int(p.pos_ids64[tok]);
      const int page = p.kv_indices[p.kv_indptr[b] + ppos / p.page_size];
      dst = reinterpret_cast<OutT*>(is_v ? p.v_cache : p.k_cache) + int64_t(page) * p.c_sp + int64_t(ppos % p.page_size) * p.c_sn +
            int64_t(h) * p.c_sh;
    } else {
      dst = reinterpret_cast<OutT*>(is_q ? p.q_out : p.k_out) + tok * (is_q ? p.qo_sn : p.ko_sn) + h * (is_q ? p.qo_sh : p.ko_sh);
    }
    const float oscale = is_q ? p.q_out_scale : (is_v ? p.v_out_scale : p.k_out_scale);
    if (is_v) {
      // V is a plain (optionally scaled / re-typed) copy: same element pattern as the rotary items
      if (item >= rot_items) {
        const int c = rd + (item - rot_items) * VN;
#pragma unroll
        for (int e = 0; e < VN; ++e) dst[c + e] = rope_cvt<OutT>(to_f32(src[c + e]) * oscale);
      } else if (p.interleave) {
        const int c = item * VN;
#pragma unroll
        for (int e = 0; e < VN; ++e) dst[c + e] = rope_cvt<OutT>(to_f32(src[c + e]) * oscale);
      } else {
        const int c = item * VN;
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          dst[c + e] = rope_cvt<OutT>(to_f32(src[c + e]) * oscale);
          dst[rd / 2 + c + e] = rope_cvt<OutT>(to_f32(src[rd / 2 + c + e]) * oscale);
        }
      }
      continue;
    }
    if (item >= rot_items) {
      const int c = rd + (item - rot_items) * VN;
      if ((const void*)(src + c) != (const void*)(dst + c) || oscale != 1.f) {
#pragma unroll
        for (int e = 0; e < VN; ++e) dst[c + e] = rope_cvt<OutT>(to_f32(src[c + e]) * oscale);
      }
      continue;
    }
    // position
    float pos;
    if (p.pos_ids) {
      pos = float(p.pos_ids[tok]);
    } else if (p.pos_ids64) {
      pos = float(p.pos_ids64[tok]);
    } else {
      int lo = 0, hi = int(p.batch);
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (p.indptr[mid] <= tok) lo = mid; else hi = mid;
      }
      pos = float(p.offsets[lo] + int(tok - p.indptr[lo]));
    }
    float x1[VN], x2[VN];
    if (p.interleave) {
      // 8 consecutive elements = 4 pairs
      const int c = item * VN;
      float cs[VN / 2], sn[VN / 2];
#pragma unroll
      for (int e = 0; e < VN / 2; ++e) {
        const int pi = c / 2 + e;
        if (p.cos_sin_cache) {
          const float* row = p.cos_sin_cache + int64_t(pos) * rd;
          cs[e] = row[pi];
          sn[e] = row[rd / 2 + pi];
        } else {
          sincosf(pos * rope_freq(p, pi), &sn[e], &cs[e]);
        }
      }
      float v[VN];
#pragma unroll
      for (int e = 0; e < VN; ++e) v[e] = to_f32(src[c + e]);
#pragma unroll
      for (int e = 0; e < VN / 2; ++e) {
        const float a = v[2 * e], b = v[2 * e + 1];
        dst[c + 2 * e] = rope_cvt<OutT>((a * cs[e] - b * sn[e]) * oscale);
        dst[c + 2 * e + 1] = rope_cvt<OutT>((b * cs[e] + a * sn[e]) * oscale);
      }
    } else {
      const int c = item * VN;
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        x1[e] = to_f32(src[c + e]);
        x2[e] = to_f32(src[rd / 2 + c + e]);
      }
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        float cs, sn;
        if (p.cos_sin_cache) {
          const float* row = p.cos_sin_cache + int64_t(pos) * rd;
          cs = row[c + e];
          sn = row[rd / 2 + c + e];
        } else {
          sincosf(pos * rope_freq(p, c + e), &sn, &cs);
        }
        dst[c + e] = rope_cvt<OutT>((x1[e] * cs - x2[e] * sn) * oscale);
        dst[rd / 2 + c + e] = rope_cvt<OutT>((x2[e] * cs + x1[e] * sn) * oscale);
      }
    }
  }
}

}  // namespace

namespace {
struct AppendArgs {
  const void* v = nullptr;
  void* k_cache = nullptr;
  void* v_cache = nullptr;
  const int32_t* batch_indices = nullptr;
  const int32_t* kv_indices = nullptr;
  const int32_t* kv_indptr = nullptr;
  int64_t v_sn = 0, v_sh = 0, c_sp = 0, c_sn = 0, c_sh = 0;
  int page_size = 1;
  float v_out_scale = 1.f;
};
thread_local AppendArgs g_append;  // consumed (and cleared) by the next rope_run call of this thread
}  // namespace

// Arms the fused paged-KV append for the NEXT rope_run call: K (after RoPE) and V are written into the cache pages
// instead of k_out.  Parity: reference rope_quantize_fp8_append_paged_kv_cache (flashinfer/rope.py:1500-1691).
extern "C" int rope_set_append(void* v, void* k_cache, void* v_cache, void* batch_indices, void* kv_indices, void* kv_indptr,
                               int64_t v_sn, int64_t v_sh, int64_t c_sp, int64_t c_sn, int64_t c_sh, int64_t page_size,
                               double v_out_scale) {
  g_append.v = v;
  g_append.k_cache = k_cache;
  g_append.v_cache = v_cache;
  g_append.batch_indices = (const int32_t*)batch_indices;
  g_append.kv_indices = (const int32_t*)kv_indices;
  g_append.kv_indptr = (const int32_t*)kv_indptr;
  g_append.v_sn = v_sn; g_append.v_sh = v_sh; g_append.c_sp = c_sp; g_append.c_sn = c_sn; g_append.c_sh = c_sh;
  g_append.page_size = (int)page_size;
  g_append.v_out_scale = (float)v_out_scale;
  return 0;
}

extern "C" int rope_run(void* q, void* k, void* q_out, void* k_out, void* pos_ids, int64_t pos_is_i64, void* indptr,
                        void* offsets, void* cos_sin_cache, int64_t nnz, int64_t batch, int64_t num_q_heads,
                        int64_t num_k_heads, int64_t head_dim, int64_t rotary_dim, int64_t interleave, int64_t q_sn,
                        int64_t q_sh, int64_t k_sn, int64_t k_sh, int64_t qo_sn, int64_t qo_sh, int64_t ko_sn,
                        int64_t ko_sh, double rope_scale, double rope_theta, int64_t llama31, double low_freq_factor,
                        double high_freq_factor, double old_context_len, double q_out_scale, double k_out_scale,
                        int64_t dtype, int64_t out_dtype, int64_t pdl, int64_t stream_) {
  if (nnz == 0) return 0;
  FIB_CHECK(rotary_dim % 16 == 0 && head_dim % 8 == 0 && rotary_dim <= head_dim, "rope: bad rotary_dim/head_dim");
  RopeParams p;
  p.q = q;
  p.k = k;
  p.q_out = q_out;
  p.k_out = k_out;
  p.pos_ids = pos_is_i64 ? nullptr : (const int32_t*)pos_ids;
  p.pos_ids64 = pos_is_i64 ? (const int64_t*)pos_ids : nullptr;
  p.indptr = (const int32_t*)indptr;
  p.offsets = (const int32_t*)offsets;
  p.cos_sin_cache = (const float*)cos_sin_cache;
  p.nnz = nnz;
  p.batch = batch;
  p.num_q_heads = (int)num_q_heads;
  p.num_k_heads = k ? (int)num_k_heads : 0;
  p.head_dim = (int)head_dim;
  p.rotary_dim = (int)rotary_dim;
  p.interleave = (int)interleave;
  p.q_sn = q_sn; p.q_sh = q_sh; p.k_sn = k_sn; p.k_sh = k_sh;
  p.qo_sn = qo_sn; p.qo_sh = qo_sh; p.ko_sn = ko_sn; p.ko_sh = ko_sh;
  p.rope_rcp_scale = (float)(1.0 / rope_scale);
  p.rope_theta_log2 = (float)log2(rope_theta);
  p.llama31 = (int)llama31;
  p.smooth_a = 0.f;
  p.smooth_b = 0.f;
  if (llama31) {
    p.smooth_a = (float)(old_context_len / (2.0 * M_PI * (high_freq_factor - low_freq_factor)));
    p.smooth_b = (float)(-1.0 / (high_freq_factor / low_freq_factor - 1.0));
  }
  p.q_out_scale = (float)q_out_scale;
  p.k_out_scale = (float)k_out_scale;
  {
    const AppendArgs a = g_append;
    g_append = AppendArgs();
    p.v = a.v; p.k_cache = a.k_cache; p.v_cache = a.v_cache; p.batch_indices = a.batch_indices; p.kv_indices = a.kv_indices;
    p.kv_indptr = a.kv_indptr; p.v_sn = a.v_sn; p.v_sh = a.v_sh; p.c_sp = a.c_sp; p.c_sn = a.c_sn; p.c_sh = a.c_sh;
    p.page_size = a.page_size; p.v_out_scale = a.v_out_scale;
    if (p.k_cache) FIB_CHECK(pos_ids != nullptr && p.batch_indices && p.kv_indices && p.kv_indptr, "rope+append needs pos_ids / batch_indices / page table");
  }
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int rot_items = interleave ? rotary_dim / 8 : rotary_dim / 16;
  const int64_t total = nnz * (p.num_q_heads + p.num_k_heads + (p.v ? p.num_k_heads : 0)) * (rot_items + (head_dim - rotary_dim) / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = int64_t(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  LaunchCfg lc(dim3((unsigned)blocks), dim3(256), 0, stream, pdl != 0);
  return FIB_DISPATCH_HALF(dtype, T, [&]() -> int {
    if (out_dtype == dtype) {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rope_kernel<T, T>, p));
    } else if (out_dtype == kE4M3) {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rope_kernel<T, __nv_fp8_e4m3>, p));
    } else if (out_dtype == kE5M2) {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, rope_kernel<T, __nv_fp8_e5m2>, p));
    } else {
      return set_error("rope: unsupported output dtype");
    }
    return 0;
  });
}
