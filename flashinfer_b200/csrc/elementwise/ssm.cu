// State-space / linear-attention recurrent kernels and the MLA K concat:
//   selective_state_update  (Mamba / Mamba-2 generation step, single- and multi-token)
//   gated_delta_rule_step   (Gated DeltaNet recurrence: decode, multi-token verify and sequential prefill)
//   concat_mla_k            (k = [k_nope | broadcast k_rope])
//
// Parity: reference flashinfer/mamba/selective_state_update.py:104 (kernel_selective_state_update_*.cuh),
// flashinfer/gdn_decode.py:118-742 / gdn_prefill.py:100 (gated delta rule), flashinfer/concat_ops.py:33
// (include/flashinfer/concat_mla.cuh:141).  All three are HBM-bound streaming updates of a resident state:
// one read + one write of the state per token, fp32 math in registers, 16-byte accesses where the layout allows.
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(__expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

struct SsuParams {
  void* state;        // [cache, H, dim, dstate]
  const void* x;      // [B, T, H, dim]
  const void* dt;     // same logical shape (strides may be 0 on dim)
  const float* A;     // [H, dim, dstate] with strides
  const void* Bm;     // [B, T, G, dstate]
  const void* Cm;
  const float* D;     // [H, dim] or null
  const void* z;      // like x or null
  const float* dt_bias;  // [H, dim] or null
  void* out;          // like x (contiguous [B, T, H, dim])
  const int32_t* state_idx;   // [B] or null
  const int32_t* dst_idx;     // [B] or null
  int64_t x_sb, x_st, x_sh;                  // element strides of x / z / out (dim stride 1)
  int64_t dt_sb, dt_st, dt_sh, dt_sd;        // dt strides
  int64_t A_sh, A_sd, A_sn;
  int64_t B_sb, B_st, B_sg;                  // B / C strides (dstate stride 1)
  int64_t D_sh, D_sd, dtb_sh, dtb_sd;
  int B, T, H, dim, dstate, G, pad_slot, softplus, update_state;
};

// one warp per (b, h, dim row); lanes stride over dstate
template <typename TX, typename TS>
__global__ void __launch_bounds__(256) ssu_kernel(const SsuParams p) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int total = p.B * p.H * p.dim;
  if (row >= total) return;
  ptx::grid_dep_wait();
  const int d = row % p.dim, h = (row / p.dim) % p.H, b = row / (p.dim * p.H);
  const int g = h / (p.H / p.G);
  int slot = p.state_idx ? p.state_idx[b] : b;
  if (slot == p.pad_slot) return;
  const int dslot = p.dst_idx ? p.dst_idx[b] : slot;
  TS* st = reinterpret_cast<TS*>(p.state) + ((int64_t(slot) * p.H + h) * p.dim + d) * p.dstate;
  TS* st_out = reinterpret_cast<TS*>(p.state) + ((int64_t(dslot) * p.H + h) * p.dim + d) * p.dstate;
  constexpr int kMaxPer = 8;  // dstate <= 256
  float s[kMaxPer];
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    const int n = lane + 32 * j;
    s[j] = n < p.dstate ? to_f32(st[n]) : 0.f;
  }
  const TX* x = reinterpret_cast<const TX*>(p.x);
  const TX* dtp = reinterpret_cast<const TX*>(p.dt);
  const TX* z = reinterpret_cast<const TX*>(p.z);
  const TX* Bm = reinterpret_cast<const TX*>(p.Bm);
  const TX* Cm = reinterpret_cast<const TX*>(p.Cm);
  TX* out = reinterpret_cast<TX*>(p.out);
  const float Dv = p.D ? p.D[h * p.D_sh + d * p.D_sd] : 0.f;
  const float dtb = p.dt_bias ? p.dt_bias[h * p.dtb_sh + d * p.dtb_sd] : 0.f;
  for (int t = 0; t < p.T; ++t) {
    const float xv = to_f32(x[b * p.x_sb + t * p.x_st + h * p.x_sh + d]);
    float dtv = to_f32(dtp[b * p.dt_sb + t * p.dt_st + h * p.dt_sh + d * p.dt_sd]) + dtb;
    if (p.softplus) dtv = softplus_f(dtv);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxPer; ++j) {
      const int n = lane + 32 * j;
      if (n < p.dstate) {
        const float a = p.A[h * p.A_sh + d * p.A_sd + n * p.A_sn];
        const float bv = to_f32(Bm[b * p.B_sb + t * p.B_st + g * p.B_sg + n]);
        const float cv = to_f32(Cm[b * p.B_sb + t * p.B_st + g * p.B_sg + n]);
        s[j] = s[j] * __expf(a * dtv) + bv * dtv * xv;
        acc += s[j] * cv;
      }
    }
    acc = warp_reduce_sum(acc);
    if (lane == 0) {
      float y = acc + Dv * xv;
      if (z) {
        const float zv = to_f32(z[b * p.x_sb + t * p.x_st + h * p.x_sh + d]);
        y *= zv * sigmoid_f(zv);
      }
      out[((int64_t(b) * p.T + t) * p.H + h) * p.dim + d] = from_f32<TX>(y);
    }
  }
  if (p.update_state) {
#pragma unroll
    for (int j = 0; j < kMaxPer; ++j) {
      const int n = lane + 32 * j;
      if (n < p.dstate) st_out[n] = from_f32<TS>(s[j]);
    }
  }
  ptx::grid_dep_launch();
}

// ------------------------------------------------------------------ gated delta rule
struct GdrParams {
  float* state;            // [N, HV, K, V] fp32 (k-major)
  const void* q;           // [B, T, H, K]
  const void* k;
  const void* v;           // [B, T, HV, V]
  const void* a;           // [B, T, HV]  (input dependent decay) or null when g is given
  const void* bgate;       // [B, T, HV]  beta logits (or beta itself)
  const float* g;          // [B, T, HV] log-decay given directly (prefill API) or null
  const float* A_log;      // [HV]
  const float* dt_bias;    // [HV]
  void* out;               // [B, T, HV, V]
  const int32_t* state_idx;  // [B] or null
  const int32_t* cu_seqlens; // [B+1] or null (varlen: tokens packed along T of batch 0)
  float scale;
  int B, T, H, HV, K, V, l2norm, beta_is_logit, update_state;
};

// CTA per (sequence, v-head); thread per V column; state column held in registers (K <= 128)
template <typename TX, int KMAX>
__global__ void __launch_bounds__(128) gdr_kernel(const GdrParams p) {
  __shared__ float sq[KMAX], sk[KMAX];
  __shared__ float red[8];
  const int hv = blockIdx.x % p.HV, b = blockIdx.x / p.HV;
  const int h = hv / (p.HV / p.H);
  const int vcol = threadIdx.x;
  ptx::grid_dep_wait();
  int t0 = 0, t1 = p.T, bq = b;
  if (p.cu_seqlens) {
    t0 = p.cu_seqlens[b];
    t1 = p.cu_seqlens[b + 1];
    bq = 0;
  }
  const int slot = p.state_idx ? p.state_idx[b] : b;
  float* S = p.state + (int64_t(slot) * p.HV + hv) * p.K * p.V;
  float s[KMAX];
  const bool active = vcol < p.V;
#pragma unroll
  for (int kk = 0; kk < KMAX; ++kk) s[kk] = (active && kk < p.K && slot >= 0) ? S[int64_t(kk) * p.V + vcol] : 0.f;
  const TX* q = reinterpret_cast<const TX*>(p.q);
  const TX* k = reinterpret_cast<const TX*>(p.k);
  const TX* v = reinterpret_cast<const TX*>(p.v);
  const TX* a = reinterpret_cast<const TX*>(p.a);
  const TX* bg = reinterpret_cast<const TX*>(p.bgate);
  TX* out = reinterpret_cast<TX*>(p.out);
  const int Ttot = p.cu_seqlens ? p.T : p.T;
  for (int t = t0; t < t1; ++t) {
    const int64_t tok = int64_t(bq) * Ttot + t;
    __syncthreads();
    // stage q, k (optionally l2-normalised) in shared memory
    float qv = 0.f, kv = 0.f;
    if (threadIdx.x < p.K) {
      qv = to_f32(q[(tok * p.H + h) * p.K + threadIdx.x]);
      kv = to_f32(k[(tok * p.H + h) * p.K + threadIdx.x]);
    }
    if (p.l2norm) {
      float q2 = warp_reduce_sum(qv * qv), k2 = warp_reduce_sum(kv * kv);
      if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5] = q2;
        red[4 + (threadIdx.x >> 5)] = k2;
      }
      __syncthreads();
      q2 = red[0] + red[1] + red[2] + red[3];
      k2 = red[4] + red[5] + red[6] + red[7];
      qv *= rsqrtf(q2 + 1e-6f);
      kv *= rsqrtf(k2 + 1e-6f);
    }
    if (threadIdx.x < KMAX) {
      sq[threadIdx.x] = qv * p.scale;
      sk[threadIdx.x] = kv;
    }
    __syncthreads();
    float gdec;
    if (p.g) {
      gdec = p.g[tok * p.HV + hv];
    } else {
      const float av = to_f32(a[tok * p.HV + hv]);
      gdec = -__expf(p.A_log[hv]) * softplus_f(av + p.dt_bias[hv]);
    }
    const float decay = __expf(gdec);
    float beta = to_f32(bg[tok * p.HV + hv]);
    if (p.beta_is_logit) beta = sigmoid_f(beta);
    float ks = 0.f;
#pragma unroll
    for (int kk = 0; kk < KMAX; ++kk) {
      s[kk] *= decay;
      ks += sk[kk] * s[kk];
    }
    const float vv = active ? to_f32(v[(tok * p.HV + hv) * p.V + vcol]) : 0.f;
    const float vnew = (vv - ks) * beta;
    float o = 0.f;
#pragma unroll
    for (int kk = 0; kk < KMAX; ++kk) {
      s[kk] += sk[kk] * vnew;
      o += sq[kk] * s[kk];
    }
    if (active) out[(tok * p.HV + hv) * p.V + vcol] = from_f32<TX>(o);
  }
  if (p.update_state && active && slot >= 0) {
#pragma unroll
    for (int kk = 0; kk < KMAX; ++kk)
      if (kk < p.K) S[int64_t(kk) * p.V + vcol] = s[kk];
  }
  ptx::grid_dep_launch();
}

// ------------------------------------------------------------------ concat_mla_k
template <typename T>
__global__ void __launch_bounds__(256)
concat_mla_k_kernel(T* __restrict__ k, const T* __restrict__ k_nope, const T* __restrict__ k_rope, int64_t tokens, int heads,
                    int nope, int rope, int64_t k_st, int64_t k_sh, int64_t n_st, int64_t n_sh, int64_t r_st) {
  constexpr int VN = 16 / sizeof(T);
  const int nv = nope / VN, rv = rope / VN;
  const int64_t per_tok = int64_t(heads) * (nv + rv);
  const int64_t total = tokens * per_tok;
  ptx::grid_dep_wait();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t t = i / per_tok;
    const int r = int(i % per_tok);
    const int h = r / (nv + rv), c = r % (nv + rv);
    Vec16<T> val;
    if (c < nv) val = ld16(k_nope + t * n_st + h * n_sh + c * VN);
    else val = ld16(k_rope + t * r_st + (c - nv) * VN);
    st16(k + t * k_st + h * k_sh + c * VN, val);
  }
  ptx::grid_dep_launch();
}

}  // namespace

// strides: int64[19] = {x_sb,x_st,x_sh, dt_sb,dt_st,dt_sh,dt_sd, A_sh,A_sd,A_sn, B_sb,B_st,B_sg, D_sh,D_sd, dtb_sh,dtb_sd, 0,0}
extern "C" int selective_state_update(void* state, void* x, void* dt, void* A, void* Bm, void* Cm, void* D, void* z,
                                      void* dt_bias, void* out, void* state_idx, void* dst_idx, void* strides_host,
                                      int64_t B, int64_t T, int64_t H, int64_t dim, int64_t dstate, int64_t G,
                                      int64_t pad_slot, int64_t softplus, int64_t update_state, int64_t x_dtype,
                                      int64_t state_dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(dstate <= 256, "selective_state_update: dstate must be <= 256");
  FIB_CHECK(H % G == 0, "selective_state_update: nheads must be a multiple of ngroups");
  if (B == 0) return 0;
  const int64_t* s = (const int64_t*)strides_host;
  SsuParams p;
  p.state = state; p.x = x; p.dt = dt; p.A = (const float*)A; p.Bm = Bm; p.Cm = Cm; p.D = (const float*)D; p.z = z;
  p.dt_bias = (const float*)dt_bias; p.out = out; p.state_idx = (const int32_t*)state_idx; p.dst_idx = (const int32_t*)dst_idx;
  p.x_sb = s[0]; p.x_st = s[1]; p.x_sh = s[2];
  p.dt_sb = s[3]; p.dt_st = s[4]; p.dt_sh = s[5]; p.dt_sd = s[6];
  p.A_sh = s[7]; p.A_sd = s[8]; p.A_sn = s[9];
  p.B_sb = s[10]; p.B_st = s[11]; p.B_sg = s[12];
  p.D_sh = s[13]; p.D_sd = s[14]; p.dtb_sh = s[15]; p.dtb_sd = s[16];
  p.B = (int)B; p.T = (int)T; p.H = (int)H; p.dim = (int)dim; p.dstate = (int)dstate; p.G = (int)G;
  p.pad_slot = (int)pad_slot; p.softplus = (int)softplus; p.update_state = (int)update_state;
  const int64_t rows = B * H * dim;
  LaunchCfg lc(dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), pdl != 0);
  return FIB_DISPATCH_HALF(x_dtype, TX, [&]() -> int {
    if (state_dtype == kF32) {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, ssu_kernel<TX, float>, p));
    } else if (state_dtype == kBF16) {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, ssu_kernel<TX, __nv_bfloat16>, p));
    } else if (state_dtype == kF16) {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, ssu_kernel<TX, __half>, p));
    } else {
      FIB_CHECK(false, "selective_state_update: state must be f32/bf16/f16");
    }
    return 0;
  });
}

extern "C" int gated_delta_rule(void* state, void* q, void* k, void* v, void* a, void* bgate, void* g, void* A_log, void* dt_bias,
                                void* out, void* state_idx, void* cu_seqlens, double scale, int64_t B, int64_t T, int64_t H,
                                int64_t HV, int64_t K, int64_t V, int64_t l2norm, int64_t beta_is_logit, int64_t update_state,
                                int64_t dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(K <= 128 && V <= 128, "gated_delta_rule: head dims must be <= 128");
  FIB_CHECK(HV % H == 0, "gated_delta_rule: HV must be a multiple of H");
  if (B == 0) return 0;
  GdrParams p;
  p.state = (float*)state; p.q = q; p.k = k; p.v = v; p.a = a; p.bgate = bgate; p.g = (const float*)g;
  p.A_log = (const float*)A_log; p.dt_bias = (const float*)dt_bias; p.out = out; p.state_idx = (const int32_t*)state_idx;
  p.cu_seqlens = (const int32_t*)cu_seqlens; p.scale = (float)scale;
  p.B = (int)B; p.T = (int)T; p.H = (int)H; p.HV = (int)HV; p.K = (int)K; p.V = (int)V; p.l2norm = (int)l2norm;
  p.beta_is_logit = (int)beta_is_logit; p.update_state = (int)update_state;
  LaunchCfg lc(dim3((unsigned)(B * HV)), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream_), pdl != 0);
  return FIB_DISPATCH_HALF(dtype, TX, [&]() -> int {
    if (K <= 64) {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, gdr_kernel<TX, 64>, p));
    } else {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, gdr_kernel<TX, 128>, p));
    }
    return 0;
  });
}

extern "C" int concat_mla_k(void* k, void* k_nope, void* k_rope, int64_t tokens, int64_t heads, int64_t nope, int64_t rope,
                            int64_t k_st, int64_t k_sh, int64_t n_st, int64_t n_sh, int64_t r_st, int64_t elt_bytes, int64_t pdl,
                            int64_t stream_) {
  if (tokens == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int64_t vn = 16 / elt_bytes;
  FIB_CHECK(nope % vn == 0 && rope % vn == 0 && k_st % vn == 0 && k_sh % vn == 0 && n_st % vn == 0 && n_sh % vn == 0 && r_st % vn == 0,
            "concat_mla_k: dims and strides must be multiples of 16 bytes");
  const int64_t total = tokens * heads * ((nope + rope) / vn);
  int64_t grid = (total + 255) / 256;
  if (grid > 8 * num_sms()) grid = 8 * num_sms();
  LaunchCfg lc(dim3((unsigned)grid), dim3(256), 0, stream, pdl != 0);
  if (elt_bytes == 2) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, concat_mla_k_kernel<__half>, (__half*)k, (const __half*)k_nope, (const __half*)k_rope,
                                      tokens, (int)heads, (int)nope, (int)rope, k_st, k_sh, n_st, n_sh, r_st));
  } else if (elt_bytes == 1) {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, concat_mla_k_kernel<uint8_t>, (uint8_t*)k, (const uint8_t*)k_nope,
                                      (const uint8_t*)k_rope, tokens, (int)heads, (int)nope, (int)rope, k_st, k_sh, n_st, n_sh, r_st));
  } else {
    FIB_CHECK(false, "concat_mla_k: element size must be 1 or 2 bytes");
  }
  return 0;
}
