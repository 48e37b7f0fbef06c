// Sorting-free sampling kernels (one CTA of 1024 threads per row, vocab streamed from L2/HBM with 16 B
// vectors).  Parity: reference flashinfer/sampling.py:737-1957 and include/flashinfer/sampling.cuh
// (OnlineSoftmax :293-560, SamplingFromProb, TopK/TopP/MinP/TopKTopP rejection sampling :835-960,
// renorm / mask kernels, ChainSpeculativeSampling :1858).
//
// Algorithms:
//  * softmax: two-pass online softmax with temperature (per-row tensor or scalar).
//  * inverse-CDF sampling: block-wide prefix scan over the row, Philox(seed, offset) uniform per row.
//  * top-k / top-p / min-p / joint top-k+top-p sampling: dual-pivot rejection sampling -- draw a token
//    from the probability mass above `low`, then count/sum the entries above pivot0 = p[token] and
//    pivot1 = (pivot0 + high) / 2 to either accept or shrink (low, high).  No sort, O(rounds * V).
//  * renorm / mask: bisection on the probability (logit) threshold with fused count + sum reductions.
#include <float.h>
#include <curand_kernel.h>
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

constexpr int kThreads = 1024;

struct Pair {
  float sum;
  int cnt;
};

__device__ __forceinline__ float block_reduce_sum(float v, float* sm) {
  v = warp_reduce_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = (l < (blockDim.x >> 5)) ? sm[l] : 0.f;
  r = warp_reduce_sum(r);
  return __shfl_sync(0xffffffffu, r, 0);
}
__device__ __forceinline__ float block_reduce_max(float v, float* sm) {
  v = warp_reduce_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = (l < (blockDim.x >> 5)) ? sm[l] : -INFINITY;
  r = warp_reduce_max(r);
  return __shfl_sync(0xffffffffu, r, 0);
}
__device__ __forceinline__ int block_reduce_sum_int(int v, int* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  int r = (l < (blockDim.x >> 5)) ? sm[l] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return __shfl_sync(0xffffffffu, r, 0);
}

// Block-wide inverse-CDF draw from the entries of `p` that satisfy `keep(x)`: returns the first index
// whose running sum of kept entries exceeds `target` (or the last kept index for round-off), -1 if none.
// kFraction: `target` is a fraction in [0, 1) of the kept mass (the scan's own total is used, no separate sum pass).
template <bool kFraction = false, typename Keep>
__device__ int block_sample(const float* __restrict__ p, int V, float target, Keep keep, float* smf, int* smi) {
  // One pass + one block scan: thread t owns the contiguous index range [t * per, (t + 1) * per), sums its kept entries,
  // a single block-wide exclusive scan locates the owner of `target`, and only that thread walks its range again.
  __shared__ float s_warp[32];
  __shared__ int s_found;
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    s_found = -1;
    s_last = -1;
  }
  __syncthreads();
  const int per = (((V + int(blockDim.x) - 1) / int(blockDim.x)) + 3) & ~3;
  const int i0 = threadIdx.x * per;
  const int i1 = min(V, i0 + per);
  float local = 0.f;
  int last_here = -1;
  if (i0 < V) {
    if ((reinterpret_cast<uintptr_t>(p + i0) & 15) == 0) {
      int i = i0;
      for (; i + 4 <= i1; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(p + i);
        const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (keep(x[e]) && x[e] > 0.f) {
            local += x[e];
            last_here = i + e;
          }
      }
      for (; i < i1; ++i) {
        const float v = p[i];
        if (keep(v) && v > 0.f) {
          local += v;
          last_here = i;
        }
      }
    } else {
      for (int i = i0; i < i1; ++i) {
        const float v = p[i];
        if (keep(v) && v > 0.f) {
          local += v;
          last_here = i;
        }
      }
    }
  }
  if (last_here >= 0) atomicMax(&s_last, last_here);
  // block exclusive scan of `local`
  float incl = local;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  float warp_off = 0.f;
  {
    const float wv = (lane < int(blockDim.x >> 5)) ? s_warp[lane] : 0.f;
    float wincl = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, wincl, o);
      if (lane >= o) wincl += t;
    }
    warp_off = __shfl_sync(0xffffffffu, wincl - wv, warp);
  }
  const float excl = warp_off + incl - local;
  if constexpr (kFraction) {
    // total = inclusive value of the last thread; every warp's lane 31 wrote its inclusive warp sum to s_warp
    float total = 0.f;
    {
      const float wv = (lane < int(blockDim.x >> 5)) ? s_warp[lane] : 0.f;
      total = warp_reduce_sum(wv);
    }
    target *= total;
  }
  if (local > 0.f && excl <= target && target < excl + local) {
    float acc = excl;
    for (int i = i0; i < i1; ++i) {
      const float v = p[i];
      if (keep(v) && v > 0.f) {
        if (target < acc + v) {
          atomicCAS(&s_found, -1, i);
          break;
        }
        acc += v;
      }
    }
  }
  __syncthreads();
  const int f = s_found >= 0 ? s_found : s_last;
  __syncthreads();
  (void)smf;
  (void)smi;
  return f;
}

__device__ __forceinline__ float row_uniform(uint64_t seed, uint64_t offset, int row) {
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)row, offset, &st);
  return curand_uniform(&st);  // (0, 1]
}

// ------------------------------------------------------------------ softmax
// One thread-block CLUSTER per row: every CTA keeps its slice of the row in shared memory (the row crosses HBM once
// in each direction), the running max and the sum are exchanged through distributed shared memory (two cluster
// barriers), so small batches still use all SMs.  Falls back to plain strided global passes when the slice does not
// fit (huge vocab with cluster size 1) or the row is not 16-byte aligned.
template <bool kSmemSlice>
__global__ void __launch_bounds__(kThreads)
softmax_kernel(const float* __restrict__ logits, float* __restrict__ probs, const float* __restrict__ temp_arr,
               float temp_val, int V, int csize, int slice) {
  __shared__ float smf[32];
  __shared__ float xch[2][8];  // [max | sum][cluster rank]
  extern __shared__ float sl[];
  const int crank = csize > 1 ? int(ptx::cluster_ctarank()) : 0;
  if (csize > 1) ptx::cluster_arrive();  // paired with the wait before the first DSMEM store: peers must be running
  const int row = blockIdx.x / csize;
  const float* x = logits + int64_t(row) * V;
  float* y = probs + int64_t(row) * V;
  const float t = temp_arr ? temp_arr[row] : temp_val;
  const float inv_t = (t > 0.f) ? 1.f / t : 1.f;
  const int i0 = crank * slice;
  const int n = max(0, min(slice, V - i0));
  float m = -INFINITY;
  if constexpr (kSmemSlice) {
    const float4* x4 = reinterpret_cast<const float4*>(x + i0);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
      float4 v = x4[i];
      v.x *= inv_t; v.y *= inv_t; v.z *= inv_t; v.w *= inv_t;
      reinterpret_cast<float4*>(sl)[i] = v;
      m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    for (int i = (n / 4) * 4 + threadIdx.x; i < n; i += blockDim.x) {
      const float v = x[i0 + i] * inv_t;
      sl[i] = v;
      m = fmaxf(m, v);
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, x[i0 + i] * inv_t);
  }
  m = block_reduce_max(m, smf);
  if (csize > 1) {
    ptx::cluster_wait();
    if (threadIdx.x < csize) ptx::st_dsmem_f32(ptx::mapa(ptx::smem_u32(&xch[0][crank]), threadIdx.x), m);
    ptx::cluster_sync();
    for (int r = 0; r < csize; ++r) m = fmaxf(m, xch[0][r]);
  }
  float s = 0.f;
  if constexpr (kSmemSlice) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float e = __expf(sl[i] - m);
      sl[i] = e;
      s += e;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += __expf(x[i0 + i] * inv_t - m);
  }
  s = block_reduce_sum(s, smf);
  if (csize > 1) {
    if (threadIdx.x < csize) ptx::st_dsmem_f32(ptx::mapa(ptx::smem_u32(&xch[1][crank]), threadIdx.x), s);
    ptx::cluster_sync();
    s = 0.f;
    for (int r = 0; r < csize; ++r) s += xch[1][r];
  }
  const float inv = 1.f / s;
  if constexpr (kSmemSlice) {
    float4* y4 = reinterpret_cast<float4*>(y + i0);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) {
      float4 v = reinterpret_cast<const float4*>(sl)[i];
      v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
      y4[i] = v;
    }
    for (int i = (n / 4) * 4 + threadIdx.x; i < n; i += blockDim.x) y[i0 + i] = sl[i] * inv;
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) y[i0 + i] = __expf(x[i0 + i] * inv_t - m) * inv;
  }
}

// ------------------------------------------------------------------ unified sampling kernel
// mode: 0 plain, 1 top-k, 2 top-p, 3 min-p, 4 top-k + top-p (joint)
__global__ void __launch_bounds__(kThreads)
sampling_kernel(const float* __restrict__ probs, int32_t* __restrict__ out, const int32_t* __restrict__ indices,
                const float* __restrict__ top_p_arr, float top_p_val, const int32_t* __restrict__ top_k_arr,
                int top_k_val, const float* __restrict__ min_p_arr, float min_p_val, int V, int mode, uint64_t seed,
                uint64_t offset, int max_rounds, uint8_t* __restrict__ success) {
  __shared__ float smf[32];
  __shared__ int smi[32];
  const int row = blockIdx.x;
  const int prow = indices ? indices[row] : row;
  const float* p = probs + int64_t(prow) * V;
  const float top_p = top_p_arr ? top_p_arr[row] : top_p_val;
  int top_k = top_k_arr ? top_k_arr[row] : top_k_val;
  if (top_k <= 0 || top_k > V) top_k = V;
  const float min_p = min_p_arr ? min_p_arr[row] : min_p_val;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)row, offset, &st);

  if (mode == 0 || mode == 3) {
    float thresh = 0.f;
    float total = 0.f;
    if (mode == 3) {
      float mx = 0.f;
      for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, p[i]);
      mx = block_reduce_max(mx, smf);
      thresh = mx * min_p;
    }
    (void)total;
    const float u = curand_uniform(&st);
    const int tok = block_sample<true>(p, V, 1.f - u, [=](float v) { return v >= thresh; }, smf, smi);
    if (threadIdx.x == 0) {
      out[row] = tok < 0 ? 0 : tok;
      if (success) success[row] = tok >= 0;
    }
    return;
  }

  // ---- dual-pivot rejection sampling ----
  float low = 0.f, high = 1.f;
  float q = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) q += p[i];
  q = block_reduce_sum(q, smf);
  {
    float mx = 0.f;
    for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, p[i]);
    high = block_reduce_max(mx, smf);
  }
  int tok = -1;
  bool accepted = false;
  for (int round = 0; round < max_rounds; ++round) {
    const float u = curand_uniform(&st);
    const float target = (1.f - u) * q;
    const float lo = low;
    tok = block_sample(p, V, target, [=](float v) { return v > lo; }, smf, smi);
    if (tok < 0) break;
    const float pivot0 = p[tok];
    const float pivot1 = 0.5f * (pivot0 + high);
    float s0 = 0.f, s1 = 0.f;
    int c0 = 0, c1 = 0;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
      const float v = p[i];
      if (v > pivot0) {
        s0 += v;
        ++c0;
      }
      if (v > pivot1) {
        s1 += v;
        ++c1;
      }
    }
    s0 = block_reduce_sum(s0, smf);
    s1 = block_reduce_sum(s1, smf);
    c0 = block_reduce_sum_int(c0, smi);
    c1 = block_reduce_sum_int(c1, smi);
    const bool k_ok0 = (mode == 2) || (c0 < top_k);
    const bool p_ok0 = (mode == 1) || (s0 < top_p);
    if (k_ok0 && p_ok0) {
      accepted = true;
      break;
    }
    const bool k_ok1 = (mode == 2) || (c1 < top_k);
    const bool p_ok1 = (mode == 1) || (s1 < top_p);
    if (k_ok1 && p_ok1) {
      low = pivot0;
      high = pivot1;
      q = s0;
    } else {
      low = pivot1;
      q = s1;
    }
  }
  if (threadIdx.x == 0) {
    out[row] = tok < 0 ? 0 : tok;
    if (success) success[row] = accepted;
  }
}

// ------------------------------------------------------------------ renorm / mask (threshold bisection)
// mode 0: top-p renorm (probs), 1: top-k renorm (probs), 2: top-k mask (logits -> -inf)
__global__ void __launch_bounds__(kThreads)
renorm_kernel(const float* __restrict__ in, float* __restrict__ outp, const float* __restrict__ top_p_arr,
              float top_p_val, const int32_t* __restrict__ top_k_arr, int top_k_val, int V, int mode) {
  __shared__ float smf[32];
  __shared__ int smi[32];
  const int row = blockIdx.x;
  const float* x = in + int64_t(row) * V;
  float* y = outp + int64_t(row) * V;
  const float top_p = top_p_arr ? top_p_arr[row] : top_p_val;
  int top_k = top_k_arr ? top_k_arr[row] : top_k_val;
  if (top_k <= 0 || top_k > V) top_k = V;
  float mx = -INFINITY, mn = INFINITY;
  // min over FINITE entries only: rows pre-masked with -inf (banned tokens, grammar masks) are the normal input of
  // top_k_mask_logits, and a -inf lower bound would stop the bisection on its first step (mid = -inf <= lo)
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float v = x[i];
    if (v == v) mx = fmaxf(mx, v);                    // NaN never moves a bound
    if (v > -FLT_MAX && v < FLT_MAX) mn = fminf(mn, v);
  }
  mx = block_reduce_max(mx, smf);
  mn = -block_reduce_max(-mn, smf);
  if (!(mn < FLT_MAX)) mn = -FLT_MAX * 0.5f;          // no finite entry at all
  if (mode == 2 && mn < -FLT_MAX * 0.25f) mn = -FLT_MAX * 0.25f;
  // find the largest threshold `lo` such that the kept set {x >= ... } still satisfies the constraint:
  //  top-p: sum(x > t) >= p  (keep everything above t);  top-k: count(x > t) >= k
  float lo = (mode == 2) ? mn - fmaxf(1.f, fabsf(mn) * 1e-6f) : 0.f, hi = mx;
  // invariant: constraint(lo) holds (kept set large enough), constraint(hi) fails
  for (int it = 0; it < 40; ++it) {
    const float mid = 0.5f * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    float s = 0.f;
    int c = 0;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
      const float v = x[i];
      if (v > mid) {
        s += v;
        ++c;
      }
    }
    bool ok;
    if (mode == 0) {
      s = block_reduce_sum(s, smf);
      ok = s >= top_p;
    } else {
      c = block_reduce_sum_int(c, smi);
      ok = c >= top_k;
    }
    if (ok) lo = mid; else hi = mid;
  }
  // kept set = {x > lo}; it is the smallest superset reachable by a threshold (ties kept together)
  if (mode == 2) {
    for (int i = threadIdx.x; i < V; i += blockDim.x) y[i] = (x[i] > lo) ? x[i] : -INFINITY;
    return;
  }
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x)
    if (x[i] > lo) s += x[i];
  s = block_reduce_sum(s, smf);
  const float inv = s > 0.f ? 1.f / s : 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) y[i] = (x[i] > lo) ? x[i] * inv : 0.f;
}

// ------------------------------------------------------------------ chain speculative sampling
// draft_probs [B, n, V], draft_ids [B, n], target_probs [B, n+1, V] -> out_ids [B, n+1] (-1 padded),
// accepted/emitted counters.  One CTA per request, sequential over draft positions.
__global__ void __launch_bounds__(kThreads)
chain_spec_kernel(const float* __restrict__ draft_probs, const int32_t* __restrict__ draft_ids,
                  const float* __restrict__ target_probs, int32_t* __restrict__ out_ids,
                  int32_t* __restrict__ accepted_num, int32_t* __restrict__ emitted_num, int n, int V, int deterministic,
                  uint64_t seed, uint64_t offset) {
  __shared__ float smf[32];
  __shared__ int smi[32];
  extern __shared__ float diff[];  // not used (streamed), kept for future caching
  (void)diff;
  const int b = blockIdx.x;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)b, offset, &st);
  int pos = 0;
  int emitted = 0;
  bool rejected = false;
  for (; pos < n; ++pos) {
    const int tok = draft_ids[b * n + pos];
    const float q = target_probs[(int64_t(b) * (n + 1) + pos) * V + tok];
    const float pd = draft_probs[(int64_t(b) * n + pos) * V + tok];
    const float u = curand_uniform(&st);
    if (u * pd < q) {  // accept with prob min(1, q/p)
      if (threadIdx.x == 0) out_ids[b * (n + 1) + pos] = tok;
      ++emitted;
    } else {
      rejected = true;
      break;
    }
  }
  // accepted_num counts how many draft tokens WOULD be accepted independently (reference semantics)
  if (threadIdx.x == 0 && accepted_num) {
    int acc = 0;
    curandStatePhilox4_32_10_t st2;
    curand_init(seed, (unsigned long long)b, offset, &st2);
    for (int i = 0; i < n; ++i) {
      const int tok = draft_ids[b * n + i];
      const float q = target_probs[(int64_t(b) * (n + 1) + i) * V + tok];
      const float pd = draft_probs[(int64_t(b) * n + i) * V + tok];
      const float u = curand_uniform(&st2);
      if (u * pd < q) ++acc;
    }
    accepted_num[b] += acc;
  }
  // sample the correction / bonus token from relu(target - draft) (or target at the bonus position)
  const float* tp = target_probs + (int64_t(b) * (n + 1) + pos) * V;
  const float* dp = rejected ? draft_probs + (int64_t(b) * n + pos) * V : nullptr;
  float total = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float d = dp ? fmaxf(tp[i] - dp[i], 0.f) : tp[i];
    total += d;
  }
  total = block_reduce_sum(total, smf);
  const float u = curand_uniform(&st);
  const float target = (1.f - u) * total;
  // inverse CDF over the (implicit) difference distribution
  __shared__ int s_tok;
  __shared__ int s_lastpos;
  if (threadIdx.x == 0) {
    s_tok = -1;
    s_lastpos = -1;
  }
  __syncthreads();
  float running = 0.f;
  for (int base = 0; base < V && s_tok < 0; base += blockDim.x) {
    const int i = base + threadIdx.x;
    float v = 0.f;
    if (i < V) v = dp ? fmaxf(tp[i] - dp[i], 0.f) : tp[i];
    float incl = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) smf[warp] = incl;
    __syncthreads();
    const float wv = (lane < (blockDim.x >> 5)) ? smf[lane] : 0.f;
    float wincl = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, wincl, o);
      if (lane >= o) wincl += t;
    }
    const float warp_off = __shfl_sync(0xffffffffu, wincl - wv, warp);
    const float tot = __shfl_sync(0xffffffffu, wincl, 31);
    const float excl = running + warp_off + incl - v;
    if (v > 0.f) {
      atomicMax(&s_lastpos, i);
      if (excl <= target && target < excl + v) atomicCAS(&s_tok, -1, i);
    }
    running += tot;
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = s_tok >= 0 ? s_tok : (s_lastpos >= 0 ? s_lastpos : 0);
    out_ids[b * (n + 1) + pos] = t;
    for (int i = pos + 1; i <= n; ++i) out_ids[b * (n + 1) + i] = -1;
    if (emitted_num) emitted_num[b] += emitted;
  }
  (void)smi;
  (void)deterministic;
}

}  // namespace

extern "C" int softmax_run(void* logits, void* probs, void* temp_arr, double temp_val, int64_t rows, int64_t V,
                           int64_t stream_) {
  if (rows == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  // cluster size: enough CTAs to cover the SMs and a slice that fits in shared memory
  int csize = 1;
  while (csize < 8 && (rows * csize < 2 * num_sms() || (V + csize - 1) / csize * 4 > 200 * 1024)) csize <<= 1;
  int slice = int((V + csize - 1) / csize);
  slice = (slice + 3) / 4 * 4;
  const bool smem_ok = (V % 4 == 0) && int64_t(slice) * 4 <= 200 * 1024;
  const size_t smem = smem_ok ? size_t(slice) * 4 : 0;
  LaunchCfg lc(dim3((unsigned)(rows * csize)), dim3(kThreads), smem, s, false, csize);
  if (smem_ok) {
    static bool set = false;
    if (!set) {
      FIB_CUDA_CHECK(cudaFuncSetAttribute(softmax_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      set = true;
    }
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, softmax_kernel<true>, (const float*)logits, (float*)probs, (const float*)temp_arr,
                                      (float)temp_val, (int)V, csize, slice));
  } else {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, softmax_kernel<false>, (const float*)logits, (float*)probs, (const float*)temp_arr,
                                      (float)temp_val, (int)V, csize, slice));
  }
  return 0;
}

extern "C" int sampling_run(void* probs, void* out, void* indices, void* top_p_arr, double top_p_val, void* top_k_arr,
                            int64_t top_k_val, void* min_p_arr, double min_p_val, int64_t rows, int64_t V, int64_t mode,
                            int64_t seed, int64_t offset, int64_t max_rounds, void* success, int64_t stream_) {
  if (rows == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  ++launch_counter();
  sampling_kernel<<<(unsigned)rows, kThreads, 0, s>>>((const float*)probs, (int32_t*)out, (const int32_t*)indices,
                                                      (const float*)top_p_arr, (float)top_p_val,
                                                      (const int32_t*)top_k_arr, (int)top_k_val, (const float*)min_p_arr,
                                                      (float)min_p_val, (int)V, (int)mode, (uint64_t)seed,
                                                      (uint64_t)offset, (int)max_rounds, (uint8_t*)success);
  FIB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int renorm_run(void* in, void* out, void* top_p_arr, double top_p_val, void* top_k_arr, int64_t top_k_val,
                          int64_t rows, int64_t V, int64_t mode, int64_t stream_) {
  if (rows == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  ++launch_counter();
  renorm_kernel<<<(unsigned)rows, kThreads, 0, s>>>((const float*)in, (float*)out, (const float*)top_p_arr,
                                                    (float)top_p_val, (const int32_t*)top_k_arr, (int)top_k_val, (int)V,
                                                    (int)mode);
  FIB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int chain_speculative_sampling_run(void* draft_probs, void* draft_ids, void* target_probs, void* out_ids,
                                              void* accepted_num, void* emitted_num, int64_t batch, int64_t n, int64_t V,
                                              int64_t deterministic, int64_t seed, int64_t offset, int64_t stream_) {
  if (batch == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  ++launch_counter();
  chain_spec_kernel<<<(unsigned)batch, kThreads, 0, s>>>((const float*)draft_probs, (const int32_t*)draft_ids,
                                                         (const float*)target_probs, (int32_t*)out_ids,
                                                         (int32_t*)accepted_num, (int32_t*)emitted_num, (int)n, (int)V,
                                                         (int)deterministic, (uint64_t)seed, (uint64_t)offset);
  FIB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
