// Radix top-k selection (one CTA of 1024 threads per row) with fused page-table / ragged index transforms.
//
// Parity: reference flashinfer/topk.py:508-911 and include/flashinfer/topk.cuh (RadixTopKKernel_Unified,
// page-table / ragged transforms used by DSA-style sparse attention).  Algorithm: 4 x 8-bit MSB-first radix
// passes over order-preserving uint32 keys locate the exact k-th largest key; one ordered compaction pass
// (ballot + prefix sums, so the output is deterministic and index-ordered) emits all elements above the
// threshold plus the required number of ties (tie-break: smaller or larger indices first).
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

constexpr int kThreads = 1024;

__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // larger float -> larger key; NaN sorts high
}

// mode 0: values + indices; 1: page-table transform (out_idx = table[batch][idx]); 2: ragged (idx + offset)
template <typename T>
__global__ void __launch_bounds__(kThreads)
topk_kernel(const T* __restrict__ input, int64_t row_stride, T* __restrict__ out_vals, int32_t* __restrict__ out_idx,
            const int32_t* __restrict__ lengths, const int32_t* __restrict__ row_starts,
            const int32_t* __restrict__ row_to_batch, const int32_t* __restrict__ page_table, int64_t table_stride,
            const int32_t* __restrict__ ragged_offsets, int max_len, int k, int mode, int tie_break) {
  __shared__ int hist[256];
  __shared__ int s_warp[2][32];
  __shared__ uint32_t s_prefix;
  __shared__ int s_krem, s_base_gt, s_base_eq;
  const int row = blockIdx.x;
  const int start = row_starts ? row_starts[row] : 0;
  const int len = lengths ? min(lengths[row], max_len - start) : max_len;
  const T* x = input + int64_t(row) * row_stride + start;
  const int kk = min(k, len);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- radix select of the kk-th largest key ----
  uint32_t prefix = 0, mask = 0;
  int krem = kk;
  if (kk > 0) {
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < len; i += blockDim.x) {
        const uint32_t key = float_key(to_f32(x[i]));
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xff], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int cum = 0, b = 255;
        for (; b > 0; --b) {
          if (cum + hist[b] >= krem) break;
          cum += hist[b];
        }
        s_prefix = prefix | (uint32_t(b) << shift);
        s_krem = krem - cum;
      }
      __syncthreads();
      prefix = s_prefix;
      krem = s_krem;
      mask |= 0xffu << shift;
      __syncthreads();
    }
  }
  // threshold key = prefix; krem = number of ties (== threshold) still to take
  const uint32_t thr = prefix;
  if (threadIdx.x == 0) {
    s_base_gt = 0;
    s_base_eq = 0;
  }
  __syncthreads();
  // output slots: [0, n_gt) for keys > thr in index order, then ties.  n_gt = kk - krem.
  const int n_gt = kk - krem;
  const int batch = row_to_batch ? row_to_batch[row] : row;
  auto emit = [&](int slot, int idx) {
    if (mode == 0) {
      out_idx[int64_t(row) * k + slot] = idx;
      if (out_vals) out_vals[int64_t(row) * k + slot] = x[idx];
    } else if (mode == 1) {
      out_idx[int64_t(row) * k + slot] = page_table[int64_t(batch) * table_stride + start + idx];
    } else {
      out_idx[int64_t(row) * k + slot] = idx + start + (ragged_offsets ? ragged_offsets[row] : 0);
    }
  };
  const int nchunks = (len + blockDim.x - 1) / blockDim.x;
  for (int c = 0; c < nchunks; ++c) {
    // tie_break LARGE walks the row backwards so that larger indices claim the tie quota first
    const int i = (tie_break == 2) ? len - 1 - (c * blockDim.x + threadIdx.x) : c * blockDim.x + threadIdx.x;
    const bool valid = i >= 0 && i < len;
    const uint32_t key = valid ? float_key(to_f32(x[i])) : 0u;
    const bool gt = valid && key > thr;
    const bool eq = valid && key == thr;
    const uint32_t bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    const int pg = __popc(bg & ((1u << lane) - 1)), pe = __popc(be & ((1u << lane) - 1));
    if (lane == 0) {
      s_warp[0][warp] = __popc(bg);
      s_warp[1][warp] = __popc(be);
    }
    __syncthreads();
    int wg = 0, we = 0, tg = 0, te = 0;
    {
      const int nw = blockDim.x >> 5;
      for (int w = 0; w < nw; ++w) {
        const int a = s_warp[0][w], b = s_warp[1][w];
        if (w < warp) {
          wg += a;
          we += b;
        }
        tg += a;
        te += b;
      }
    }
    const int base_gt = s_base_gt, base_eq = s_base_eq;
    if (gt) emit(base_gt + wg + pg, i);
    if (eq) {
      const int e = base_eq + we + pe;
      if (e < krem) emit(n_gt + e, i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      s_base_gt = base_gt + tg;
      s_base_eq = base_eq + te;
    }
    __syncthreads();
  }
  // pad rows shorter than k
  for (int j = kk + threadIdx.x; j < k; j += blockDim.x) {
    out_idx[int64_t(row) * k + j] = -1;
    if (mode == 0 && out_vals) out_vals[int64_t(row) * k + j] = from_f32<T>(-INFINITY);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Cluster flavour: ONE ROW PER THREAD-BLOCK CLUSTER (2 / 4 / 8 CTAs).  Few long rows (the DSA indexer at decode: a handful of
// rows of up to 128 K scores, k = 2048) leave a one-CTA-per-row kernel on a few SMs; here every CTA of the cluster owns a
// contiguous slice of the row, keeps its order-preserving keys in shared memory (one global read of the row), and the
// per-pass digit histograms / the compaction bases are combined through distributed shared memory (ld.shared::cluster).
// The output is identical to topk_kernel (deterministic, index ordered).  Parity: reference
// include/flashinfer/fast_topk_clusters_exact.cuh:408-497 and topk.cuh:1094-1550 (multi-CTA radix top-k).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_cluster_s32(uint32_t addr) {
  int v;
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
topk_cluster_kernel(const T* __restrict__ input, int64_t row_stride, T* __restrict__ out_vals, int32_t* __restrict__ out_idx,
                    const int32_t* __restrict__ lengths, const int32_t* __restrict__ row_starts,
                    const int32_t* __restrict__ row_to_batch, const int32_t* __restrict__ page_table, int64_t table_stride,
                    const int32_t* __restrict__ ragged_offsets, int max_len, int k, int mode, int tie_break, int cs, int key_cap) {
  extern __shared__ uint32_t s_keys[];  // [key_cap]: the slice's keys when it fits
  __shared__ int hist[256];
  __shared__ int tot[256];
  __shared__ int s_warp[2][32];
  __shared__ uint32_t s_prefix;
  __shared__ int s_krem, s_base_gt, s_base_eq;
  __shared__ int s_cnt[2];
  const int row = blockIdx.x / cs;
  const int crank = int(ptx::cluster_ctarank());
  const int start = row_starts ? row_starts[row] : 0;
  const int len = lengths ? min(lengths[row], max_len - start) : max_len;
  const T* x = input + int64_t(row) * row_stride + start;
  const int kk = min(k, len);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (len + cs - 1) / cs;
  const int lo = min(len, crank * per), hi = min(len, lo + per);
  const int n = hi - lo;
  const bool cached = n <= key_cap;
  if (cached)
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_keys[i] = float_key(to_f32(x[lo + i]));
  auto key_at = [&](int i) { return cached ? s_keys[i] : float_key(to_f32(x[lo + i])); };
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  __syncthreads();

  uint32_t prefix = 0, mask = 0;
  int krem = kk;
  if (kk > 0) {
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t key = key_at(i);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xff], 1);
      }
      __syncthreads();
      ptx::cluster_sync();  // every CTA's histogram of this pass is complete
      if (threadIdx.x < 256) {
        int t = 0;
        const uint32_t a = ptx::smem_u32(&hist[threadIdx.x]);
        for (int r = 0; r < cs; ++r) t += ld_cluster_s32(ptx::mapa(a, uint32_t(r)));
        tot[threadIdx.x] = t;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int cum = 0, b = 255;
        for (; b > 0; --b) {
          if (cum + tot[b] >= krem) break;
          cum += tot[b];
        }
        s_prefix = prefix | (uint32_t(b) << shift);
        s_krem = krem - cum;
      }
      __syncthreads();
      prefix = s_prefix;
      krem = s_krem;
      mask |= 0xffu << shift;
      ptx::cluster_sync();  // peers have read my histogram: it may be cleared for the next pass
    }
  }
  const uint32_t thr = prefix;
  const int n_gt = kk - krem;
  // ---- local counts of keys above / equal to the threshold, exchanged over the cluster -> this CTA's output bases
  {
    int cg = 0, ce = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = key_at(i);
      cg += key > thr;
      ce += key == thr;
    }
    cg = warp_reduce_sum(cg);
    ce = warp_reduce_sum(ce);
    if (lane == 0 && kk > 0) {
      atomicAdd(&s_cnt[0], cg);
      atomicAdd(&s_cnt[1], ce);
    }
  }
  __syncthreads();
  ptx::cluster_sync();
  if (threadIdx.x == 0) {
    int bg = 0, be = 0;
    for (int r = 0; r < cs; ++r) {
      const bool before = (tie_break == 2) ? (r > crank) : (r < crank);  // LARGE: the row is walked backwards, last slice first
      if (before) {
        bg += ld_cluster_s32(ptx::mapa(ptx::smem_u32(&s_cnt[0]), uint32_t(r)));
        be += ld_cluster_s32(ptx::mapa(ptx::smem_u32(&s_cnt[1]), uint32_t(r)));
      }
    }
    s_base_gt = bg;
    s_base_eq = be;
  }
  __syncthreads();
  const int batch = row_to_batch ? row_to_batch[row] : row;
  auto emit = [&](int slot, int idx) {
    if (mode == 0) {
      out_idx[int64_t(row) * k + slot] = idx;
      if (out_vals) out_vals[int64_t(row) * k + slot] = x[idx];
    } else if (mode == 1) {
      out_idx[int64_t(row) * k + slot] = page_table[int64_t(batch) * table_stride + start + idx];
    } else {
      out_idx[int64_t(row) * k + slot] = idx + start + (ragged_offsets ? ragged_offsets[row] : 0);
    }
  };
  const int nchunks = kk > 0 ? (n + blockDim.x - 1) / blockDim.x : 0;
  for (int c = 0; c < nchunks; ++c) {
    const int li = (tie_break == 2) ? n - 1 - (c * blockDim.x + threadIdx.x) : c * blockDim.x + threadIdx.x;
    const bool valid = li >= 0 && li < n;
    const uint32_t key = valid ? key_at(li) : 0u;
    const bool gt = valid && key > thr;
    const bool eq = valid && key == thr;
    const uint32_t bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    const int pg = __popc(bg & ((1u << lane) - 1)), pe = __popc(be & ((1u << lane) - 1));
    if (lane == 0) {
      s_warp[0][warp] = __popc(bg);
      s_warp[1][warp] = __popc(be);
    }
    __syncthreads();
    int wg = 0, we = 0, tg = 0, te = 0;
    {
      const int nw = blockDim.x >> 5;
      for (int w = 0; w < nw; ++w) {
        const int a = s_warp[0][w], b = s_warp[1][w];
        if (w < warp) {
          wg += a;
          we += b;
        }
        tg += a;
        te += b;
      }
    }
    const int base_gt = s_base_gt, base_eq = s_base_eq;
    if (gt) emit(base_gt + wg + pg, lo + li);
    if (eq) {
      const int e = base_eq + we + pe;
      if (e < krem) emit(n_gt + e, lo + li);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      s_base_gt = base_gt + tg;
      s_base_eq = base_eq + te;
    }
    __syncthreads();
  }
  if (crank == 0)
    for (int j = kk + threadIdx.x; j < k; j += blockDim.x) {
      out_idx[int64_t(row) * k + j] = -1;
      if (mode == 0 && out_vals) out_vals[int64_t(row) * k + j] = from_f32<T>(-INFINITY);
    }
  ptx::cluster_sync();  // nobody exits while a peer may still read its counters
}

}  // namespace

// `clusters`: 0 = auto (cluster kernel when the rows alone cannot fill the machine), 1 = one CTA per row, 2 / 4 / 8 = cluster size
extern "C" int topk_run_ex(void* input, int64_t row_stride, void* out_vals, void* out_idx, void* lengths, void* row_starts,
                           void* row_to_batch, void* page_table, int64_t table_stride, void* ragged_offsets, int64_t num_rows,
                           int64_t max_len, int64_t k, int64_t mode, int64_t tie_break, int64_t dtype, int64_t clusters, int64_t pdl,
                           int64_t stream_) {
  if (num_rows == 0 || k == 0) return 0;
  FIB_CHECK(clusters == 0 || clusters == 1 || clusters == 2 || clusters == 4 || clusters == 8, "topk: clusters must be 0, 1, 2, 4 or 8");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  int cs = int(clusters);
  if (cs == 0) {
    cs = 1;
    const int sms = num_sms();
    if (max_len >= 8192)
      while (cs < 8 && num_rows * cs * 2 <= sms && max_len / (cs * 2) >= 2048) cs *= 2;
  }
  if (cs == 1) {
    ++launch_counter();
    return FIB_DISPATCH_FLOAT(dtype, T, [&]() -> int {
      topk_kernel<T><<<(unsigned)num_rows, kThreads, 0, s>>>(
          (const T*)input, row_stride, (T*)out_vals, (int32_t*)out_idx, (const int32_t*)lengths, (const int32_t*)row_starts,
          (const int32_t*)row_to_batch, (const int32_t*)page_table, table_stride, (const int32_t*)ragged_offsets, (int)max_len, (int)k,
          (int)mode, (int)tie_break);
      FIB_CUDA_CHECK(cudaGetLastError());
      return 0;
    });
  }
  const int per = int((max_len + cs - 1) / cs);
  int key_cap = per <= 48 * 1024 ? per : 0;  // keys of the slice in shared memory (<= 192 KB), else re-read from global / L2
  const size_t smem = size_t(key_cap) * 4;
  return FIB_DISPATCH_FLOAT(dtype, T, [&]() -> int {
    auto kern = topk_cluster_kernel<T>;
    FIB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LaunchCfg lc(dim3((unsigned)(num_rows * cs)), dim3(kThreads), smem, s, pdl != 0, cs);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, (const T*)input, row_stride, (T*)out_vals, (int32_t*)out_idx,
                                      (const int32_t*)lengths, (const int32_t*)row_starts, (const int32_t*)row_to_batch,
                                      (const int32_t*)page_table, table_stride, (const int32_t*)ragged_offsets, (int)max_len, (int)k,
                                      (int)mode, (int)tie_break, cs, key_cap));
    return 0;
  });
}

extern "C" int topk_run(void* input, int64_t row_stride, void* out_vals, void* out_idx, void* lengths, void* row_starts,
                        void* row_to_batch, void* page_table, int64_t table_stride, void* ragged_offsets,
                        int64_t num_rows, int64_t max_len, int64_t k, int64_t mode, int64_t tie_break, int64_t dtype,
                        int64_t stream_) {
  if (num_rows == 0 || k == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  ++launch_counter();
  return FIB_DISPATCH_FLOAT(dtype, T, [&]() -> int {
    topk_kernel<T><<<(unsigned)num_rows, kThreads, 0, s>>>(
        (const T*)input, row_stride, (T*)out_vals, (int32_t*)out_idx, (const int32_t*)lengths,
        (const int32_t*)row_starts, (const int32_t*)row_to_batch, (const int32_t*)page_table, table_stride,
        (const int32_t*)ragged_offsets, (int)max_len, (int)k, (int)mode, (int)tie_break);
    FIB_CUDA_CHECK(cudaGetLastError());
    return 0;
  });
}
