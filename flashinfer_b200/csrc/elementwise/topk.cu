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

}  // namespace

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
