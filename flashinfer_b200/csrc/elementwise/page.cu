// Paged KV-cache maintenance ops (SIMT, 16 B vectors, PDL).
// Parity: reference flashinfer/page.py:128-406 and include/flashinfer/page.cuh:223-560
// (append_paged_kv_cache, append_paged_mla_kv_cache, get_batch_indices_positions).
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

// one warp-slice per (token, head): copies head_dim elements of K and V into the page slot.
template <typename T>
__global__ void __launch_bounds__(256)
append_paged_kv_kernel(const T* __restrict__ key, const T* __restrict__ value, const int32_t* __restrict__ batch_indices,
                       const int32_t* __restrict__ positions, T* __restrict__ k_cache, T* __restrict__ v_cache,
                       const int32_t* __restrict__ kv_indices, const int32_t* __restrict__ kv_indptr, int64_t nnz,
                       int num_heads, int head_dim, int page_size, int64_t key_sn, int64_t key_sh, int64_t val_sn,
                       int64_t val_sh, int64_t c_sp, int64_t c_sn, int64_t c_sh) {
  constexpr int VN = 16 / sizeof(T);
  const int vec_per_head = head_dim / VN;
  const int64_t total = nnz * num_heads * vec_per_head;
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();  // early trigger: dependents overlap their prologue, they still wait for our completion
  for (int64_t w = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; w < total; w += int64_t(gridDim.x) * blockDim.x) {
    const int v = int(w % vec_per_head);
    const int64_t rest = w / vec_per_head;
    const int h = int(rest % num_heads);
    const int64_t i = rest / num_heads;
    const int b = batch_indices[i];
    const int pos = positions[i];
    const int page = kv_indices[kv_indptr[b] + pos / page_size];
    const int entry = pos % page_size;
    const int64_t off = int64_t(page) * c_sp + int64_t(entry) * c_sn + int64_t(h) * c_sh + v * VN;
    st16(k_cache + off, ld16(key + i * key_sn + h * key_sh + v * VN));
    st16(v_cache + off, ld16(value + i * val_sn + h * val_sh + v * VN));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
append_paged_mla_kernel(const T* __restrict__ ckv, const T* __restrict__ kpe, const int32_t* __restrict__ batch_indices,
                        const int32_t* __restrict__ positions, T* __restrict__ ckv_cache, T* __restrict__ kpe_cache,
                        const int32_t* __restrict__ kv_indices, const int32_t* __restrict__ kv_indptr, int64_t nnz,
                        int ckv_dim, int kpe_dim, int page_size, int64_t ckv_sn, int64_t kpe_sn, int64_t cc_sp,
                        int64_t cc_sn, int64_t kc_sp, int64_t kc_sn) {
  constexpr int VN = 16 / sizeof(T);
  const int vc = ckv_dim / VN, vk = kpe_dim / VN;
  const int64_t total = nnz * (vc + vk);
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();  // early trigger: dependents overlap their prologue, they still wait for our completion
  for (int64_t w = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; w < total; w += int64_t(gridDim.x) * blockDim.x) {
    const int v = int(w % (vc + vk));
    const int64_t i = w / (vc + vk);
    const int b = batch_indices[i];
    const int pos = positions[i];
    const int page = kv_indices[kv_indptr[b] + pos / page_size];
    const int entry = pos % page_size;
    if (v < vc) {
      st16(ckv_cache + int64_t(page) * cc_sp + int64_t(entry) * cc_sn + v * VN, ld16(ckv + i * ckv_sn + v * VN));
    } else {
      st16(kpe_cache + int64_t(page) * kc_sp + int64_t(entry) * kc_sn + (v - vc) * VN,
           ld16(kpe + i * kpe_sn + (v - vc) * VN));
    }
  }
}

__global__ void batch_indices_positions_kernel(const int32_t* __restrict__ append_indptr,
                                               const int32_t* __restrict__ seq_lens, int32_t* __restrict__ batch_indices,
                                               int32_t* __restrict__ positions, int batch) {
  const int b = blockIdx.x;
  if (b >= batch) return;
  const int s = append_indptr[b], e = append_indptr[b + 1];
  const int len = e - s;
  const int base = seq_lens[b] - len;
  for (int j = threadIdx.x; j < len; j += blockDim.x) {
    batch_indices[s + j] = b;
    positions[s + j] = base + j;
  }
}

}  // namespace

extern "C" int append_paged_kv_cache(void* key, void* value, void* batch_indices, void* positions, void* k_cache,
                                     void* v_cache, void* kv_indices, void* kv_indptr, int64_t nnz, int64_t num_heads,
                                     int64_t head_dim, int64_t page_size, int64_t key_sn, int64_t key_sh, int64_t val_sn,
                                     int64_t val_sh, int64_t c_sp, int64_t c_sn, int64_t c_sh, int64_t dtype,
                                     int64_t pdl, int64_t stream_) {
  if (nnz == 0) return 0;
  const int esz = dtype_size(dtype);
  FIB_CHECK(head_dim % (16 / esz) == 0, "append_paged_kv_cache: head_dim must be a multiple of the 16B vector width");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int64_t total = nnz * num_heads * (head_dim / (16 / esz));
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = int64_t(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  LaunchCfg lc(dim3((unsigned)blocks), dim3(256), 0, stream, pdl != 0);
#define FIB_APPEND(T)                                                                                              \
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, append_paged_kv_kernel<T>, (const T*)key, (const T*)value,            \
                                    (const int32_t*)batch_indices, (const int32_t*)positions, (T*)k_cache,         \
                                    (T*)v_cache, (const int32_t*)kv_indices, (const int32_t*)kv_indptr, nnz,       \
                                    (int)num_heads, (int)head_dim, (int)page_size, key_sn, key_sh, val_sn, val_sh, \
                                    c_sp, c_sn, c_sh))
  if (esz == 2) {
    FIB_APPEND(uint16_t);
  } else if (esz == 1) {
    FIB_APPEND(uint8_t);
  } else if (esz == 4) {
    FIB_APPEND(uint32_t);
  } else {
    return set_error("append_paged_kv_cache: unsupported dtype size");
  }
#undef FIB_APPEND
  return 0;
}

extern "C" int append_paged_mla_kv_cache(void* ckv, void* kpe, void* batch_indices, void* positions, void* ckv_cache,
                                         void* kpe_cache, void* kv_indices, void* kv_indptr, int64_t nnz,
                                         int64_t ckv_dim, int64_t kpe_dim, int64_t page_size, int64_t ckv_sn,
                                         int64_t kpe_sn, int64_t cc_sp, int64_t cc_sn, int64_t kc_sp, int64_t kc_sn,
                                         int64_t dtype, int64_t pdl, int64_t stream_) {
  if (nnz == 0) return 0;
  const int esz = dtype_size(dtype);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int64_t total = nnz * ((ckv_dim + kpe_dim) / (16 / esz));
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = int64_t(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  LaunchCfg lc(dim3((unsigned)blocks), dim3(256), 0, stream, pdl != 0);
#define FIB_APPEND_MLA(T)                                                                                       \
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, append_paged_mla_kernel<T>, (const T*)ckv, (const T*)kpe,          \
                                    (const int32_t*)batch_indices, (const int32_t*)positions, (T*)ckv_cache,    \
                                    (T*)kpe_cache, (const int32_t*)kv_indices, (const int32_t*)kv_indptr, nnz,  \
                                    (int)ckv_dim, (int)kpe_dim, (int)page_size, ckv_sn, kpe_sn, cc_sp, cc_sn,   \
                                    kc_sp, kc_sn))
  if (esz == 2) {
    FIB_APPEND_MLA(uint16_t);
  } else if (esz == 1) {
    FIB_APPEND_MLA(uint8_t);
  } else {
    return set_error("append_paged_mla_kv_cache: unsupported dtype size");
  }
#undef FIB_APPEND_MLA
  return 0;
}

extern "C" int get_batch_indices_positions(void* append_indptr, void* seq_lens, void* batch_indices, void* positions,
                                           int64_t batch, int64_t stream_) {
  if (batch == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  batch_indices_positions_kernel<<<(unsigned)batch, 128, 0, stream>>>((const int32_t*)append_indptr,
                                                                      (const int32_t*)seq_lens, (int32_t*)batch_indices,
                                                                      (int32_t*)positions, (int)batch);
  FIB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Row gather: dst[i, :] = idx[i] >= 0 ? src[idx[i], :] : 0   (rows of `row_bytes` bytes, multiple of 16).  The compaction
// step of sparse (top-k) MLA: the selected KV tokens of every query are packed into a dense per-query cache that the tcgen05
// MLA kernel streams with full-size TMA boxes (reference: sparse_mla_top_k of trtllm_batch_decode_with_kv_cache_mla,
// flashinfer/mla/_core.py:631-940, where the closed kernel gathers by token index).
// ---------------------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint8_t* __restrict__ src, const int32_t* __restrict__ idx,
                                                          uint8_t* __restrict__ dst, int64_t n_rows, int64_t src_rows, int vecs,
                                                          int64_t src_pitch, int64_t dst_pitch) {
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();
  const int64_t total = n_rows * vecs;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / vecs;
    const int v = int(i - r * vecs);
    const int32_t s = idx[r];
    int4 val = make_int4(0, 0, 0, 0);
    if (s >= 0 && s < src_rows) val = __ldg(reinterpret_cast<const int4*>(src + int64_t(s) * src_pitch) + v);
    reinterpret_cast<int4*>(dst + r * dst_pitch)[v] = val;
  }
}
}  // namespace

extern "C" int gather_rows(void* src, void* idx, void* dst, int64_t n_rows, int64_t src_rows, int64_t row_bytes, int64_t src_pitch,
                           int64_t dst_pitch, int64_t pdl, int64_t stream_) {
  FIB_CHECK(row_bytes % 16 == 0 && src_pitch % 16 == 0 && dst_pitch % 16 == 0, "gather_rows: row size / pitches must be multiples of 16 bytes");
  FIB_CHECK((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "gather_rows: 16-byte aligned tensors");
  if (n_rows == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int vecs = int(row_bytes / 16);
  int64_t blocks = (n_rows * vecs + 255) / 256;
  const int64_t cap = int64_t(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  LaunchCfg lc(dim3((unsigned)blocks), dim3(256), 0, stream, pdl != 0);
  FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, gather_rows_kernel, (const uint8_t*)src, (const int32_t*)idx, (uint8_t*)dst, n_rows,
                                    src_rows, vecs, src_pitch, dst_pitch));
  return 0;
}
