// MoE routing + token permutation utilities (SIMT, warp-per-token).
//
// Parity: reference routing kernels csrc/fused_moe/trtllm_backend/trtllm_fused_moe_routing_{deepseek,llama4,custom}.cu,
// csrc/fused_moe/noAuxTcKernels.cu (fused_topk_deepseek) and the permute / finalize kernels of
// csrc/fused_moe/trtllm_backend/trtllm_fused_moe_dev_kernel.cu:635-900, csrc/nv_internal/.../moeUtils.cu.
// Routing methods (flashinfer/tllm_enums.py:6-27):
//   0 Default (softmax -> top-k)        1 Renormalize (top-k -> softmax)     2 DeepSeekV3 (sigmoid + bias, group top-2
//   sums -> top groups -> top-k, normalise * scale)   3 Llama4 (top-1 -> sigmoid)   4 RenormalizeNaive (softmax ->
//   top-k -> renormalise)   5 TopK (raw)   6 SigmoidRenorm   7 MiniMax2 (sigmoid + bias -> top-k -> sum-normalise)
//   8 Sigmoid (no renormalisation)
#include <fib200/common.cuh>
#include <fib200/ptx.cuh>

using namespace fib200;

FIB_EXPORT_LAST_ERROR()

namespace {

constexpr int kMaxPerLane = 16;  // up to 512 experts
constexpr int kMaxTopK = 32;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// One warp per token.  logits [T, E] (f32 or bf16), bias [E] optional.
// kSlots = register slots per lane (experts / 32, rounded up to a power of two): every unrolled loop is sized by it.
template <typename TL, int kSlots>
__global__ void __launch_bounds__(256)
routing_kernel(const TL* __restrict__ logits, const float* __restrict__ bias, int32_t* __restrict__ topk_ids,
               float* __restrict__ topk_w, int T, int E, int K, int method, int n_group, int topk_group,
               float routed_scale, int norm_topk_prob) {
  const int lane = threadIdx.x & 31;
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= T) return;
  ptx::grid_dep_wait();
  const int per = (E + 31) / 32;
  float score[kSlots];   // value used for selection
  float orig[kSlots];    // value used for the output weight
  // expert e lives in lane e % 32, slot e / 32
  float mx = -INFINITY;
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    const int e = s * 32 + lane;
    float x = (s < per && e < E) ? to_f32(logits[int64_t(tok) * E + e]) : -INFINITY;
    score[s] = x;
    mx = fmaxf(mx, x);
  }
  mx = warp_reduce_max(mx);
  const bool softmax_first = (method == 0 || method == 4);
  const bool sigmoid_first = (method == 2 || method == 6 || method == 7 || method == 8);
  if (softmax_first) {
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      if (s < per) {
        score[s] = (score[s] == -INFINITY) ? 0.f : __expf(score[s] - mx);
        sum += score[s];
      }
    }
    sum = warp_reduce_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int e = s * 32 + lane;
      score[s] = (s < per && e < E) ? score[s] * inv : -INFINITY;
      orig[s] = score[s];
    }
  } else if (sigmoid_first) {
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int e = s * 32 + lane;
      if (s < per && e < E) {
        const float sg = sigmoidf_(score[s]);
        orig[s] = sg;
        score[s] = sg + ((bias && (method == 2 || method == 7)) ? bias[e] : 0.f);
      } else {
        orig[s] = 0.f;
        score[s] = -INFINITY;
      }
    }
  } else {
#pragma unroll
    for (int s = 0; s < kSlots; ++s) orig[s] = score[s];
  }

  // DeepSeek-V3 group limiting: keep only experts of the `topk_group` best groups (score = sum of top-2 in group)
  if (method == 2 && n_group > 1) {
    const int gsz = E / n_group;
    // group scores are computed redundantly by lane g (n_group <= 32)
    float gscore = -INFINITY;
    // gather every expert score into smem-free fashion: use shuffles per slot
    float top1 = -INFINITY, top2 = -INFINITY;
    if (gsz % 32 == 0) {
      // fast path: a group is a whole number of register slots -> per-lane top-2 of the group's slots, then a
      // butterfly merge of (top1, top2) pairs; lane g keeps the score of group g
      const int spg = gsz / 32;
#pragma unroll 4
      for (int g = 0; g < n_group; ++g) {
        float a1 = -INFINITY, a2 = -INFINITY;
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          if (s >= g * spg && s < (g + 1) * spg) {
            const float v = score[s];
            if (v > a1) {
              a2 = a1;
              a1 = v;
            } else if (v > a2) {
              a2 = v;
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float b1 = __shfl_xor_sync(0xffffffffu, a1, o), b2 = __shfl_xor_sync(0xffffffffu, a2, o);
          const float hi = fmaxf(a1, b1);
          const float lo = fmaxf(fminf(a1, b1), fmaxf(a2, b2));
          a1 = hi;
          a2 = lo;
        }
        if (lane == g) {
          top1 = a1;
          top2 = a2;
        }
      }
    } else {
    for (int s = 0; s < per; ++s) {
      for (int l = 0; l < 32; ++l) {
        const float v = __shfl_sync(0xffffffffu, score[s], l);
        const int e = s * 32 + l;
        if (e < E && e / gsz == lane) {
          if (v > top1) {
            top2 = top1;
            top1 = v;
          } else if (v > top2) {
            top2 = v;
          }
        }
      }
    }
    }
    if (lane < n_group) gscore = top1 + top2;
    // select topk_group groups: rank of my group
    int rank = 0;
    for (int l = 0; l < n_group; ++l) {
      const float v = __shfl_sync(0xffffffffu, gscore, l);
      if (v > gscore || (v == gscore && l < lane)) ++rank;
    }
    const unsigned keep_mask = __ballot_sync(0xffffffffu, lane < n_group && rank < topk_group);
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int e = s * 32 + lane;
      if (s < per && e < E) {
        if (!((keep_mask >> (e / gsz)) & 1u)) score[s] = -INFINITY;
      }
    }
  }

  // iterative top-K (K <= 32): argmax over the warp, ties -> smaller expert id.  Selection k is kept by lane k.
  float my_w = 0.f;
  int my_id = 0;
  float wsum = 0.f;
  for (int k = 0; k < K; ++k) {
    float bv = -INFINITY;
    int be = 1 << 30;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int e = s * 32 + lane;
      if (s < per && e < E && (score[s] > bv || (score[s] == bv && e < be))) {
        bv = score[s];
        be = e;
      }
    }
    {
      // warp argmax with two redux.sync ops: max of the order-preserving integer image of the score, then the smallest
      // expert id among the lanes that hold it (ties -> smaller id, like torch.topk on the sorted order)
      const uint32_t bits = __float_as_uint(bv);
      const uint32_t key = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
      const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
      be = (int)__reduce_min_sync(0xffffffffu, key == kmax ? (unsigned)be : 0x7fffffffu);
      const uint32_t b2 = (kmax & 0x80000000u) ? (kmax & 0x7fffffffu) : ~kmax;
      bv = __uint_as_float(b2);
    }
    // owner lane fetches the output weight and removes the expert
    float w = 0.f;
    if ((be & 31) == lane && be < E) {
      const int s = be >> 5;
#pragma unroll
      for (int ss = 0; ss < kSlots; ++ss)
        if (ss == s) {
          w = orig[ss];
          score[ss] = -INFINITY;
        }
    }
    w = __shfl_sync(0xffffffffu, w, be & 31);
    if (lane == k) {
      my_w = w;
      my_id = be;
    }
    wsum += w;
  }
  const bool have = lane < K;
  // post-processing of the selected weights (lane k owns selection k)
  if (method == 1) {  // top-k -> softmax over the selected logits
    const float m2 = warp_reduce_max(have ? my_w : -INFINITY);
    const float ex = have ? __expf(my_w - m2) : 0.f;
    const float s2 = warp_reduce_sum(ex);
    my_w = ex / s2;
  } else if (method == 3) {  // llama4: sigmoid of the selected logit(s)
    my_w = sigmoidf_(my_w);
  } else if (method == 2 || method == 7) {
    const float inv = (norm_topk_prob || method == 7) ? 1.f / (wsum + 1e-20f) : 1.f;
    my_w = my_w * inv * routed_scale;
  } else if (method == 4 || method == 6) {
    my_w *= 1.f / (wsum + 1e-20f);
  }
  if (have) {
    topk_ids[int64_t(tok) * K + lane] = my_id;
    topk_w[int64_t(tok) * K + lane] = my_w;
  }
  ptx::grid_dep_launch();
}

// ------------------------------------------------------------------ sort / permutation (single CTA)
// Builds the expert-grouped, tile-padded permutation.
//   counts[e], offsets[e] (padded to `tile`), expanded_to_permuted[T*K] (-1 for non-local experts),
//   permuted_to_token[P] (-1 padding), tile_expert[P / tile] (-1 unused), meta[0] = num tiles, meta[1] = padded rows
// Three small kernels (deterministic: rows of one expert keep ascending expanded-index order):
//   count   : grid = chunks of 1024 expanded entries; per-chunk smem histogram -> chunk_hist[chunk][e]
//   scan    : one CTA; per expert exclusive prefix over the chunks, padded expert offsets, tile_expert, meta
//   scatter : grid = chunks; rank inside the chunk via warp match + warps taking turns on a smem cursor
constexpr int kSortChunk = 1024;

__global__ void __launch_bounds__(kSortChunk)
moe_count_kernel(const int32_t* __restrict__ topk_ids, int n, int local_offset, int local_num,
                 int32_t* __restrict__ chunk_hist) {
  extern __shared__ int sm[];
  ptx::grid_dep_wait();
  for (int i = threadIdx.x; i < local_num; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  const int i = blockIdx.x * kSortChunk + threadIdx.x;
  if (i < n) {
    const int e = topk_ids[i] - local_offset;
    if (e >= 0 && e < local_num) atomicAdd(&sm[e], 1);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < local_num; e += blockDim.x) chunk_hist[int64_t(blockIdx.x) * local_num + e] = sm[e];
  ptx::grid_dep_launch();
}

__global__ void __launch_bounds__(1024)
moe_scan_kernel(int32_t* __restrict__ chunk_hist, int nchunks, int local_num, int tile, int max_rows,
                int32_t* __restrict__ permuted_to_token, int32_t* __restrict__ tile_expert,
                int32_t* __restrict__ expert_offsets, int32_t* __restrict__ meta) {
  extern __shared__ int sm[];
  int* cnt = sm;              // [local_num]
  int* off = sm + local_num;  // [local_num + 1]
  __shared__ int warp_tot[32];
  ptx::grid_dep_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // one thread per expert (strided): exclusive prefix over the chunks, independent loads in flight
  for (int e = threadIdx.x; e < local_num; e += blockDim.x) {
    int acc = 0;
    int c = 0;
    for (; c + 4 <= nchunks; c += 4) {
      const int v0 = chunk_hist[int64_t(c) * local_num + e], v1 = chunk_hist[int64_t(c + 1) * local_num + e];
      const int v2 = chunk_hist[int64_t(c + 2) * local_num + e], v3 = chunk_hist[int64_t(c + 3) * local_num + e];
      chunk_hist[int64_t(c) * local_num + e] = acc;
      chunk_hist[int64_t(c + 1) * local_num + e] = acc + v0;
      chunk_hist[int64_t(c + 2) * local_num + e] = acc + v0 + v1;
      chunk_hist[int64_t(c + 3) * local_num + e] = acc + v0 + v1 + v2;
      acc += v0 + v1 + v2 + v3;
    }
    for (; c < nchunks; ++c) {
      const int v = chunk_hist[int64_t(c) * local_num + e];
      chunk_hist[int64_t(c) * local_num + e] = acc;
      acc += v;
    }
    cnt[e] = acc;
  }
  __syncthreads();
  // block-wide exclusive scan of the tile-padded counts (local_num <= 1024 per pass)
  int base = 0;
  for (int e0 = 0; e0 < local_num; e0 += blockDim.x) {
    const int e = e0 + threadIdx.x;
    const int padded = e < local_num ? (cnt[e] + tile - 1) / tile * tile : 0;
    int incl = padded;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) wbase += warp_tot[w];
      total += warp_tot[w];
    }
    if (e < local_num) off[e] = base + wbase + incl - padded;
    base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    off[local_num] = base;
    meta[0] = base / tile;
    meta[1] = base;
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= local_num; i += blockDim.x) expert_offsets[i] = off[i];
  for (int i = off[local_num] / tile + threadIdx.x; i < max_rows / tile; i += blockDim.x) tile_expert[i] = -1;
  for (int i = off[local_num] + threadIdx.x; i < max_rows; i += blockDim.x) permuted_to_token[i] = -1;  // dead tail rows
  // per expert: tile -> expert map and the -1 padding rows at the tail of its last tile
  for (int e = warp; e < local_num; e += blockDim.x >> 5) {
    for (int r = off[e] + lane * tile; r < off[e + 1]; r += 32 * tile) tile_expert[r / tile] = e;
    for (int r = off[e] + cnt[e] + lane; r < off[e + 1]; r += 32) permuted_to_token[r] = -1;
  }
  ptx::grid_dep_launch();
}

__global__ void __launch_bounds__(kSortChunk)
moe_scatter_kernel(const int32_t* __restrict__ topk_ids, int n, int K, int local_offset, int local_num,
                   const int32_t* __restrict__ chunk_base, const int32_t* __restrict__ expert_offsets,
                   int32_t* __restrict__ expanded_to_permuted, int32_t* __restrict__ permuted_to_token) {
  extern __shared__ int sm[];  // cursor[local_num]
  ptx::grid_dep_wait();
  for (int e = threadIdx.x; e < local_num; e += blockDim.x)
    sm[e] = expert_offsets[e] + chunk_base[int64_t(blockIdx.x) * local_num + e];
  __syncthreads();
  const int i = blockIdx.x * kSortChunk + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int e = -1;
  if (i < n) {
    e = topk_ids[i] - local_offset;
    if (e < 0 || e >= local_num) e = -1;
  }
  const uint32_t peers = __match_any_sync(0xffffffffu, e);
  const int rank = __popc(peers & ((1u << lane) - 1));
  const bool leader = rank == 0;
  int base = 0;
  for (int w = 0; w < kSortChunk / 32; ++w) {
    if (w == warp && e >= 0 && leader) {
      base = sm[e];
      sm[e] = base + __popc(peers);
    }
    __syncthreads();
  }
  base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
  if (i < n) {
    if (e >= 0) {
      const int pos = base + rank;
      expanded_to_permuted[i] = pos;
      permuted_to_token[pos] = i / K;
    } else {
      expanded_to_permuted[i] = -1;
    }
  }
  ptx::grid_dep_launch();
}

// Small problems (T * K <= 32 K entries, 32 * local_num ints of smem): the whole sort in ONE CTA, still deterministic.
// Warp w owns the contiguous entry range [w * per, (w + 1) * per) and walks it 32 entries at a time, so every step is
// warp-local (match.any + a warp-private histogram row); the only block-wide steps are the expert scan in the middle.
template <int ITERS>  // 32-entry groups per warp (entries are preloaded: one round of global latency instead of ITERS)
__global__ void __launch_bounds__(1024)
moe_sort_small_kernel(const int32_t* __restrict__ topk_ids, int n, int K, int local_offset, int local_num, int tile,
                      int max_rows, int32_t* __restrict__ expanded_to_permuted, int32_t* __restrict__ permuted_to_token,
                      int32_t* __restrict__ tile_expert, int32_t* __restrict__ expert_offsets, int32_t* __restrict__ meta) {
  extern __shared__ int sm[];
  int* hist = sm;                          // [32][local_num] per-warp counts, then per-warp cursors
  int* cnt = sm + 32 * local_num;          // [local_num]
  int* off = cnt + local_num;              // [local_num + 1]
  __shared__ int warp_tot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * local_num; i += blockDim.x) hist[i] = 0;
  ptx::grid_dep_wait();
  __syncthreads();
  const int per = ITERS * 32;  // entries per warp
  const int lo = warp * per, hi = min(n, lo + per);
  int* myh = hist + warp * local_num;
  int ev[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int i = lo + it * 32 + lane;
    int e = -1;
    if (i < hi) {
      e = topk_ids[i] - local_offset;
      if (e < 0 || e >= local_num) e = -1;
    }
    ev[it] = e;
  }
  // rank / count among the 32 lanes by an all-to-all shuffle sweep (match.any is a long-latency serialising op here)
  auto peers_of = [&](int e, int& rank, int& cnt, int& first) {
    rank = 0;
    cnt = 0;
    first = 32;
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const bool same = __shfl_sync(0xffffffffu, e, l) == e;
      cnt += same;
      rank += (same && l < lane);
      if (same && first == 32) first = l;
    }
  };
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int e = ev[it];
    int rank, cnt, first;
    peers_of(e, rank, cnt, first);
    if (e >= 0 && rank == 0) myh[e] += cnt;
    __syncwarp();
  }
  __syncthreads();
  // per expert: exclusive prefix over the 32 warps (in place), total count
  for (int e = threadIdx.x; e < local_num; e += blockDim.x) {
    int acc = 0;
#pragma unroll 8
    for (int w = 0; w < 32; ++w) {
      const int v = hist[w * local_num + e];
      hist[w * local_num + e] = acc;
      acc += v;
    }
    cnt[e] = acc;
  }
  __syncthreads();
  int base = 0;
  for (int e0 = 0; e0 < local_num; e0 += blockDim.x) {
    const int e = e0 + threadIdx.x;
    const int padded = e < local_num ? (cnt[e] + tile - 1) / tile * tile : 0;
    int incl = padded;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) wbase += warp_tot[w];
      total += warp_tot[w];
    }
    if (e < local_num) off[e] = base + wbase + incl - padded;
    base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    off[local_num] = base;
    meta[0] = base / tile;
    meta[1] = base;
  }
  __syncthreads();
  ptx::grid_dep_launch();
  for (int i = threadIdx.x; i <= local_num; i += blockDim.x) expert_offsets[i] = off[i];
  for (int i = off[local_num] / tile + threadIdx.x; i < max_rows / tile; i += blockDim.x) tile_expert[i] = -1;
  for (int i = off[local_num] + threadIdx.x; i < max_rows; i += blockDim.x) permuted_to_token[i] = -1;
  for (int e = warp; e < local_num; e += blockDim.x >> 5) {
    for (int r = off[e] + lane * tile; r < off[e + 1]; r += 32 * tile) tile_expert[r / tile] = e;
    for (int r = off[e] + cnt[e] + lane; r < off[e + 1]; r += 32) permuted_to_token[r] = -1;
  }
  // scatter: same walk, positions = expert offset + warp prefix + running cursor + rank inside the 32-entry group
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int i = lo + it * 32 + lane;
    const int e = ev[it];
    int rank, cnt, first;
    peers_of(e, rank, cnt, first);
    int b = 0;
    if (e >= 0 && rank == 0) {
      b = myh[e];
      myh[e] = b + cnt;
    }
    b = __shfl_sync(0xffffffffu, b, first);
    if (i < hi) {
      if (e >= 0) {
        const int pos = off[e] + b + rank;
        expanded_to_permuted[i] = pos;
        permuted_to_token[pos] = i / K;
      } else {
        expanded_to_permuted[i] = -1;
      }
    }
    __syncwarp();
  }
}

__device__ __forceinline__ int ld_dsmem_i32(uint32_t addr) {
  int v;
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// Mid-size problems (T * K <= 16 K entries): the three phases in ONE launch on a thread-block cluster - CTA c owns the 1024
// entries [1024 c, 1024 c + 1024), the per-CTA histograms are combined through distributed shared memory, every CTA then
// knows the padded expert offsets and its own base inside every expert and scatters its chunk (same deterministic order as
// the 3-kernel path: rows of an expert ascend with the expanded index).  A single SM is NOT enough for this: match.any and
// the shuffle sweeps that could replace it are issue / unit bound (measured 26 - 34 us for 8 K entries on one SM).
__global__ void __launch_bounds__(kSortChunk)
moe_sort_cluster_kernel(const int32_t* __restrict__ topk_ids, int n, int K, int local_offset, int local_num, int tile,
                        int max_rows, int32_t* __restrict__ expanded_to_permuted, int32_t* __restrict__ permuted_to_token,
                        int32_t* __restrict__ tile_expert, int32_t* __restrict__ expert_offsets, int32_t* __restrict__ meta) {
  extern __shared__ int sm[];
  int* hist = sm;                   // [local_num] my chunk's histogram, later my running cursor
  int* cnt = sm + local_num;        // [local_num] totals over the cluster
  int* off = cnt + local_num;       // [local_num + 1] padded expert offsets
  __shared__ int warp_tot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int crank = int(ptx::cluster_ctarank()), csize = int(ptx::cluster_nctarank());
  for (int i = threadIdx.x; i < local_num; i += blockDim.x) hist[i] = 0;
  ptx::grid_dep_wait();
  __syncthreads();
  const int i = crank * kSortChunk + threadIdx.x;
  int e = -1;
  if (i < n) {
    e = topk_ids[i] - local_offset;
    if (e < 0 || e >= local_num) e = -1;
  }
  const uint32_t peers = __match_any_sync(0xffffffffu, e);
  const int rank = __popc(peers & ((1u << lane) - 1));
  if (e >= 0 && rank == 0) atomicAdd(&hist[e], __popc(peers));
  ptx::cluster_sync();  // all histograms complete and visible cluster-wide
  // totals and my prefix (CTAs of lower rank) for every expert
  for (int x = threadIdx.x; x < local_num; x += blockDim.x) {
    int tot = 0, before = 0;
    for (int c = 0; c < csize; ++c) {
      const int v = c == crank ? hist[x] : ld_dsmem_i32(ptx::mapa(ptx::smem_u32(&hist[x]), uint32_t(c)));
      tot += v;
      if (c < crank) before += v;
    }
    cnt[x] = tot;
    off[x] = before;  // parked here until the padded offsets are known
  }
  ptx::cluster_sync();  // every remote read of `hist` is done: it may be overwritten (and CTAs may exit later on)
  // block-wide exclusive scan of the tile-padded totals (every CTA computes the same offsets)
  int base = 0;
  for (int e0 = 0; e0 < local_num; e0 += blockDim.x) {
    const int x = e0 + threadIdx.x;
    const int padded = x < local_num ? (cnt[x] + tile - 1) / tile * tile : 0;
    int incl = padded;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < warp) wbase += warp_tot[w];
      total += warp_tot[w];
    }
    if (x < local_num) {
      const int o_x = base + wbase + incl - padded;
      hist[x] = o_x + off[x];  // my cursor: expert offset + rows of lower-rank CTAs
      off[x] = o_x;
    }
    base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[local_num] = base;
  __syncthreads();
  ptx::grid_dep_launch();
  // bookkeeping, spread over the cluster: expert x is written by CTA x % csize, the dead tail by everybody
  if (crank == 0 && threadIdx.x == 0) {
    meta[0] = base / tile;
    meta[1] = base;
  }
  for (int x = crank * int(blockDim.x) + threadIdx.x; x <= local_num; x += csize * blockDim.x) expert_offsets[x] = off[x];
  for (int x = base / tile + crank * int(blockDim.x) + threadIdx.x; x < max_rows / tile; x += csize * blockDim.x) tile_expert[x] = -1;
  for (int x = base + crank * int(blockDim.x) + threadIdx.x; x < max_rows; x += csize * blockDim.x) permuted_to_token[x] = -1;
  for (int x = crank + warp * csize; x < local_num; x += csize * (blockDim.x >> 5)) {
    for (int r = off[x] + lane * tile; r < off[x + 1]; r += 32 * tile) tile_expert[r / tile] = x;
    for (int r = off[x] + cnt[x] + lane; r < off[x + 1]; r += 32) permuted_to_token[r] = -1;
  }
  // scatter my chunk: rank inside the 32-entry group, warps take turns on the shared cursor (deterministic)
  int b = 0;
  for (int w = 0; w < kSortChunk / 32; ++w) {
    if (w == warp && e >= 0 && rank == 0) {
      b = hist[e];
      hist[e] = b + __popc(peers);
    }
    __syncthreads();
  }
  b = __shfl_sync(0xffffffffu, b, __ffs(peers) - 1);
  if (i < n) {
    if (e >= 0) {
      const int pos = b + rank;
      expanded_to_permuted[i] = pos;
      permuted_to_token[pos] = i / K;
    } else {
      expanded_to_permuted[i] = -1;
    }
  }
}

// permuted_x[p, :] = x[token(p), :]  (zero rows for padding)
template <typename T>
__global__ void __launch_bounds__(256)
moe_gather_kernel(const T* __restrict__ x, T* __restrict__ out, const int32_t* __restrict__ permuted_to_token,
                  const int32_t* __restrict__ meta, int64_t hidden, int64_t x_stride, int zero_pad,
                  const int32_t* __restrict__ row_list, int64_t n_list, int list_div) {
  constexpr int VN = 16 / sizeof(T);
  ptx::grid_dep_wait();
  const int64_t vec = hidden / VN;
  if (row_list) {
    // live rows only: entry j of the expanded (token, k) list lands in permuted row row_list[j]; padding rows stay
    // uninitialised (GEMM rows are independent and nothing reads the padding rows of the result)
    const int64_t total = n_list * vec;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
      int64_t j, v;
      fast_divmod(i, vec, j, v);
      const int r = row_list[j];
      if (r < 0) continue;
      st16(out + int64_t(r) * hidden + v * VN, ld16(x + int64_t(uint32_t(j) / uint32_t(list_div)) * x_stride + v * VN));
    }
    ptx::grid_dep_launch();
    return;
  }
  const int rows = meta[1];
  const int64_t total = int64_t(rows) * vec;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / vec, v = i % vec;
    const int tok = permuted_to_token[r];
    Vec16<T> val;
    if (tok >= 0) {
      val = ld16(x + int64_t(tok) * x_stride + v * VN);
    } else {
      if (!zero_pad) continue;  // consumers that skip padding rows do not need them initialised
      *reinterpret_cast<int4*>(&val) = make_int4(0, 0, 0, 0);
    }
    st16(out + r * hidden + v * VN, val);
  }
  ptx::grid_dep_launch();
}

// out[t, :] = sum_k w[t,k] * y[perm(t,k), :]   (skips non-local experts)
template <typename T>
__global__ void __launch_bounds__(256)
moe_finalize_kernel(const T* __restrict__ y, T* __restrict__ out, const int32_t* __restrict__ expanded_to_permuted,
                    const float* __restrict__ topk_w, int Tn, int K, int64_t hidden, int accumulate) {
  constexpr int VN = 16 / sizeof(T);
  ptx::grid_dep_wait();
  const int64_t vec = hidden / VN;
  const int64_t total = int64_t(Tn) * vec;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int64_t t, v;
    fast_divmod(i, vec, t, v);
    float acc[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[e] = 0.f;
    if (accumulate) {
      const Vec16<T> o = ld16(out + t * hidden + v * VN);
#pragma unroll
      for (int e = 0; e < VN; ++e) acc[e] = to_f32(o.v[e]);
    }
    for (int k = 0; k < K; ++k) {
      const int p = expanded_to_permuted[t * K + k];
      if (p < 0) continue;
      const float w = topk_w[t * K + k];
      const Vec16<T> val = ld16(y + int64_t(p) * hidden + v * VN);
#pragma unroll
      for (int e = 0; e < VN; ++e) acc[e] += w * to_f32(val.v[e]);
    }
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < VN; ++e) o.v[e] = from_f32<T>(acc[e]);
    st16(out + t * hidden + v * VN, o);
  }
  ptx::grid_dep_launch();
}

inline int grid_for(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = int64_t(num_sms()) * 16;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

}  // namespace

extern "C" int moe_routing(void* logits, void* bias, void* topk_ids, void* topk_w, int64_t T, int64_t E, int64_t K,
                           int64_t method, int64_t n_group, int64_t topk_group, double routed_scale,
                           int64_t norm_topk_prob, int64_t logits_dtype, int64_t pdl, int64_t stream_) {
  FIB_CHECK(E <= 32 * kMaxPerLane, "moe_routing: at most 512 experts");
  FIB_CHECK(E >= 1, "moe_routing: no experts");
  FIB_CHECK(K >= 1 && K <= kMaxTopK, "moe_routing: top_k must be in [1,32]");
  FIB_CHECK(method != 2 || n_group <= 32, "moe_routing: n_group must be <= 32");
  if (T == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  LaunchCfg lc(dim3((unsigned)((T + 7) / 8)), dim3(256), 0, stream, pdl != 0);
  auto go = [&](auto kern, auto* lg) -> int {
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, lg, (const float*)bias, (int32_t*)topk_ids, (float*)topk_w, (int)T, (int)E,
                                      (int)K, (int)method, (int)(n_group > 0 ? n_group : 1),
                                      (int)(topk_group > 0 ? topk_group : 1), (float)routed_scale, (int)norm_topk_prob));
    return 0;
  };
  const int slots = (int)((E + 31) / 32);
#define FIB_ROUTE(TL_)                                                        \
  {                                                                           \
    const TL_* lg = (const TL_*)logits;                                       \
    if (slots <= 1) return go(routing_kernel<TL_, 1>, lg);                    \
    if (slots <= 2) return go(routing_kernel<TL_, 2>, lg);                    \
    if (slots <= 4) return go(routing_kernel<TL_, 4>, lg);                    \
    if (slots <= 8) return go(routing_kernel<TL_, 8>, lg);                    \
    return go(routing_kernel<TL_, 16>, lg);                                   \
  }
  if (logits_dtype == kF32) FIB_ROUTE(float)
  if (logits_dtype == kBF16) FIB_ROUTE(__nv_bfloat16)
#undef FIB_ROUTE
  return set_error("moe_routing: logits must be float32 or bfloat16");
}

// workspace: int32 [ceil(T*K / 1024) * local_num]
extern "C" int moe_sort(void* topk_ids, int64_t T, int64_t K, int64_t E, int64_t local_offset, int64_t local_num,
                        int64_t tile, int64_t max_rows, void* expanded_to_permuted, void* permuted_to_token,
                        void* tile_expert, void* expert_offsets, void* meta, void* workspace, int64_t pdl,
                        int64_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int n = (int)(T * K);
  const int nchunks = n > 0 ? (n + kSortChunk - 1) / kSortChunk : 0;
  (void)E;
  static const int sort_mode = [] {
    const char* v = getenv("FIB200_MOE_SORT");
    return v ? atoi(v) : 0;  // 0 auto, 1 single CTA, 2 cluster, 3 three kernels
  }();
  if (n > 0 && nchunks <= 16 && local_num <= 4096 && (sort_mode == 0 || sort_mode == 2)) {
    // one launch, one cluster of ceil(n / 1024) CTAs (clusters above 8 need the non-portable opt-in)
    static bool np_set = false;
    if (!np_set) {
      FIB_CUDA_CHECK(cudaFuncSetAttribute(moe_sort_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      np_set = true;
    }
    LaunchCfg lc(dim3(nchunks), dim3(kSortChunk), (size_t(3) * local_num + 1) * sizeof(int), stream, pdl != 0, nchunks);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, moe_sort_cluster_kernel, (const int32_t*)topk_ids, n, (int)K, (int)local_offset,
                                      (int)local_num, (int)tile, (int)max_rows, (int32_t*)expanded_to_permuted,
                                      (int32_t*)permuted_to_token, (int32_t*)tile_expert, (int32_t*)expert_offsets,
                                      (int32_t*)meta));
    return 0;
  }
  const size_t small_smem = (size_t(34) * local_num + 1) * sizeof(int);
  if (n > 0 && n <= 32768 && small_smem <= 48 * 1024 && sort_mode == 1) {
    LaunchCfg lc(dim3(1), dim3(1024), small_smem, stream, pdl != 0);
    const int iters = (n + 1023) / 1024;  // 32-entry groups per warp
    auto go = [&](auto kern) -> int {
      FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, kern, (const int32_t*)topk_ids, n, (int)K, (int)local_offset, (int)local_num,
                                        (int)tile, (int)max_rows, (int32_t*)expanded_to_permuted, (int32_t*)permuted_to_token,
                                        (int32_t*)tile_expert, (int32_t*)expert_offsets, (int32_t*)meta));
      return 0;
    };
    if (iters <= 1) return go(moe_sort_small_kernel<1>);
    if (iters <= 2) return go(moe_sort_small_kernel<2>);
    if (iters <= 4) return go(moe_sort_small_kernel<4>);
    if (iters <= 8) return go(moe_sort_small_kernel<8>);
    if (iters <= 16) return go(moe_sort_small_kernel<16>);
    return go(moe_sort_small_kernel<32>);
  }
  if (nchunks > 0) {
    LaunchCfg lc(dim3(nchunks), dim3(kSortChunk), local_num * sizeof(int), stream, pdl != 0);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, moe_count_kernel, (const int32_t*)topk_ids, n, (int)local_offset,
                                      (int)local_num, (int32_t*)workspace));
  }
  {
    LaunchCfg lc(dim3(1), dim3(1024), (2 * local_num + 1) * sizeof(int), stream, pdl != 0);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, moe_scan_kernel, (int32_t*)workspace, nchunks, (int)local_num, (int)tile,
                                      (int)max_rows, (int32_t*)permuted_to_token, (int32_t*)tile_expert,
                                      (int32_t*)expert_offsets, (int32_t*)meta));
  }
  if (nchunks > 0) {
    LaunchCfg lc(dim3(nchunks), dim3(kSortChunk), local_num * sizeof(int), stream, pdl != 0);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, moe_scatter_kernel, (const int32_t*)topk_ids, n, (int)K,
                                      (int)local_offset, (int)local_num, (const int32_t*)workspace,
                                      (const int32_t*)expert_offsets, (int32_t*)expanded_to_permuted,
                                      (int32_t*)permuted_to_token));
  }
  return 0;
}

extern "C" int moe_gather(void* x, void* out, void* permuted_to_token, void* meta, int64_t max_rows, int64_t hidden,
                          int64_t x_stride, int64_t zero_pad, void* row_list, int64_t n_list, int64_t list_div, int64_t dtype,
                          int64_t pdl, int64_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  return FIB_DISPATCH_HALF(dtype, T, [&]() -> int {
    LaunchCfg lc(dim3(grid_for((row_list ? n_list : max_rows) * (hidden / 8))), dim3(256), 0, stream, pdl != 0);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, moe_gather_kernel<T>, (const T*)x, (T*)out,
                                      (const int32_t*)permuted_to_token, (const int32_t*)meta, hidden, x_stride, (int)zero_pad,
                                      (const int32_t*)row_list, n_list, (int)(list_div > 0 ? list_div : 1)));
    return 0;
  });
}

extern "C" int moe_finalize(void* y, void* out, void* expanded_to_permuted, void* topk_w, int64_t T, int64_t K,
                            int64_t hidden, int64_t accumulate, int64_t dtype, int64_t pdl, int64_t stream_) {
  if (T == 0) return 0;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  return FIB_DISPATCH_HALF(dtype, Tt, [&]() -> int {
    LaunchCfg lc(dim3(grid_for(T * (hidden / 8))), dim3(256), 0, stream, pdl != 0);
    FIB_CUDA_CHECK(cudaLaunchKernelEx(&lc.cfg, moe_finalize_kernel<Tt>, (const Tt*)y, (Tt*)out,
                                      (const int32_t*)expanded_to_permuted, (const float*)topk_w, (int)T, (int)K, hidden,
                                      (int)accumulate));
    return 0;
  });
}
