// Host-side work planners (C++).  Pure CPU code: callable (and unit-tested) without a GPU.
//
// Parity target: the reference's plan() layer — DecodePlan / PrefillPlan / MLAPlan in
// include/flashinfer/attention/scheduler.cuh:426-492,694-797,1440-1710 — but re-designed for
// persistent sm_100a kernels: instead of choosing a grid and a kv-chunk size by occupancy and
// binary search, every planner flattens the work into uniform "tiles" and cuts the flattened
// space into equal per-CTA quotas (stream-K over the KV / Q-tile space).  Output is a set of
// int32 arrays written into one pinned host buffer that python uploads with a single H2D copy
// (same contract as the reference: plan() is host work + one cudaMemcpyAsync, run() is launch-only
// and CUDA-graph capturable because every array has a fixed upper-bound size).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <queue>
#include <string>
#include <vector>

namespace {
thread_local std::string g_err;
int fail(const std::string& m) {
  g_err = m;
  return 1;
}

constexpr int kTileKV = 128;
constexpr int kSegInts = 12;
constexpr int kMergeInts = 8;

inline int64_t num_kv_tiles(int64_t kv_len, int64_t page_size) {
  if (kv_len <= 0) return 0;
  if (page_size <= kTileKV) {
    const int64_t tpt = (kTileKV / page_size) * page_size;
    return (kv_len + tpt - 1) / tpt;
  }
  const int64_t tpp = (page_size + kTileKV - 1) / kTileKV;
  const int64_t full_pages = kv_len / page_size;
  const int64_t rem = kv_len % page_size;
  return full_pages * tpp + (rem + kTileKV - 1) / kTileKV;
}
}  // namespace

extern "C" const char* fib200_last_error() { return g_err.c_str(); }

// ---------------------------------------------------------------------------------------------
// Decode / small-q append plan.
//   kv_page_indptr [B+1] : page-list offsets          kv_lens [B] : tokens per request
//   q_indptr [B+1] or nullptr (=> q_len 1, q_start b)
// Outputs:
//   seg_info [max_segs][12] : {req, kv_head, tile_begin, tile_end, slot(-1 = final), q_start, q_len,
//                              kv_len, page_start, num_pages, first_slot, num_parts}
//   cta_seg_indptr [num_ctas+1]
//   merge_items [max_merge][8] : {slot0, nparts, q_start, q_len, kv_head, req, 0, 0}
//   counts[0]=nseg, [1]=nmerge, [2]=nslots, [3]=max_q_rows(q_len*group), [4]=total_tiles, [5]=quota
// ---------------------------------------------------------------------------------------------
extern "C" int decode_plan(const int32_t* kv_page_indptr, const int32_t* kv_lens, const int32_t* q_indptr, int64_t batch,
                           int64_t num_kv_heads, int64_t group, int64_t page_size, int64_t num_ctas,
                           int64_t min_tiles_per_cta, int32_t* seg_info, int64_t max_segs, int32_t* cta_seg_indptr,
                           int32_t* merge_items, int64_t max_merge, int64_t* counts) {
  if (batch < 0 || num_kv_heads <= 0 || page_size <= 0 || num_ctas <= 0) return fail("decode_plan: bad arguments");
  std::vector<int64_t> tiles(batch);
  int64_t total = 0;
  int64_t max_q_rows = 1;
  for (int64_t b = 0; b < batch; ++b) {
    if (kv_lens[b] < 0) return fail("decode_plan: negative kv_len");
    tiles[b] = num_kv_tiles(kv_lens[b], page_size);
    total += tiles[b] * num_kv_heads;
    const int64_t ql = q_indptr ? (q_indptr[b + 1] - q_indptr[b]) : 1;
    max_q_rows = std::max(max_q_rows, ql * group);
  }
  int64_t quota = (total + num_ctas - 1) / num_ctas;
  quota = std::max<int64_t>(quota, std::max<int64_t>(1, min_tiles_per_cta));
  int64_t nseg = 0, nmerge = 0, nslots = 0;
  int64_t cta = 0, room = quota;  // tiles still available in the current CTA
  cta_seg_indptr[0] = 0;
  for (int64_t b = 0; b < batch; ++b) {
    const int64_t nt = tiles[b];
    if (nt == 0) continue;
    const int32_t q_start = q_indptr ? q_indptr[b] : (int32_t)b;
    const int32_t q_len = q_indptr ? (q_indptr[b + 1] - q_indptr[b]) : 1;
    if (q_len <= 0) continue;
    const int32_t page_start = kv_page_indptr[b];
    const int32_t npages = kv_page_indptr[b + 1] - kv_page_indptr[b];
    for (int64_t h = 0; h < num_kv_heads; ++h) {
      int64_t t = 0;
      const int64_t first_seg = nseg;
      while (t < nt) {
        if (room == 0) {
          ++cta;
          if (cta >= num_ctas) return fail("decode_plan: internal error (cta overflow)");
          cta_seg_indptr[cta] = (int32_t)nseg;
          room = quota;
        }
        const int64_t take = std::min(room, nt - t);
        if (nseg >= max_segs) return fail("decode_plan: seg_info capacity exceeded");
        int32_t* s = seg_info + nseg * kSegInts;
        s[0] = (int32_t)b;
        s[1] = (int32_t)h;
        s[2] = (int32_t)t;
        s[3] = (int32_t)(t + take);
        s[4] = -1;
        s[5] = q_start;
        s[6] = q_len;
        s[7] = kv_lens[b];
        s[8] = page_start;
        s[9] = npages;
        s[10] = 0;
        s[11] = 0;
        ++nseg;
        t += take;
        room -= take;
      }
      const int64_t parts = nseg - first_seg;
      if (parts > 1) {
        if (nmerge >= max_merge) return fail("decode_plan: merge_items capacity exceeded");
        for (int64_t i = 0; i < parts; ++i) {
          seg_info[(first_seg + i) * kSegInts + 4] = (int32_t)(nslots + i);
          seg_info[(first_seg + i) * kSegInts + 10] = (int32_t)nslots;  // first slot of the item (= counter id)
          seg_info[(first_seg + i) * kSegInts + 11] = (int32_t)parts;   // number of partial states to merge
        }
        int32_t* m = merge_items + nmerge * kMergeInts;
        m[0] = (int32_t)nslots;
        m[1] = (int32_t)parts;
        m[2] = q_start;
        m[3] = q_len;
        m[4] = (int32_t)h;
        m[5] = (int32_t)b;
        m[6] = 0;
        m[7] = 0;
        nslots += parts;
        ++nmerge;
      }
    }
  }
  for (int64_t c = cta + 1; c <= num_ctas; ++c) cta_seg_indptr[c] = (int32_t)nseg;
  counts[0] = nseg;
  counts[1] = nmerge;
  counts[2] = nslots;
  counts[3] = max_q_rows;
  counts[4] = total;
  counts[5] = quota;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Prefill / append plan (ragged or paged KV): work unit = one (request, q_tile, kv_head-group) with
// cost ~ number of kv tiles it has to visit (causal aware).  Units are distributed over CTAs with
// the LPT rule (longest processing time first onto the least-loaded CTA, min-heap), the same
// balancing idea as the reference's PrefillSM90Plan (scheduler.cuh:870-1019), and each CTA's list
// is emitted contiguously.
//   work_info [max_work][8] : {req, q_tile_start(row), q_rows, kv_head? (-1 = loop heads), kv_len, q_len,
//                              qo_start, kv_start}
// ---------------------------------------------------------------------------------------------
extern "C" int prefill_plan(const int32_t* qo_indptr, const int32_t* kv_lens, const int32_t* kv_start, int64_t batch,
                            int64_t num_qo_heads, int64_t tile_q, int64_t tile_kv, int64_t causal, int64_t window_left,
                            int64_t num_ctas, int32_t* work_info, int64_t max_work, int32_t* cta_work_indptr,
                            int64_t* counts) {
  struct Unit {
    int32_t req, q0, rows, head, kv_len, q_len, qo_start, kv_start;
    int64_t cost;
  };
  std::vector<Unit> units;
  for (int64_t b = 0; b < batch; ++b) {
    const int32_t q_len = qo_indptr[b + 1] - qo_indptr[b];
    const int32_t kv_len = kv_lens[b];
    if (q_len <= 0) continue;
    for (int32_t q0 = 0; q0 < q_len; q0 += (int32_t)tile_q) {
      const int32_t rows = std::min<int32_t>((int32_t)tile_q, q_len - q0);
      int64_t kv_hi = kv_len;
      if (causal) kv_hi = std::min<int64_t>(kv_len, (int64_t)kv_len - q_len + q0 + rows);
      int64_t kv_lo = 0;
      if (window_left >= 0) kv_lo = std::max<int64_t>(0, (int64_t)kv_len - q_len + q0 - window_left);
      if (kv_hi < 0) kv_hi = 0;
      const int64_t ntiles = kv_hi > kv_lo ? (kv_hi - (kv_lo / tile_kv) * tile_kv + tile_kv - 1) / tile_kv : 0;
      for (int64_t h = 0; h < num_qo_heads; ++h) {
        Unit u{(int32_t)b, q0, rows, (int32_t)h, kv_len, q_len, qo_indptr[b], kv_start ? kv_start[b] : 0,
               ntiles + 1};
        units.push_back(u);
      }
    }
  }
  if ((int64_t)units.size() > max_work) return fail("prefill_plan: work_info capacity exceeded");
  // LPT: sort by cost desc, assign to least-loaded CTA
  std::vector<int64_t> order(units.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int64_t)i;
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return units[a].cost > units[b].cost; });
  using Load = std::pair<int64_t, int64_t>;  // (load, cta)
  std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
  for (int64_t c = 0; c < num_ctas; ++c) heap.push({0, c});
  std::vector<std::vector<int64_t>> per_cta(num_ctas);
  for (int64_t idx : order) {
    Load l = heap.top();
    heap.pop();
    per_cta[l.second].push_back(idx);
    heap.push({l.first + units[idx].cost, l.second});
  }
  int64_t n = 0;
  int64_t max_load = 0;
  for (int64_t c = 0; c < num_ctas; ++c) {
    cta_work_indptr[c] = (int32_t)n;
    int64_t load = 0;
    for (int64_t idx : per_cta[c]) {
      const Unit& u = units[idx];
      int32_t* w = work_info + n * 8;
      w[0] = u.req;
      w[1] = u.q0;
      w[2] = u.rows;
      w[3] = u.head;
      w[4] = u.kv_len;
      w[5] = u.q_len;
      w[6] = u.qo_start;
      w[7] = u.kv_start;
      load += u.cost;
      ++n;
    }
    max_load = std::max(max_load, load);
  }
  cta_work_indptr[num_ctas] = (int32_t)n;
  counts[0] = n;
  counts[1] = max_load;
  return 0;
}
