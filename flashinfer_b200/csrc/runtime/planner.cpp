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
namespace {
struct DecUnit {
  int32_t b, h, nt, q_start, q_len, kv_len, page_start, npages;
};
struct DecSeg {
  int32_t unit, t0, t1, cta;
};

// Two candidate schedules, scored by  max over CTAs of (tiles + kSegCost * segments):
//  * stream-K: walk (request, kv_head, tile) in order, cut into equal quotas -> perfect tile balance, but units
//    straddle CTAs (extra pipeline start-ups + split-KV merges); best when every CTA has many tiles.
//  * LPT: units stay whole unless longer than the piece size; pieces go longest-first to the least-loaded CTA ->
//    few segments and merges; best for small batches / short sequences where start-up latency dominates.
constexpr double kSegCost = 2.5;  // pipeline start-up of a segment measured in KV tiles (~3.5 us vs ~1.4 us / tile)

double score(const std::vector<DecSeg>& segs, int64_t num_ctas) {
  std::vector<double> load(num_ctas, 0.0);
  for (const DecSeg& s : segs) load[s.cta] += double(s.t1 - s.t0) + kSegCost;
  double mx = 0;
  for (double l : load) mx = std::max(mx, l);
  return mx;
}

std::vector<DecSeg> plan_streamk(const std::vector<DecUnit>& units, int64_t quota) {
  std::vector<DecSeg> segs;
  int64_t cta = 0, room = quota;
  for (size_t u = 0; u < units.size(); ++u) {
    int64_t t = 0;
    const int64_t nt = units[u].nt;
    while (t < nt) {
      if (room == 0) {
        ++cta;
        room = quota;
      }
      const int64_t take = std::min(room, nt - t);
      segs.push_back({(int32_t)u, (int32_t)t, (int32_t)(t + take), (int32_t)cta});
      t += take;
      room -= take;
    }
  }
  return segs;
}

std::vector<DecSeg> plan_lpt(const std::vector<DecUnit>& units, int64_t piece, int64_t num_ctas) {
  std::vector<DecSeg> segs;
  for (size_t u = 0; u < units.size(); ++u) {
    const int64_t nt = units[u].nt;
    const int64_t parts = (nt + piece - 1) / piece;
    for (int64_t i = 0; i < parts; ++i)
      segs.push_back({(int32_t)u, (int32_t)(nt * i / parts), (int32_t)(nt * (i + 1) / parts), 0});
  }
  std::vector<int64_t> order(segs.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int64_t)i;
  std::stable_sort(order.begin(), order.end(),
                   [&](int64_t x, int64_t y) { return segs[x].t1 - segs[x].t0 > segs[y].t1 - segs[y].t0; });
  using Load = std::pair<double, int64_t>;
  std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
  for (int64_t c = 0; c < num_ctas; ++c) heap.push({0.0, c});
  for (int64_t idx : order) {
    Load l = heap.top();
    heap.pop();
    segs[idx].cta = (int32_t)l.second;
    heap.push({l.first + double(segs[idx].t1 - segs[idx].t0) + kSegCost, l.second});
  }
  return segs;
}
}  // namespace

extern "C" int decode_plan(const int32_t* kv_page_indptr, const int32_t* kv_lens, const int32_t* q_indptr, int64_t batch,
                           int64_t num_kv_heads, int64_t group, int64_t page_size, int64_t num_ctas,
                           int64_t min_tiles_per_cta, int32_t* seg_info, int64_t max_segs, int32_t* cta_seg_indptr,
                           int32_t* merge_items, int64_t max_merge, int64_t* counts) {
  if (batch < 0 || num_kv_heads <= 0 || page_size <= 0 || num_ctas <= 0) return fail("decode_plan: bad arguments");
  std::vector<DecUnit> units;
  int64_t total = 0;
  int64_t max_q_rows = 1;
  for (int64_t b = 0; b < batch; ++b) {
    if (kv_lens[b] < 0) return fail("decode_plan: negative kv_len");
    const int64_t nt = num_kv_tiles(kv_lens[b], page_size);
    const int64_t ql = q_indptr ? (q_indptr[b + 1] - q_indptr[b]) : 1;
    max_q_rows = std::max(max_q_rows, ql * group);
    if (nt == 0 || ql <= 0) continue;
    total += nt * num_kv_heads;
    for (int64_t h = 0; h < num_kv_heads; ++h)
      units.push_back({(int32_t)b, (int32_t)h, (int32_t)nt, q_indptr ? q_indptr[b] : (int32_t)b, (int32_t)ql, kv_lens[b],
                       kv_page_indptr[b], kv_page_indptr[b + 1] - kv_page_indptr[b]});
  }
  const bool no_split = min_tiles_per_cta >= (int64_t(1) << 29);
  int64_t quota = (total + num_ctas - 1) / num_ctas;
  quota = std::max<int64_t>(quota, no_split ? 1 : std::max<int64_t>(1, min_tiles_per_cta));
  std::vector<DecSeg> segs;
  if (no_split) {
    segs = plan_lpt(units, int64_t(1) << 30, num_ctas);  // whole units only
  } else {
    segs = plan_streamk(units, quota);
    const std::vector<DecSeg> alt = plan_lpt(units, std::max<int64_t>(quota, 8), num_ctas);
    if (score(alt, num_ctas) < score(segs, num_ctas)) segs = alt;
  }
  if ((int64_t)segs.size() > max_segs) return fail("decode_plan: seg_info capacity exceeded");
  // merge bookkeeping per unit
  std::vector<int32_t> parts(units.size(), 0), slot0(units.size(), -1), seen(units.size(), 0);
  for (const DecSeg& sg : segs) ++parts[sg.unit];
  int64_t nmerge = 0, nslots = 0;
  for (size_t u = 0; u < units.size(); ++u) {
    if (parts[u] > 1) {
      if (nmerge >= max_merge) return fail("decode_plan: merge_items capacity exceeded");
      slot0[u] = (int32_t)nslots;
      int32_t* m = merge_items + nmerge * kMergeInts;
      m[0] = (int32_t)nslots;
      m[1] = parts[u];
      m[2] = units[u].q_start;
      m[3] = units[u].q_len;
      m[4] = units[u].h;
      m[5] = units[u].b;
      m[6] = 0;
      m[7] = 0;
      nslots += parts[u];
      ++nmerge;
    }
  }
  // emit grouped by CTA (stable: keeps kv order inside a unit)
  std::vector<int64_t> order(segs.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int64_t)i;
  std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return segs[x].cta < segs[y].cta; });
  int64_t nseg = 0;
  int64_t cur = 0;
  cta_seg_indptr[0] = 0;
  for (int64_t idx : order) {
    const DecSeg& sg = segs[idx];
    if (sg.cta >= num_ctas) return fail("decode_plan: internal error (cta overflow)");
    while (cur < sg.cta) cta_seg_indptr[++cur] = (int32_t)nseg;
    const DecUnit& u = units[sg.unit];
    int32_t* s = seg_info + nseg * kSegInts;
    s[0] = u.b;
    s[1] = u.h;
    s[2] = sg.t0;
    s[3] = sg.t1;
    s[4] = parts[sg.unit] > 1 ? slot0[sg.unit] + seen[sg.unit] : -1;
    s[5] = u.q_start;
    s[6] = u.q_len;
    s[7] = u.kv_len;
    s[8] = u.page_start;
    s[9] = u.npages;
    s[10] = parts[sg.unit] > 1 ? slot0[sg.unit] : 0;
    s[11] = parts[sg.unit] > 1 ? parts[sg.unit] : 0;
    ++seen[sg.unit];
    ++nseg;
  }
  for (int64_t c = cur + 1; c <= num_ctas; ++c) cta_seg_indptr[c] = (int32_t)nseg;
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

// ---------------------------------------------------------------------------------------------
// MLA decode plan (reference include/flashinfer/attention/scheduler.cuh MLAPlan :1440-1710).
//   One work item = (query token, KV chunk); a chunk is a multiple of the 32-token MLA tile.  The chunk length is the smallest
//   one for which all items fit ONE wave of CTA pairs (every query token is at least one item), found by bisection.
//   qo_indptr / kv_page_indptr [batch + 1], kv_lens [batch] (int64).  causal: token i of a q_len-token request sees
//   kv_len - (q_len - 1 - i) keys (MTP / speculative decode).
// Outputs: work [num_work][8] = {q_row, page_start, kv_lo, kv_hi, kv_visible, partial_slot, num_pages, (kmax << 16) | nsplits},
//          row_parts [n_q] = splits of every query row, counts = {num_work, kmax, n_q, chunk}.
// ---------------------------------------------------------------------------------------------
extern "C" int mla_plan(const int64_t* qo_indptr, const int64_t* kv_page_indptr, const int64_t* kv_lens, int64_t batch, int64_t causal,
                        int64_t num_ctas, int64_t tile, int32_t* work, int64_t max_work, int32_t* row_parts, int64_t max_rows,
                        int64_t* counts) {
  if (batch < 0 || num_ctas <= 0 || tile <= 0) return fail("mla_plan: bad arguments");
  struct Row {
    int64_t q_row, page_start, vis, npages;
  };
  std::vector<Row> rows;
  int64_t total_tokens = 0, max_vis = tile;
  for (int64_t b = 0; b < batch; ++b) {
    const int64_t ql = qo_indptr[b + 1] - qo_indptr[b];
    if (kv_lens[b] < 0 || ql < 0) return fail("mla_plan: negative length");
    for (int64_t i = 0; i < ql; ++i) {
      int64_t vis = causal ? kv_lens[b] - (ql - 1 - i) : kv_lens[b];
      if (vis < 0) vis = 0;
      rows.push_back({qo_indptr[b] + i, kv_page_indptr[b], vis, kv_page_indptr[b + 1] - kv_page_indptr[b]});
      total_tokens += vis;
      max_vis = std::max(max_vis, vis);
    }
  }
  const int64_t n_q = batch > 0 ? qo_indptr[batch] : 0;
  if (n_q > max_rows) return fail("mla_plan: row_parts capacity exceeded");
  auto ceil_div = [](int64_t a, int64_t b) { return (a + b - 1) / b; };
  auto count = [&](int64_t c) {
    int64_t n = 0;
    for (const Row& r : rows) n += std::max<int64_t>(1, ceil_div(r.vis, c));
    return n;
  };
  int64_t chunk = std::max<int64_t>(4 * tile, ceil_div(total_tokens, num_ctas));
  chunk = ceil_div(chunk, tile) * tile;
  const int64_t target = std::max<int64_t>(num_ctas, (int64_t)rows.size());
  if (count(chunk) > target) {
    int64_t lo = chunk / tile, hi = std::max<int64_t>(chunk / tile, ceil_div(max_vis, tile));
    while (lo < hi) {
      const int64_t mid = (lo + hi) / 2;
      if (count(mid * tile) > target) lo = mid + 1;
      else hi = mid;
    }
    chunk = lo * tile;
  }
  int64_t kmax = 1;
  for (const Row& r : rows) kmax = std::max<int64_t>(kmax, ceil_div(r.vis, chunk));
  if (kmax >= (int64_t(1) << 15)) return fail("mla_plan: too many KV splits");
  for (int64_t i = 0; i < n_q; ++i) row_parts[i] = 1;
  int64_t nw = 0;
  for (const Row& r : rows) {
    const int64_t nsp = std::max<int64_t>(1, ceil_div(r.vis, chunk));
    row_parts[r.q_row] = (int32_t)nsp;
    for (int64_t s = 0; s < nsp; ++s) {
      if (nw >= max_work) return fail("mla_plan: work capacity exceeded");
      const int64_t lo = s * chunk, hi = std::min<int64_t>(r.vis, (s + 1) * chunk);
      int32_t* w = work + nw * 8;
      w[0] = (int32_t)r.q_row;
      w[1] = (int32_t)r.page_start;
      w[2] = (int32_t)lo;
      w[3] = (int32_t)std::max(hi, lo);
      w[4] = (int32_t)r.vis;
      w[5] = (int32_t)(r.q_row * kmax + s);
      w[6] = (int32_t)std::max<int64_t>(r.npages, 1);
      w[7] = (int32_t)((kmax << 16) | nsp);
      ++nw;
    }
  }
  counts[0] = nw;
  counts[1] = kmax;
  counts[2] = n_q;
  counts[3] = chunk;
  return 0;
}
