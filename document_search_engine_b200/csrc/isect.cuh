// k_score_isect — candidate-driven AND (included by bm25f.cu after stream.cuh).
//
// Whoosh evaluates And([...]) with IntersectionMatcher: the matcher of the rarer side is advanced
// and the other side is asked to skip_to() its docid; postings of the
// long lists that cannot match are never scored.  This kernel is the same idea for the GPU and is
// used for AND / AND-of-OR queries whose smallest group is much sparser than the rest (the host
// decides per query; symmetric ANDs stay on the streaming kernels):
//
//  * every WARP is an independent worker that pulls items (query, document range) from a global
//    counter; no shared-memory accumulators at all, so occupancy is bounded by registers only;
//  * the CANDIDATES are the postings of the leaves of the smallest group, 32 per step, one per lane
//    (docid, running score, mask of groups matched in registers);
//  * for every other leaf the warp first finds, cooperatively (32 probes per round), the end of the
//    window that can hold the step's docids - which is also that leaf's cursor for the next step -
//    and then every lane binary-searches its own docid inside the window;
//  * a candidate that collected every group is a match: counted, and offered to the warp's top-k
//    (one 64-bit key per lane, k <= 32) if its score reaches the current k-th best.
//
// Results are identical to the streaming kernels: same impacts, same FMA chain per document (own
// leaf first, then the other leaves in leaf order), same keys.
#pragma once

struct IsectParams {
  const uint2* pairs;
  const LeafRec* leaves;
  const QueryRec* queries;
  const ItemRec* items;            // tile_begin / tile_end hold the item's document range [lo, hi)
  unsigned long long* part_keys;   // [n_parts * k]
  unsigned long long* totals;      // [Q]
  unsigned int* queue;             // work counter, zeroed before the launch
  uint32_t n_items;
  uint32_t doc_base;
  int k;
};

constexpr int IS_WARPS = 8;

// Requires: k <= 32, <= 32 leaves, <= 32 groups, every leaf weight > 0, no after_key, no postings of
// deleted documents in the store.
__global__ void __launch_bounds__(IS_WARPS * 32) k_score_isect(IsectParams ip) {
  const int lane = threadIdx.x & 31;
  const uint2* __restrict__ store = ip.pairs;

  for (;;) {
    uint32_t item_idx = 0;
    if (lane == 0) item_idx = atomicAdd(ip.queue, 1u);
    item_idx = __shfl_sync(0xFFFFFFFFu, item_idx, 0);
    if (item_idx >= ip.n_items) break;

    const ItemRec item = ip.items[item_idx];
    const QueryRec q = ip.queries[item.q];
    const int L = (int)q.n_leaves;
    const uint32_t full = (q.n_groups >= 32u) ? 0xFFFFFFFFu : ((1u << q.n_groups) - 1u);
    const uint32_t d_lo = item.tile_begin, d_hi = item.tile_end;

    // lane l keeps leaf l: list origin, length, cursor (first posting not yet passed), weight, group
    unsigned long long s_off = 0ull;
    uint32_t s_df = 0u, s_start = 0u, s_grp = 0xFFFFFFFFu;
    float s_w = 0.0f;
    if (lane < L) {
      const LeafRec lf = ip.leaves[q.leaf_begin + lane];
      s_off = lf.off;
      s_df = lf.df;
      s_w = lf.w;
      s_grp = lf.group;
    }
    if (d_lo > 0u) {
      for (int l = 0; l < L; ++l) {
        const uint32_t st = warp_lower_bound(store + __shfl_sync(0xFFFFFFFFu, s_off, l), __shfl_sync(0xFFFFFFFFu, s_df, l), d_lo, lane);
        if (lane == l) s_start = st;
      }
    }
    const int n_cand = __popc(__ballot_sync(0xFFFFFFFFu, s_grp == 0u));    // leaves of the smallest group come first

    unsigned long long top = 0ull;            // lane i: i-th best key of this item so far
    unsigned long long thr_key = 0ull;
    float thr = 0.0f;
    unsigned int tot = 0;

    for (int c = 0; c < n_cand; ++c) {
      const uint2* __restrict__ cp = store + __shfl_sync(0xFFFFFFFFu, s_off, c);
      const uint32_t c_df = __shfl_sync(0xFFFFFFFFu, s_df, c);
      const float c_w = __shfl_sync(0xFFFFFFFFu, s_w, c);
      uint32_t s_cur = s_start;               // the other leaves' cursors restart with every candidate leaf
      for (uint32_t row = __shfl_sync(0xFFFFFFFFu, s_start, c); row < c_df; row += 32u) {
        const uint32_t i = row + (uint32_t)lane;
        uint2 r = make_uint2(0xFFFFFFFFu, 0u);
        if (i < c_df) r = ldg_pair(cp + i);
        const bool valid = r.x < d_hi;        // docids ascend: the valid lanes are a prefix
        const unsigned vm = __ballot_sync(0xFFFFFFFFu, valid);
        if (vm == 0u) break;
        const uint32_t doc = r.x;
        const uint32_t d_last = __shfl_sync(0xFFFFFFFFu, doc, 31 - __clz(vm));
        float score = c_w * __uint_as_float(r.y);
        uint32_t sat = 1u;                    // group 0
        bool dead = !valid;
        for (int l = 0; l < L; ++l) {
          if (l == c) continue;
          const uint32_t df = __shfl_sync(0xFFFFFFFFu, s_df, l);
          const uint32_t cur = __shfl_sync(0xFFFFFFFFu, s_cur, l);
          if (cur >= df) continue;            // list exhausted: nobody finds anything
          const uint2* __restrict__ lp = store + __shfl_sync(0xFFFFFFFFu, s_off, l);
          const uint32_t g = __shfl_sync(0xFFFFFFFFu, s_grp, l);
          const float wl = __shfl_sync(0xFFFFFFFFu, s_w, l);
          // end of the window: first posting past the step's last docid (the cursor of the next step)
          const uint32_t wend = cur + warp_lower_bound(lp + cur, df - cur, d_last + 1u, lane);
          uint32_t lo = cur, hi = wend;
          if (!valid) hi = lo;
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&lp[mid].x) < doc) lo = mid + 1u; else hi = mid;
          }
          if (valid && lo < wend) {
            const uint2 h = ldg_pair(lp + lo);
            if (h.x == doc) {
              score = fmaf(wl, __uint_as_float(h.y), score);
              sat |= 1u << g;
              if (g == 0u && l < c) dead = true;   // already a candidate of an earlier leaf of the group
            }
          }
          __syncwarp();
          if (lane == l) s_cur = wend;
        }
        const bool alive = !dead && sat == full;
        tot += alive ? 1u : 0u;
        unsigned long long key = 0ull;
        if (alive && score >= thr) key = make_key(score, ip.doc_base + doc);
        unsigned pm = __ballot_sync(0xFFFFFFFFu, key > thr_key);
        while (pm) {
          const int src = __ffs(pm) - 1;
          pm &= pm - 1u;
          const unsigned long long bk = __shfl_sync(0xFFFFFFFFu, key, src);
          if (bk > thr_key) {
            warp_topk_insert(top, bk, lane);
            thr_key = __shfl_sync(0xFFFFFFFFu, top, ip.k - 1);
          }
        }
        if (thr_key != 0ull) thr = key_score(thr_key);
      }
    }

    unsigned long long* out = ip.part_keys + (size_t)item.part * ip.k;
    if (lane < ip.k) out[lane] = top;
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xFFFFFFFFu, tot, o);
    if (lane == 0 && tot) atomicAdd(ip.totals + item.q, (unsigned long long)tot);
  }
}
