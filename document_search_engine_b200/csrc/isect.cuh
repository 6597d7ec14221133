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
//    (KR 64-bit keys per lane, k <= 32 * KR) if its score reaches the current k-th best.
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
  const uint32_t* n_items_dev;     // batches planned on the device (k_plan_*): the item count and the position of this
  const uint32_t* items_off_dev;   // kernel's items inside `items` live there (else null)
  uint32_t doc_base;
  int k;
  // FINAL instantiation only (a weighting with a final() step, reference my_whoosh.py:127-154):
  const double* final_add;         // [n_docs] date term of the document (NaN: the document has no date)
  unsigned int* part_lo;           // [n_parts * k] low halves of the keys (~global docnum); part_keys holds the high halves
};

constexpr int IS_WARPS = 8;

template <int KR, bool FINAL>
__global__ void __launch_bounds__(IS_WARPS * 32, (FINAL || KR > 4) ? 2 : KR == 1 ? 5 : 3) k_score_isect(IsectParams ip) {   // 48 registers for k <= 32: two CTAs fit beside the stream kernel
  const int lane = threadIdx.x & 31;
  const uint2* __restrict__ store = ip.pairs;
  if (ip.n_items_dev && blockIdx.x * (uint32_t)IS_WARPS >= __ldg(ip.n_items_dev)) return;   // device-planned batch, full grid

  for (;;) {
    uint32_t item_idx = 0;
    if (lane == 0) item_idx = atomicAdd(ip.queue, 1u);
    item_idx = __shfl_sync(0xFFFFFFFFu, item_idx, 0);
    if (item_idx >= (ip.n_items_dev ? __ldg(ip.n_items_dev) : ip.n_items)) break;

    const ItemRec item = ip.items[(ip.items_off_dev ? __ldg(ip.items_off_dev) : 0u) + item_idx];
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
    // candidate leaves: the leaves of the smallest group (rank 0; leaves of NOT clauses precede them)
    const unsigned cand_mask = __ballot_sync(0xFFFFFFFFu, s_grp == 0u);

    unsigned long long top[KR];               // lane i, row j: the (32 j + i)-th best key of this item so far
#pragma unroll
    for (int j = 0; j < KR; ++j) top[j] = 0ull;
    unsigned long long thr_key = 0ull;
    float thr = 0.0f;
    unsigned int tot = 0;
    uint32_t topl[FINAL ? KR : 1];            // FINAL: the low halves of the keys; thr_key / thr_lo = the k-th best
    uint32_t thr_lo = 0u;
#pragma unroll
    for (int j = 0; j < (FINAL ? KR : 1); ++j) topl[j] = 0u;

    for (unsigned cm = cand_mask; cm; cm &= cm - 1u) {
      const int c = __ffs(cm) - 1;
      const uint2* __restrict__ cp = store + __shfl_sync(0xFFFFFFFFu, s_off, c);
      const uint32_t c_df = __shfl_sync(0xFFFFFFFFu, s_df, c);
      const float c_w = __shfl_sync(0xFFFFFFFFu, s_w, c);
      uint32_t s_cur = s_start;               // the other leaves' cursors restart with every candidate leaf
      uint32_t s_win = 64u;                   // size of the leaf's previous window
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
          // end of the window: first posting past the step's last docid (the cursor of the next step).
          // Windows of consecutive steps are about the same size, so look in twice the previous window
          // first (nearby lines, two probe rounds) and in the rest of the list only if that fails.
          const uint32_t near = min(df - cur, 2u * __shfl_sync(0xFFFFFFFFu, s_win, l) + 32u);
          uint32_t wlen = warp_lower_bound(lp + cur, near, d_last + 1u, lane);
          if (wlen == near && near < df - cur) wlen = near + warp_lower_bound(lp + cur + near, df - cur - near, d_last + 1u, lane);
          const uint32_t wend = cur + wlen;
          // Each lane now looks for its own docid in [cur, wend).  One strided load puts the last posting of 32
          // equal buckets of the window into the lanes, five shuffle rounds find every lane's bucket (the first
          // whose last docid is not below the lane's), and only the inside of that bucket is searched through
          // memory: a window of up to 32 postings costs one dependent load, one of 1024 five instead of ten.
          // (Measured: config 3, all AND-of-OR, 4.76 -> 4.36 ms per 10k queries; config 2 unchanged within 1 %.)
          const uint32_t stride = (wlen + 31u) >> 5;
          uint32_t sx = 0xFFFFFFFFu;
          {
            const uint32_t b0 = (uint32_t)lane * stride;
            if (b0 < wlen) sx = __ldg(&lp[cur + min(b0 + stride, wlen) - 1u].x);
          }
          uint32_t bk = 0u;
#pragma unroll
          for (uint32_t st = 16u; st; st >>= 1) {
            const uint32_t v = __shfl_sync(0xFFFFFFFFu, sx, bk + st - 1u);
            if (v < doc) bk += st;
          }
          uint32_t lo = min(cur + bk * stride, wend), hi = min(lo + stride, wend);
          if (!valid) hi = lo;
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&lp[mid].x) < doc) lo = mid + 1u; else hi = mid;
          }
          if (valid && lo < wend) {
            const uint2 h = ldg_pair(lp + lo);
            if (h.x == doc) {
              score = fmaf(wl, __uint_as_float(h.y), score);
              sat |= 1u << g;                      // a leaf of a NOT clause sets bit NEG_GROUP, which `full` never has
              if (g == 0u && l < c) dead = true;   // already a candidate of an earlier leaf of the group
            }
          }
          __syncwarp();
          if (lane == l) { s_cur = wend; s_win = wlen; }
        }
        const bool alive = !dead && sat == full;
        tot += alive ? 1u : 0u;
        if constexpr (FINAL) {
          // final() of every match, then the same insertion with 96-bit keys
          unsigned long long kh = 0ull;
          uint32_t kl = 0u;
          if (alive) {
            kh = orderable_f64(final_value(score, __ldg(ip.final_add + doc)));
            kl = 0xFFFFFFFFu - (ip.doc_base + doc);
          }
          unsigned pm = __ballot_sync(0xFFFFFFFFu, key2_wanted(kh, kl, thr_key, thr_lo, q.after_key, q.after_lo));
          while (pm) {
            const int src = __ffs(pm) - 1;
            pm &= pm - 1u;
            const unsigned long long bh = __shfl_sync(0xFFFFFFFFu, kh, src);
            const uint32_t bl = __shfl_sync(0xFFFFFFFFu, kl, src);
            if (key2_gt(bh, bl, thr_key, thr_lo)) {
              warp_topk2_insert_rows<KR>(top, topl, bh, bl, lane);
              warp_topk2_kth<KR>(top, topl, ip.k, thr_key, thr_lo);
            }
          }
        } else {
        unsigned long long key = 0ull;
        if (alive && score >= thr) key = make_key(score, ip.doc_base + doc);
        unsigned pm = __ballot_sync(0xFFFFFFFFu, key > thr_key);
        while (pm) {
          const int src = __ffs(pm) - 1;
          pm &= pm - 1u;
          const unsigned long long bk = __shfl_sync(0xFFFFFFFFu, key, src);
          if (bk > thr_key) {
            warp_topk_insert_rows<KR>(top, bk, lane);
            thr_key = warp_topk_kth<KR>(top, ip.k);
          }
        }
        if (thr_key != 0ull) thr = key_score(thr_key);
        }
      }
    }

    unsigned long long* out = ip.part_keys + (size_t)item.part * ip.k;
#pragma unroll
    for (int j = 0; j < KR; ++j)
      if (32 * j + lane < ip.k) out[32 * j + lane] = top[j];
    if constexpr (FINAL) {
      unsigned int* out_lo = ip.part_lo + (size_t)item.part * ip.k;
#pragma unroll
      for (int j = 0; j < KR; ++j)
        if (32 * j + lane < ip.k) out_lo[32 * j + lane] = topl[j];
    }
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xFFFFFFFFu, tot, o);
    if (lane == 0 && tot) atomicAdd(ip.totals + item.q, (unsigned long long)tot);
  }
}
