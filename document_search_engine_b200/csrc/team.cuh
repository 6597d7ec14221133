// k_score_team — the hot kernel of libbm25f (included by bm25f.cu after stream.cuh).
//
// Replaces Whoosh's matcher loop + BM25FScorer + TopCollector for one batch of lowered queries
// (reference call sites my_flask.py:208, :211, :304).
//
// Same arithmetic and slot formats as k_score_stream; what changes is who finds the postings:
//
//  * PERSISTENT grid, one CTA per SM.  A CTA (a TEAM of warps) pulls one work item (query,
//    document range) at a time, heaviest first, from a global counter.
//  * The item's document range is cut into SLICES of as many documents as one warp's accumulator
//    slots hold.  All threads of the team first build a BOUNDS table
//        bounds[slice boundary][leaf] = first posting with docid >= boundary
//    (independent binary searches, four interleaved per thread so their latencies overlap).
//  * Then the warps take the slices round-robin and, for each leaf that has postings in the slice,
//    read exactly that range: at most 32 postings are one load per lane (issued before the previous
//    leaf is processed); longer ranges stream super-rows of 128 postings (four 64-bit loads in
//    flight per lane, the next super-row requested before the current one is processed), interior
//    super-rows with no masks at all.  No cursors, no tails, no barriers between leaves: a warp's
//    slots are private.  A slice of an AND in which some group has no posting is skipped without
//    reading anything.
//  * Every warp keeps its own hot list and top-k (one key per lane, k <= 32); the admission
//    threshold is shared through shared memory (atomicMax of the k-th best score), and the lists
//    are merged once per item.  The team synchronises three times per table window, not per leaf.
#pragma once

constexpr int TM_MAX_LEAVES = 8;
constexpr int TM_BOUNDS_WORDS = 4096;   // table window: (slices + 1) * leaves <= this
constexpr int TM_MAX_WARPS = 16;
constexpr int TM_HOT = 64;              // hot-list entries per warp

struct TeamParams {
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
  uint32_t slot_bytes;             // accumulator bytes per warp (multiple of 512)
  uint32_t doc_base;
  uint32_t prefetch;               // slices ahead to bulk-prefetch into L2 (0: off)
  int k;
};

struct TeamShared {
  uint32_t bounds[TM_BOUNDS_WORDS];
  unsigned long long keys[TM_MAX_WARPS][32];
  uint16_t hot[TM_MAX_WARPS][TM_HOT];
  uint32_t nhot[TM_MAX_WARPS];
  uint32_t thr_bits;                        // max over warps of the k-th best score (positive float bits order as integers)
  uint32_t item;
  unsigned long long total;
};

// Requires: k <= 32, <= TM_MAX_LEAVES leaves, every leaf weight > 0, no after_key, no postings of
// deleted documents in the store (bm25f_create compacts them away).
__global__ void __launch_bounds__(TM_MAX_WARPS * 32, 1) k_score_team(TeamParams tp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int NW = nthr >> 5;
  const uint32_t slot_bytes = tp.slot_bytes;
  if (tp.n_items_dev && blockIdx.x >= __ldg(tp.n_items_dev)) return;      // device-planned batch, full grid: nothing left for this CTA
  TeamShared& sh = *reinterpret_cast<TeamShared*>(smem_raw + (size_t)NW * slot_bytes);
  SubCtx cx;
  cx.slots_addr = smem_u32(smem_raw) + (uint32_t)warp * slot_bytes;
  cx.hot_addr = smem_u32(sh.hot[warp]);
  cx.cnt_addr = smem_u32(&sh.nhot[warp]);

  for (uint32_t o = (uint32_t)lane * 16u; o < slot_bytes; o += 512u) sts_zero16(cx.slots_addr + o);
  if (lane == 0) sh.nhot[warp] = 0u;

  for (;;) {
    __syncthreads();                       // the previous item is finished by everybody
    if (tid == 0) sh.item = atomicAdd(tp.queue, 1u);
    __syncthreads();
    const uint32_t item_idx = sh.item;
    if (item_idx >= (tp.n_items_dev ? __ldg(tp.n_items_dev) : tp.n_items)) break;

    const ItemRec item = tp.items[(tp.items_off_dev ? __ldg(tp.items_off_dev) : 0u) + item_idx];
    const QueryRec q = tp.queries[item.q];
    const uint32_t L = q.n_leaves;
    const uint32_t G = q.n_groups;
    const bool simple_or = (q.flags & QF_SIMPLE_OR) != 0;
    const uint32_t shift = simple_or ? 2u : 3u;
    const uint32_t SW = slot_bytes >> shift;            // documents per slice
    const uint32_t d_lo = item.tile_begin, d_hi = item.tile_end;
    const uint32_t n_slices = (d_hi - d_lo + SW - 1u) / SW;
    const uint32_t win = (uint32_t)TM_BOUNDS_WORDS / L - 1u;   // slices per table window

    if (tid == 0) { sh.thr_bits = 0x00800000u; sh.total = 0ull; }   // FLT_MIN: every first hit is hot

    // lane l keeps leaf l's constants; lane g keeps the leaf mask of group g
    unsigned long long lf_base = 0ull;
    float lf_w = 0.0f;
    uint32_t lf_grp = 0xFFFFFFFFu;
    if ((uint32_t)lane < L) {
      const LeafRec lf = tp.leaves[q.leaf_begin + lane];
      lf_base = lf.off - (lf.off & 31ull);
      lf_w = lf.w;
      lf_grp = lf.group;
    }
    uint32_t gm = 0u;
    for (uint32_t l = 0; l < L; ++l)
      if (__shfl_sync(0xFFFFFFFFu, lf_grp, l) == (uint32_t)lane) gm |= 1u << l;

    unsigned long long top = 0ull;            // lane i: i-th best key this warp has seen in this item
    unsigned long long thr_key = 0ull;
    unsigned int tot = 0;
    const uint32_t thr_addr = smem_u32(&sh.thr_bits);

    for (uint32_t win_lo = 0; win_lo < n_slices; win_lo += win) {
      const uint32_t win_n = min(win, n_slices - win_lo);
      __syncthreads();                     // the previous window's table is no longer in use
      // ---- bounds table: entry i = boundary (i / L) of leaf (i % L); four searches per thread per
      // round, interleaved so that their load latencies overlap
      const uint32_t n_ent = (win_n + 1u) * L;
      for (uint32_t i0 = (uint32_t)tid; i0 < n_ent; i0 += 4u * (uint32_t)nthr) {
        const uint2* p[4];
        uint32_t lo[4], hi[4], tg[4], al[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t i = i0 + (uint32_t)j * (uint32_t)nthr;
          lo[j] = hi[j] = tg[j] = al[j] = 0u;
          p[j] = tp.pairs;
          if (i < n_ent) {
            const uint32_t b = i / L, l = i - b * L;
            const LeafRec lf = tp.leaves[q.leaf_begin + l];
            p[j] = tp.pairs + lf.off;
            al[j] = (uint32_t)(lf.off & 31ull);
            const unsigned long long t64 = (unsigned long long)d_lo + (unsigned long long)(win_lo + b) * SW;
            tg[j] = (uint32_t)min(t64, (unsigned long long)d_hi);
            hi[j] = (tg[j] == 0u) ? 0u : lf.df;
          }
        }
        while ((lo[0] < hi[0]) | (lo[1] < hi[1]) | (lo[2] < hi[2]) | (lo[3] < hi[3])) {
          uint32_t mid[4], dv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            mid[j] = (lo[j] + hi[j]) >> 1;
            dv[j] = 0u;
            if (lo[j] < hi[j]) dv[j] = __ldg(&p[j][mid[j]].x);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (lo[j] < hi[j]) {
              if (dv[j] < tg[j]) lo[j] = mid[j] + 1u; else hi[j] = mid[j];
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t i = i0 + (uint32_t)j * (uint32_t)nthr;
          if (i < n_ent) sh.bounds[i] = al[j] + lo[j];
        }
      }
      __syncthreads();

      // ---- slices of this window, round-robin over the warps ------------------------------------
      for (uint32_t s = (uint32_t)warp; s < win_n; s += (uint32_t)NW) {
        const uint32_t* blo = sh.bounds + s * L;
        uint32_t my_lo = 0u, my_hi = 0u;
        if ((uint32_t)lane < L) { my_lo = blo[lane]; my_hi = blo[L + lane]; }
        unsigned todo = __ballot_sync(0xFFFFFFFFu, my_lo < my_hi);            // leaves with postings in here
        if (tp.prefetch) {
          // Memory-level parallelism: the table knows exactly which postings this warp will need for
          // its slice `prefetch` rounds from now; pull those lines into L2 while this slice is worked on.
          const uint32_t s2 = s + tp.prefetch * (uint32_t)NW;
          if (s2 < win_n) {
            uint32_t pa = 0u, pb = 0u;
            if ((uint32_t)lane < L) { pa = sh.bounds[s2 * L + lane]; pb = sh.bounds[(s2 + 1u) * L + lane]; }
            unsigned pt = __ballot_sync(0xFFFFFFFFu, pa < pb);
            while (pt) {
              const int pl = __ffs(pt) - 1;
              pt &= pt - 1u;
              const uint32_t a = __shfl_sync(0xFFFFFFFFu, pa, pl), b = __shfl_sync(0xFFFFFFFFu, pb, pl);
              const unsigned long long pbase = __shfl_sync(0xFFFFFFFFu, lf_base, pl);
              // 128-byte lines (16 postings) that hold [a, b)
              for (uint32_t line = (a >> 4) + (uint32_t)lane; line <= ((b - 1u) >> 4); line += 32u) prefetch_l2(tp.pairs + pbase + (line << 4));
            }
          }
        }
        if (todo == 0u) continue;
        if (!simple_or && __ballot_sync(0xFFFFFFFFu, ((uint32_t)lane < G) && ((todo & gm) == 0u))) continue;   // an AND needs every group
        const unsigned todo0 = todo;
        const uint32_t sub_lo = d_lo + (win_lo + s) * SW;
        const uint32_t n_slots = min(SW, d_hi - sub_lo);
        cx.sbase = cx.slots_addr - (sub_lo << shift);
        // until k hits exist (threshold still FLT_MIN) nothing is pushed: the slice is scanned instead
        const uint32_t tb = lds_u32(thr_addr);               // other warps raise it
        // postings of this slice (the table knows): a sparse slice is cheaper to push than to scan, and
        // cheaper to clear by walking its postings again than by zeroing every slot
        uint32_t n_post = ((uint32_t)lane < L) ? my_hi - my_lo : 0u;
        n_post = __reduce_add_sync(0xFFFFFFFFu, n_post);
        const bool boot = simple_or && (tb == 0x00800000u) && n_post > (uint32_t)TM_HOT;
        cx.thr = boot ? __uint_as_float(0x7F800000u) : __uint_as_float(tb);

        // ---- visits.  A leaf with at most 32 postings in the slice is one load per lane, issued
        // before the previous leaf is processed; longer ranges stream super-rows.
        int l = __ffs(todo) - 1;                 // ascending leaf order = ascending group rank
        todo &= todo - 1u;
        uint32_t lo = __shfl_sync(0xFFFFFFFFu, my_lo, l), hi = __shfl_sync(0xFFFFFFFFu, my_hi, l);
        const uint2* __restrict__ pairs = tp.pairs + __shfl_sync(0xFFFFFFFFu, lf_base, l);
        uint2 r = make_uint2(0u, 0u);
        if (hi - lo <= 32u && (uint32_t)lane < hi - lo) r = ldg_pair(pairs + lo + (uint32_t)lane);
        for (;;) {
          // the next leaf, and its postings if they are few
          int l2 = -1;
          uint32_t lo2 = 0u, hi2 = 0u;
          const uint2* __restrict__ pairs2 = pairs;
          uint2 r2 = make_uint2(0u, 0u);
          if (todo) {
            l2 = __ffs(todo) - 1;
            todo &= todo - 1u;
            lo2 = __shfl_sync(0xFFFFFFFFu, my_lo, l2);
            hi2 = __shfl_sync(0xFFFFFFFFu, my_hi, l2);
            pairs2 = tp.pairs + __shfl_sync(0xFFFFFFFFu, lf_base, l2);
            if (hi2 - lo2 <= 32u && (uint32_t)lane < hi2 - lo2) r2 = ldg_pair(pairs2 + lo2 + (uint32_t)lane);
          }
          const float w = __shfl_sync(0xFFFFFFFFu, lf_w, l);
          const uint32_t g = __shfl_sync(0xFFFFFFFFu, lf_grp, l);
          const bool lastg = (g + 1u == G);
          if (hi - lo <= 32u) {
            if ((uint32_t)lane < hi - lo) {
              if (simple_or) or_one(cx, w, r.x, r.y, tot);
              else and_one<false>(cx, w, g, lastg, r.x, r.y, tot);
            }
          } else {
            uint32_t i0 = lo & ~127u;
            const uint32_t i_last = (hi - 1u) & ~127u;
            uint2 qa[4], qb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) qa[e] = ldg_pair(pairs + i0 + (uint32_t)lane + 32u * e);
            for (;;) {
              const bool have_next = i0 < i_last;
              if (have_next) {
#pragma unroll
                for (int e = 0; e < 4; ++e) qb[e] = ldg_pair(pairs + i0 + 128u + (uint32_t)lane + 32u * e);
              }
              if (i0 >= lo && i0 + 128u <= hi) {           // interior super-row: no masks
                if (simple_or) or_four(cx, w, qa, tot);
                else and_four<false>(cx, w, g, lastg, qa, tot);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint32_t idx = i0 + (uint32_t)lane + 32u * e;
                  if (idx >= lo && idx < hi) {
                    if (simple_or) or_one(cx, w, qa[e].x, qa[e].y, tot);
                    else and_one<false>(cx, w, g, lastg, qa[e].x, qa[e].y, tot);
                  }
                }
              }
              if (!have_next) break;
#pragma unroll
              for (int e = 0; e < 4; ++e) qa[e] = qb[e];
              i0 += 128u;
            }
          }
          if (l2 < 0) break;
          l = l2; lo = lo2; hi = hi2; pairs = pairs2; r = r2;
        }

        // ---- slice epilogue: only documents that crossed the threshold are looked at -------------
        __syncwarp();
        const uint32_t nhot = boot ? 0xFFFFFFFFu : lds_u32(cx.cnt_addr);
        if (nhot) {
          const float fthr = boot ? 1.17549435e-38f : cx.thr;
          const bool overflow = nhot > (uint32_t)TM_HOT;
          const uint32_t n = overflow ? n_slots : nhot;
          for (uint32_t j0 = 0; j0 < n; j0 += 32u) {
            const uint32_t j = j0 + (uint32_t)lane;
            unsigned long long key = 0ull;
            if (j < n) {
              const uint32_t slot = overflow ? j : lds_u16(cx.hot_addr + j * 2u);
              float sc;
              bool ok;
              if (simple_or) {
                sc = lds_f32(cx.slots_addr + (slot << 2));
                ok = sc != 0.0f;
              } else {
                const uint2 v = lds_v2(cx.slots_addr + (slot << 3));
                sc = __uint_as_float(v.y);
                ok = v.x == G;
              }
              if (ok && sc >= fthr) key = make_key(sc, tp.doc_base + sub_lo + slot);
            }
            unsigned pm = __ballot_sync(0xFFFFFFFFu, key > thr_key);
            while (pm) {
              const int src = __ffs(pm) - 1;
              pm &= pm - 1u;
              const unsigned long long bk = __shfl_sync(0xFFFFFFFFu, key, src);
              if (bk > thr_key) {
                warp_topk_insert(top, bk, lane);
                thr_key = __shfl_sync(0xFFFFFFFFu, top, tp.k - 1);
              }
            }
          }
          if (lane == 0) {
            if (!boot) sts_u32(cx.cnt_addr, 0u);
            if (thr_key != 0ull) atomicMax(&sh.thr_bits, __float_as_uint(key_score(thr_key)));
          }
        }
        if (n_post <= 96u) {
          // clear by walking the slice's postings again (L1 / L2 hits): zero stores only where needed
          unsigned cl = todo0;
          while (cl) {
            const int l3 = __ffs(cl) - 1;
            cl &= cl - 1u;
            const uint32_t lo3 = __shfl_sync(0xFFFFFFFFu, my_lo, l3), hi3 = __shfl_sync(0xFFFFFFFFu, my_hi, l3);
            const uint2* __restrict__ p3 = tp.pairs + __shfl_sync(0xFFFFFFFFu, lf_base, l3);
            for (uint32_t i = lo3 + (uint32_t)lane; i < hi3; i += 32u) {
              const uint32_t a = cx.sbase + (ldg_pair(p3 + i).x << shift);
              if (simple_or) sts_f32(a, 0.0f); else sts_v2(a, 0u, 0u);
            }
          }
        } else {
          const uint32_t bytes = n_slots << shift;
          uint32_t o = (uint32_t)lane * 16u;
          for (; o + 1536u < bytes; o += 2048u) {
            sts_zero16(cx.slots_addr + o);
            sts_zero16(cx.slots_addr + o + 512u);
            sts_zero16(cx.slots_addr + o + 1024u);
            sts_zero16(cx.slots_addr + o + 1536u);
          }
          for (; o < bytes; o += 512u) sts_zero16(cx.slots_addr + o);
        }
        __syncwarp();
      }
    }

    // ---- item epilogue: merge the warps' lists ---------------------------------------------------
    sh.keys[warp][lane] = (lane < tp.k) ? top : 0ull;
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xFFFFFFFFu, tot, o);
    if (lane == 0 && tot) atomicAdd(&sh.total, (unsigned long long)tot);
    __syncthreads();
    if (warp == 0) {
      unsigned long long best = sh.keys[0][lane], bthr = 0ull;
      if (tp.k <= 32) bthr = __shfl_sync(0xFFFFFFFFu, best, tp.k - 1);
      for (int w2 = 1; w2 < NW; ++w2) {
        const unsigned long long key = sh.keys[w2][lane];
        unsigned pm = __ballot_sync(0xFFFFFFFFu, key > bthr);
        while (pm) {
          const int src = __ffs(pm) - 1;
          pm &= pm - 1u;
          const unsigned long long bk = __shfl_sync(0xFFFFFFFFu, key, src);
          if (bk > bthr) {
            warp_topk_insert(best, bk, lane);
            bthr = __shfl_sync(0xFFFFFFFFu, best, tp.k - 1);
          }
        }
      }
      if (lane < tp.k) tp.part_keys[(size_t)item.part * tp.k + lane] = best;
      if (lane == 0 && sh.total) atomicAdd(tp.totals + item.q, sh.total);
    }
  }
}
