// k_score_ring — k_score_stream with asynchronous posting rings (included by bm25f.cu inside its
// anonymous namespace, after stream.cuh whose helpers and parameter block it shares).
//
// Same work items, accumulators, hot list and top-k as k_score_stream.  What changes is how postings
// reach the warp.  k_score_stream parks ONE row (32 postings) per leaf in shared memory and fetches
// every further row of a visit with a load it then waits for, so a (sub-range, leaf) visit of the
// typical ~90 postings pays two or three dependent memory latencies with one load in flight per warp.
// Here every leaf owns a RING of R rows in shared memory (RG_ROWS rows per warp shared out between the
// leaves), filled with cp.async (LDGSTS: global -> shared, no registers, no wait).  A visit consumes rows
// from the ring; when it is done, the slots of the rows it finished are refilled with the rows that
// follow the ring, and the warp moves on to the next leaf.  The one `cp.async.wait_all` sits at the top
// of the NEXT sub-range, a whole sub-range of work after the requests were issued.  Every lane reads
// back exactly the 8 bytes it requested itself, so no barrier is needed either.
// A visit that outruns its ring (dense lists) falls through to the register-streamed super-rows of
// k_score_stream and re-seeds the ring where it stops.
#pragma once

constexpr uint32_t RG_ROWS = 16;        // ring rows (256 B each) per warp, shared out between the query's leaves

__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

// Requires: k <= 32 * KR, <= ST_MAX_LEAVES leaves, every leaf weight > 0, no after_key, no postings of
// deleted documents in the store (bm25f_create compacts them away).
template <int KR>
__global__ void __maxnreg__(80) k_score_ring(StreamParams sp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const uint32_t slot_bytes = sp.slot_bytes;
  // shared memory: [nwarps][slot_bytes] slots | [nwarps][RG_ROWS][32] ring rows (8 B) | [nwarps][ST_HOT] hot (2 B) | [nwarps] hot counters
  SubCtx cx;
  const uint32_t smem0 = smem_u32(smem_raw);
  cx.slots_addr = smem0 + (uint32_t)warp * slot_bytes;
  const uint32_t rings_addr = smem0 + (uint32_t)nwarps * slot_bytes + (uint32_t)warp * (RG_ROWS * 256u) + (uint32_t)lane * 8u;   // this lane's column
  cx.hot_addr = smem0 + (uint32_t)nwarps * (slot_bytes + RG_ROWS * 256u) + (uint32_t)warp * (ST_HOT * 2);
  cx.cnt_addr = smem0 + (uint32_t)nwarps * (slot_bytes + RG_ROWS * 256u + ST_HOT * 2) + (uint32_t)warp * 4u;

  for (uint32_t o = (uint32_t)lane * 16u; o < slot_bytes; o += 512u) sts_zero16(cx.slots_addr + o);
  if (lane == 0) sts_u32(cx.cnt_addr, 0u);
  __syncwarp();

  for (;;) {
    uint32_t item_idx = 0;
    if (lane == 0) item_idx = atomicAdd(sp.queue, 1u);
    item_idx = __shfl_sync(0xFFFFFFFFu, item_idx, 0);
    if (item_idx >= sp.n_items) break;

    const ItemRec item = sp.items[item_idx];
    const QueryRec q = sp.queries[item.q];
    const int L = (int)q.n_leaves;
    const uint32_t G = q.n_groups;
    const bool simple_or = (q.flags & QF_SIMPLE_OR) != 0;
    const uint32_t shift = simple_or ? 2u : 3u;
    const uint32_t SW = slot_bytes >> shift;            // documents per sub-range
    const uint32_t d_lo = item.tile_begin, d_hi = item.tile_end;

    // ---- leaf state: lane l holds leaf l ------------------------------------------------------
    // Index space of a leaf: absolute posting index minus s_base, where s_base = off & ~31, so every
    // row is 256-byte aligned.  The list occupies [off & 31, s_end).
    // Rows [s_cur >> 5, s_fill) of the leaf are in its ring (or on their way); row r sits in slot r % R.
    unsigned long long s_base = 0ull;
    uint32_t s_cur = 0u, s_end = 0u, s_grp = 0u, s_next = 0xFFFFFFFFu, s_fill = 0u;
    float s_w = 0.0f;
    const uint32_t rshift = L <= 1 ? 4u : L <= 2 ? 3u : L <= 4 ? 2u : 1u;   // log2 of R, the ring rows per leaf
    const uint32_t rmask = (1u << rshift) - 1u;
    cp_async_wait_all();                        // nothing of the previous item may still be landing in the rings
    if (lane < L) {
      const LeafRec lf = sp.leaves[q.leaf_begin + lane];
      const uint32_t a = (uint32_t)(lf.off & 31ull);
      s_base = lf.off - a;
      s_cur = a;
      s_end = a + lf.df;
      s_w = lf.w;
      s_grp = lf.group;
    }
    for (int l = 0; l < L; ++l) {
      const unsigned long long base = __shfl_sync(0xFFFFFFFFu, s_base, l);
      const uint32_t end = __shfl_sync(0xFFFFFFFFu, s_end, l);
      uint32_t cur = __shfl_sync(0xFFFFFFFFu, s_cur, l);
      const uint2* __restrict__ pairs = sp.pairs + base;
      if (d_lo > 0u) cur += warp_lower_bound(pairs + cur, end - cur, d_lo, lane);
      // seed the ring with the row that holds the cursor and the R - 1 rows after it
      const uint32_t ring = rings_addr + (((uint32_t)l << rshift) << 8);
      const uint32_t row0 = cur >> 5;
      const uint32_t lim = min(row0 + rmask + 1u, (end + 31u) >> 5);
      for (uint32_t rr = row0; rr < lim; ++rr) {
        const uint32_t idx = (rr << 5) + (uint32_t)lane;
        const uint32_t dst = ring + ((rr & rmask) << 8);
        if (idx < end) cp_async_8(dst, pairs + idx);
        else sts_v2(dst, 0xFFFFFFFFu, 0u);       // past the list: never inside a sub-range
      }
      if (lane == l) { s_cur = cur; s_fill = max(lim, row0); }
      if (sp.pf_dist) {
        // chunks [cur, cur + pf_dist + chunk), one per lane
        const uint32_t c0 = (cur & ~(ST_PF_CHUNK - 1u)) + (uint32_t)lane * ST_PF_CHUNK;
        if (c0 < end && c0 <= cur + sp.pf_dist) bulk_prefetch_l2(pairs + c0, min(ST_PF_CHUNK, (end - c0 + 1u) & ~1u) * 8u);
      }
    }

    // the docid at every leaf's cursor, one load for all the leaves
    if (lane < L && s_cur < s_end) s_next = __ldg(&(sp.pairs + s_base)[s_cur].x);

    unsigned long long top[KR];               // lane i, row j: the (32 j + i)-th best key of this item so far
#pragma unroll
    for (int j = 0; j < KR; ++j) top[j] = 0ull;
    unsigned long long thr_key = 0ull;
    cx.thr = 1.17549435e-38f;                 // FLT_MIN until k hits exist: every first hit is hot
    unsigned int tot = 0;

    uint32_t sub_lo = d_lo;
    while (sub_lo < d_hi) {
      cx.sub_hi = min(sub_lo + SW, d_hi);
      unsigned todo = __ballot_sync(0xFFFFFFFFu, s_next < cx.sub_hi);      // leaves with a posting in here
      if (todo == 0u) {
        // nothing in this sub-range: jump to the one that holds the nearest posting
        const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, s_next);
        if (m >= d_hi) break;
        sub_lo += ((m - sub_lo) / SW) * SW;
        continue;
      }
      cx.sbase = cx.slots_addr - (sub_lo << shift);
      cp_async_wait_all();                      // the rows requested while the previous sub-range was swept

      while (todo) {
        const int l = __ffs(todo) - 1;          // ascending leaf order = ascending group rank
        todo &= todo - 1u;
        const unsigned long long base = __shfl_sync(0xFFFFFFFFu, s_base, l);
        const uint32_t end = __shfl_sync(0xFFFFFFFFu, s_end, l);
        uint32_t cur = __shfl_sync(0xFFFFFFFFu, s_cur, l);
        const float w = __shfl_sync(0xFFFFFFFFu, s_w, l);
        const uint32_t g = __shfl_sync(0xFFFFFFFFu, s_grp, l);
        const bool lastg = (g + 1u == G);
        const uint2* __restrict__ pairs = sp.pairs + base;
        uint32_t fill = __shfl_sync(0xFFFFFFFFu, s_fill, l);
        const uint32_t ring = rings_addr + (((uint32_t)l << rshift) << 8);

        uint2 r = make_uint2(0xFFFFFFFFu, 0u);  // ends up as the row that holds the cursor
        uint32_t off = cur & 31u;               // lanes before the cursor in its row are already consumed
        for (;;) {
          const uint32_t row = cur >> 5;
          if (row < fill) {
            // ---- up to four rows out of the ring at once: the lanes at or after the cursor whose docid is
            // inside the sub-range form one contiguous run over the rows (lists are sorted)
            uint2 q4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              q4[e] = make_uint2(0xFFFFFFFFu, 0u);
              if (row + e < fill) q4[e] = lds_v2(ring + (((row + e) & rmask) << 8));
            }
            uint32_t n = 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const bool act = (q4[e].x < cx.sub_hi) && (e > 0 || (uint32_t)lane >= off);
              n += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, act));
              if (act) {
                if (simple_or) or_one(cx, w, q4[e].x, q4[e].y, tot);
                else and_one<true>(cx, w, g, lastg, q4[e].x, q4[e].y, tot);
              }
            }
            const uint32_t cap = min(fill - row, 4u) * 32u - off;
            cur += n;
            const uint32_t es = (cur >> 5) - row;
            r = es == 0u ? q4[0] : es == 1u ? q4[1] : es == 2u ? q4[2] : q4[3];
            if (n < cap || cur >= end) break;   // the sub-range (or the list) ends inside these rows
            off = 0u;
            continue;
          }
          // ---- the ring is dry (a dense list): rows straight from memory.  Super-rows while whole ones
          // fit, each requested one super-row before it is processed
          if ((cur & 127u) == 0u && cur + 128u <= end) {
            uint2 qa[4], qb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) qa[e] = ldg_pair(pairs + cur + (uint32_t)lane + 32u * e);
            bool stop = false;
            for (;;) {
              const bool have_next = cur + 256u <= end;
              if (have_next) {
#pragma unroll
                for (int e = 0; e < 4; ++e) qb[e] = ldg_pair(pairs + cur + 128u + (uint32_t)lane + 32u * e);
              }
              if (sp.pf_dist && (cur & (ST_PF_CHUNK - 1u)) == 0u) {
                const uint32_t c0 = cur + sp.pf_dist;
                if (lane == 0 && c0 < end) bulk_prefetch_l2(pairs + c0, min(ST_PF_CHUNK, (end - c0 + 1u) & ~1u) * 8u);
              }
              const uint32_t dlast = __shfl_sync(0xFFFFFFFFu, qa[3].x, 31);
              if (dlast < cx.sub_hi) {          // entirely inside the sub-range: no masks
                if (simple_or) or_four(cx, w, qa, tot);
                else and_four<true>(cx, w, g, lastg, qa, tot);
                cur += 128u;
                if (!have_next) break;          // fewer than 128 postings left: back to single rows
#pragma unroll
                for (int e = 0; e < 4; ++e) qa[e] = qb[e];
              } else {
                // the sub-range ends inside this super-row
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (!stop) {
                    const bool a2 = qa[e].x < cx.sub_hi;
                    const unsigned m2 = __ballot_sync(0xFFFFFFFFu, a2);
                    if (a2) {
                      if (simple_or) or_one(cx, w, qa[e].x, qa[e].y, tot);
                      else and_one<true>(cx, w, g, lastg, qa[e].x, qa[e].y, tot);
                    }
                    const uint32_t n2 = (uint32_t)__popc(m2);
                    cur += n2;
                    if (n2 < 32u) { r = qa[e]; stop = true; }
                  }
                }
                break;
              }
            }
            if (stop || cur >= end) break;      // r holds the row of the cursor (or the list is finished)
            continue;
          }
          // ---- one row
          {
            const uint32_t idx = (row << 5) + (uint32_t)lane;
            r = make_uint2(0xFFFFFFFFu, 0u);
            if (idx < end) r = ldg_pair(pairs + idx);
            const bool act = ((uint32_t)lane >= off) && (r.x < cx.sub_hi);
            const unsigned mk = __ballot_sync(0xFFFFFFFFu, act);
            if (act) {
              if (simple_or) or_one(cx, w, r.x, r.y, tot);
              else and_one<true>(cx, w, g, lastg, r.x, r.y, tot);
            }
            const uint32_t n = (uint32_t)__popc(mk);
            cur += n;
            if (n == 0u || (cur & 31u) != 0u || cur >= end) break;
            off = 0u;
          }
        }
        // ---- refill: the slots of the rows this visit finished take the rows that follow the ring
        {
          const uint32_t row0 = cur >> 5;
          fill = max(fill, row0);               // a visit that streamed past its ring re-seeds it at the cursor
          const uint32_t lim = min(row0 + rmask + 1u, (end + 31u) >> 5);
          for (uint32_t rr = fill; rr < lim; ++rr) {
            const uint32_t idx = (rr << 5) + (uint32_t)lane;
            const uint32_t dst = ring + ((rr & rmask) << 8);
            if (idx < end) cp_async_8(dst, pairs + idx);
            else sts_v2(dst, 0xFFFFFFFFu, 0u);
            if (sp.pf_dist && lane < 2 && idx + sp.pf_dist < end) prefetch_l2(pairs + (rr << 5) + sp.pf_dist + (uint32_t)lane * 16u);
          }
          fill = max(fill, lim);
        }
        const uint32_t nd = __shfl_sync(0xFFFFFFFFu, r.x, cur & 31u);
        if (lane == l) { s_cur = cur; s_fill = fill; s_next = (cur < end) ? nd : 0xFFFFFFFFu; }
      }

      // ---- sub-range epilogue ------------------------------------------------------------------
      __syncwarp();
      const uint32_t nhot = lds_u32(cx.cnt_addr);
      if (nhot) {
        const bool overflow = nhot > (uint32_t)ST_HOT;
        const uint32_t n = overflow ? (cx.sub_hi - sub_lo) : nhot;
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
            if (ok && sc >= cx.thr) key = make_key(sc, sp.doc_base + sub_lo + slot);
          }
          unsigned pm = __ballot_sync(0xFFFFFFFFu, key > thr_key);
          while (pm) {
            const int src = __ffs(pm) - 1;
            pm &= pm - 1u;
            const unsigned long long bk = __shfl_sync(0xFFFFFFFFu, key, src);
            if (bk > thr_key) {
              warp_topk_insert_rows<KR>(top, bk, lane);
              thr_key = warp_topk_kth<KR>(top, sp.k);
            }
          }
        }
        if (thr_key != 0ull) cx.thr = key_score(thr_key);
        if (lane == 0) sts_u32(cx.cnt_addr, 0u);
      }
      for (uint32_t o = (uint32_t)lane * 16u; o < slot_bytes; o += 512u) sts_zero16(cx.slots_addr + o);
      __syncwarp();
      sub_lo += SW;
    }

    // ---- item epilogue -------------------------------------------------------------------------
    unsigned long long* out = sp.part_keys + (size_t)item.part * sp.k;
#pragma unroll
    for (int j = 0; j < KR; ++j)
      if (32 * j + lane < sp.k) out[32 * j + lane] = top[j];
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xFFFFFFFFu, tot, o);
    if (lane == 0 && tot) atomicAdd(sp.totals + item.q, (unsigned long long)tot);
  }
}
