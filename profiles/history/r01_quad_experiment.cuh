// k_score_quad — k_score_stream with one visit loop of four masked rows (included by bm25f.cu after
// stream.cuh, whose helpers and parameter block it shares).
//
// k_score_stream walks a (sub-range, leaf) visit row by row (one load in flight, then wait) unless the
// visit is long enough for its register-streamed super-rows.  Most visits are short (~90 postings), so the
// kernel is bound by dependent latencies at 16 warps per SM (profiles/r01_notes.md).  Here every round of a
// visit loads the four rows from the cursor's row on, all in flight together, and processes them with
// masks; there is no parked tail, no separate streaming path, and 8 fewer registers of row buffers.  The
// intended geometry is FEWER warps with MORE accumulator bytes each (longer visits, fewer of them), the
// instruction-level parallelism of the four rows standing in for the warps given up.
#pragma once

// Requires: k <= 32 * KR, <= ST_MAX_LEAVES leaves, every leaf weight > 0, no after_key, no postings of
// deleted documents in the store (bm25f_create compacts them away).
template <int KR>
__global__ void __launch_bounds__(ST_MAX_WARPS * 32, 1) k_score_quad(StreamParams sp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const uint32_t slot_bytes = sp.slot_bytes;
  // shared memory: [nwarps][slot_bytes] slots | [nwarps][ST_HOT] hot (2 B) | [nwarps] hot counters
  SubCtx cx;
  const uint32_t smem0 = smem_u32(smem_raw);
  cx.slots_addr = smem0 + (uint32_t)warp * slot_bytes;
  cx.hot_addr = smem0 + (uint32_t)nwarps * slot_bytes + (uint32_t)warp * (ST_HOT * 2);
  cx.cnt_addr = smem0 + (uint32_t)nwarps * (slot_bytes + ST_HOT * 2) + (uint32_t)warp * 4u;

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
    unsigned long long s_base = 0ull;
    uint32_t s_cur = 0u, s_end = 0u, s_grp = 0u, s_next = 0xFFFFFFFFu;
    float s_w = 0.0f;
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
      if (lane == l) s_cur = cur;
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

        // Four rows (128 postings) per round, all four loads in flight before the first is used.  The lanes at
        // or after the cursor whose docid is inside the sub-range form one contiguous run over the rows (lists
        // are sorted), so the number of active lanes is the distance the cursor moves.
        uint2 r = make_uint2(0xFFFFFFFFu, 0u);  // ends up as the row that holds the cursor
        uint32_t off = cur & 31u;               // lanes before the cursor in its row are already consumed
        for (;;) {
          const uint32_t row0 = cur & ~31u;
          uint2 q4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t idx = row0 + 32u * e + (uint32_t)lane;
            q4[e] = make_uint2(0xFFFFFFFFu, 0u);
            if (idx < end) q4[e] = ldg_pair(pairs + idx);
          }
          if (sp.pf_dist && (row0 & (ST_PF_CHUNK - 1u)) < 128u) {
            const uint32_t c0 = (row0 & ~(ST_PF_CHUNK - 1u)) + sp.pf_dist;
            if (lane == 0 && c0 < end) bulk_prefetch_l2(pairs + c0, min(ST_PF_CHUNK, (end - c0 + 1u) & ~1u) * 8u);
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
          cur += n;
          const uint32_t es = (cur - row0) >> 5;
          r = es == 0u ? q4[0] : es == 1u ? q4[1] : es == 2u ? q4[2] : q4[3];
          if (n < 128u - off || cur >= end) break;   // the sub-range (or the list) ends inside these rows
          off = 0u;
        }
        const uint32_t nd = __shfl_sync(0xFFFFFFFFu, r.x, cur & 31u);
        if (lane == l) { s_cur = cur; s_next = (cur < end) ? nd : 0xFFFFFFFFu; }
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
