// k_score_hash — flat OR with one dominant leaf (included by bm25f.cu after isect.cuh).
//
// In an OR, a document that occurs in exactly one leaf needs no accumulator: its score is
// w * impact.  When the densest leaf D of a query outweighs all the other leaves S together, almost
// every posting is such a document.  This kernel therefore keeps accumulators only for the documents
// of S, in a small per-warp HASH TABLE (open addressing, {docid, score} entries in shared memory), and
// streams D past it:
//
//  * every WARP is an independent worker pulling items (query, document range) from a global counter;
//  * the range is swept in BLOCKS whose width is chosen per query so that about HS_TARGET postings
//    of S fall into a block (a query is ~10-60 blocks, not hundreds of fixed sub-ranges);
//  * per block: the postings of every S leaf in the block are inserted / accumulated (atomicCAS on
//    the key claims a slot; docids are unique inside a list, so the score update needs no atomic);
//    then D's postings of the block are streamed in super-rows of 128 (four 64-bit loads in flight per
//    lane): test the block's exact BITMAP of S documents (one shared load + a bit test) - nearly always
//    clear, then score = w * impact is compared with the k-th best score and that is all; only a set
//    bit walks the hash chain and raises the entry's score instead; finally the table is flushed: every
//    entry is a match and is offered to the top-k;
//  * matches = |D| + |entries| - |hits|, exactly; every posting is read exactly once.
//
// Results are identical to the other kernels (same impacts, FMA order: S leaves in leaf order, D last).
#pragma once

constexpr int HS_WARPS = 4;       // 32 KB of tables per CTA: several CTAs per SM
constexpr uint32_t HS_SLOTS = 1024;      // table entries per warp (8 KB)
constexpr uint32_t HS_TARGET = 320;      // postings of S aimed at per block
constexpr uint32_t HS_MAX_FILL = 640;    // a block with more is halved
constexpr uint32_t HS_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t HS_BITMAP_DOCS = 16384; // widest block: one bit per document of the block (2 KB per warp)

struct HashParams {
  const uint2* pairs;
  const LeafRec* leaves;
  const QueryRec* queries;
  const ItemRec* items;            // tile_begin / tile_end hold the item's document range [lo, hi)
  unsigned long long* part_keys;   // [n_parts * k]
  unsigned long long* totals;      // [Q]
  unsigned int* queue;             // work counter, zeroed before the launch
  uint32_t n_items;
  uint32_t doc_base;
  uint32_t n_docs;
  int k;
};

__device__ __forceinline__ uint32_t hs_hash(uint32_t d) { return (d * 2654435761u) >> 22; }   // 10 bits
__device__ __forceinline__ uint32_t atoms_cas(uint32_t addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

// Requires: flat OR (QF_SIMPLE_OR | QF_STREAM_LAST: the densest leaf is the last one), k <= 32,
// <= 32 leaves, every leaf weight > 0, no after_key, no postings of deleted documents in the store.
__global__ void __launch_bounds__(HS_WARPS * 32) k_score_hash(HashParams hp) {
  __shared__ __align__(16) uint2 s_table[HS_WARPS][HS_SLOTS];
  __shared__ __align__(16) uint32_t s_bits[HS_WARPS][HS_BITMAP_DOCS / 32];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint2* __restrict__ store = hp.pairs;
  const uint32_t tab = smem_u32(&s_table[warp][0]);
  const uint32_t bits = smem_u32(&s_bits[warp][0]);

  for (uint32_t j = (uint32_t)lane; j < HS_SLOTS; j += 32u) sts_v2(tab + j * 8u, HS_EMPTY, 0u);
  for (uint32_t o = (uint32_t)lane * 16u; o < HS_BITMAP_DOCS / 8u; o += 512u) sts_zero16(bits + o);
  __syncwarp();

  for (;;) {
    uint32_t item_idx = 0;
    if (lane == 0) item_idx = atomicAdd(hp.queue, 1u);
    item_idx = __shfl_sync(0xFFFFFFFFu, item_idx, 0);
    if (item_idx >= hp.n_items) break;

    const ItemRec item = hp.items[item_idx];
    const QueryRec q = hp.queries[item.q];
    const int L = (int)q.n_leaves;
    const int dl = L - 1;                                   // the dense leaf
    const uint32_t d_lo = item.tile_begin, d_hi = item.tile_end;

    // lane l keeps leaf l: list origin, length, cursor (first posting at or after the block), weight
    unsigned long long s_off = 0ull;
    uint32_t s_df = 0u, s_cur = 0u;
    float s_w = 0.0f;
    if (lane < L) {
      const LeafRec lf = hp.leaves[q.leaf_begin + lane];
      s_off = lf.off;
      s_df = lf.df;
      s_w = lf.w;
    }
    if (d_lo > 0u) {
      for (int l = 0; l < L; ++l) {
        const uint32_t st = warp_lower_bound(store + __shfl_sync(0xFFFFFFFFu, s_off, l), __shfl_sync(0xFFFFFFFFu, s_df, l), d_lo, lane);
        if (lane == l) s_cur = st;
      }
    }
    // block width: about HS_TARGET postings of the sparse leaves per block (documents are spread evenly
    // in the synthetic corpora; a block that turns out fuller than HS_MAX_FILL is halved below)
    uint32_t s_total = (lane < dl) ? s_df : 0u;
    s_total = __reduce_add_sync(0xFFFFFFFFu, s_total);
    uint32_t W = min(d_hi - d_lo, HS_BITMAP_DOCS);
    if (s_total > 0u) {
      const unsigned long long w64 = (unsigned long long)hp.n_docs * HS_TARGET / s_total;
      W = (uint32_t)min((unsigned long long)W, max(w64, 64ull));
    }
    uint32_t s_win = 64u;                     // lane l: size of leaf l's previous block (where to look first)
    const uint2* __restrict__ dp = store + __shfl_sync(0xFFFFFFFFu, s_off, dl);
    const uint32_t d_df = __shfl_sync(0xFFFFFFFFu, s_df, dl);
    const float d_w = __shfl_sync(0xFFFFFFFFu, s_w, dl);

    unsigned long long top = 0ull;            // lane i: i-th best key of this item so far
    unsigned long long thr_key = 0ull;
    float thr = 0.0f;
    unsigned int tot = 0;                     // per-lane partial match count

    uint32_t blo = d_lo;
    while (blo < d_hi) {
      // ---- block bounds: first posting of every leaf at or after the block's end --------------------
      uint32_t bw = min(W, d_hi - blo);
      uint32_t bhi, s_end, n_s;
      for (;;) {
        bhi = blo + bw;
        s_end = s_cur;
        for (int l = 0; l < L; ++l) {
          const uint32_t cur = __shfl_sync(0xFFFFFFFFu, s_cur, l);
          const uint32_t df = __shfl_sync(0xFFFFFFFFu, s_df, l);
          uint32_t e = cur;
          if (cur < df) {
            // blocks of one query are about the same size: look in twice the previous block first
            const uint2* __restrict__ lp = store + __shfl_sync(0xFFFFFFFFu, s_off, l) + cur;
            const uint32_t near = min(df - cur, 2u * __shfl_sync(0xFFFFFFFFu, s_win, l) + 32u);
            uint32_t n = warp_lower_bound(lp, near, bhi, lane);
            if (n == near && near < df - cur) n = near + warp_lower_bound(lp + near, df - cur - near, bhi, lane);
            e = cur + n;
          }
          if (lane == l) s_end = e;
        }
        n_s = (lane < dl) ? s_end - s_cur : 0u;
        n_s = __reduce_add_sync(0xFFFFFFFFu, n_s);
        if (n_s <= HS_MAX_FILL || bw <= 1u) break;
        bw = (bw + 1u) >> 1;                                // fuller than the table likes: halve the block
      }

      // ---- the sparse leaves: insert / accumulate ------------------------------------------------
      uint32_t n_ent = 0;                                   // per-lane count of entries created
      for (int l = 0; l < dl; ++l) {
        const uint32_t lo = __shfl_sync(0xFFFFFFFFu, s_cur, l), hi = __shfl_sync(0xFFFFFFFFu, s_end, l);
        if (lo >= hi) continue;
        const uint2* __restrict__ lp = store + __shfl_sync(0xFFFFFFFFu, s_off, l);
        const float w = __shfl_sync(0xFFFFFFFFu, s_w, l);
        for (uint32_t i = lo + (uint32_t)lane; i < hi; i += 32u) {
          const uint2 r = ldg_pair(lp + i);
          const uint32_t rel = r.x - blo;
          asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(bits + (rel >> 5) * 4u), "r"(1u << (rel & 31u)) : "memory");
          uint32_t h = hs_hash(r.x);
          for (;;) {
            const uint32_t a = tab + h * 8u;
            const uint32_t old = atoms_cas(a, HS_EMPTY, r.x);
            if (old == HS_EMPTY) { sts_f32(a + 4u, w * __uint_as_float(r.y)); ++n_ent; break; }
            if (old == r.x) { sts_f32(a + 4u, fmaf(w, __uint_as_float(r.y), lds_f32(a + 4u))); break; }
            h = (h + 1u) & (HS_SLOTS - 1u);
          }
        }
        __syncwarp();                                       // the next leaf may meet the same documents
      }

      // ---- the dense leaf: probe, score, compare --------------------------------------------------
      const uint32_t lo = __shfl_sync(0xFFFFFFFFu, s_cur, dl), hi = __shfl_sync(0xFFFFFFFFu, s_end, dl);
      uint32_t n_hit = 0;
      if (lo < hi) {
        uint32_t i0 = lo & ~127u;
        const uint32_t i_last = (hi - 1u) & ~127u;
        uint2 qa[4], qb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) qa[e] = ldg_pair(dp + i0 + (uint32_t)lane + 32u * e);
        for (;;) {
          const bool have_next = i0 < i_last;
          if (have_next) {
#pragma unroll
            for (int e = 0; e < 4; ++e) qb[e] = ldg_pair(dp + i0 + 128u + (uint32_t)lane + 32u * e);
          }
          const bool inner = i0 >= lo && i0 + 128u <= hi;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t idx = i0 + (uint32_t)lane + 32u * e;
            const bool valid = inner || (idx >= lo && idx < hi);
            const uint32_t d = qa[e].x;
            float score = d_w * __uint_as_float(qa[e].y);
            bool miss = valid;
            if (valid && n_s && ((lds_u32(bits + ((d - blo) >> 5) * 4u) >> ((d - blo) & 31u)) & 1u)) {
              uint32_t h = hs_hash(d);                       // in a sparse leaf too: its entry exists
              for (;;) {
                const uint32_t a = tab + h * 8u;
                const uint2 t = lds_v2(a);
                if (t.x == HS_EMPTY) break;
                if (t.x == d) {                              // also in a sparse leaf: raise its entry
                  sts_f32(a + 4u, fmaf(d_w, __uint_as_float(qa[e].y), __uint_as_float(t.y)));
                  miss = false;
                  ++n_hit;
                  break;
                }
                h = (h + 1u) & (HS_SLOTS - 1u);
              }
            }
            unsigned long long key = 0ull;
            if (miss && score >= thr) key = make_key(score, hp.doc_base + d);
            unsigned pm = __ballot_sync(0xFFFFFFFFu, key > thr_key);
            if (pm) {
              do {
                const int src = __ffs(pm) - 1;
                pm &= pm - 1u;
                const unsigned long long bk = __shfl_sync(0xFFFFFFFFu, key, src);
                if (bk > thr_key) {
                  warp_topk_insert(top, bk, lane);
                  thr_key = __shfl_sync(0xFFFFFFFFu, top, hp.k - 1);
                }
              } while (pm);
              if (thr_key != 0ull) thr = key_score(thr_key);
            }
          }
          if (!have_next) break;
#pragma unroll
          for (int e = 0; e < 4; ++e) qa[e] = qb[e];
          i0 += 128u;
        }
        if (lane == 0) tot += hi - lo;                       // every posting of D in the block is a match ...
      }
      tot += n_ent;                                          // ... so is every entry ...
      tot -= n_hit;                                          // ... and the hits were counted twice (mod 2^32 per lane, summed below)

      // ---- flush: every entry is a match; offer it, empty the slot ---------------------------------
      if (n_s) {
        __syncwarp();
        for (uint32_t j = (uint32_t)lane; j < HS_SLOTS; j += 32u) {
          const uint32_t a = tab + j * 8u;
          const uint2 t = lds_v2(a);
          unsigned long long key = 0ull;
          if (t.x != HS_EMPTY) {
            sts_u32(a, HS_EMPTY);
            const float sc = __uint_as_float(t.y);
            if (sc >= thr) key = make_key(sc, hp.doc_base + t.x);
          }
          unsigned pm = __ballot_sync(0xFFFFFFFFu, key > thr_key);
          if (pm) {
            do {
              const int src = __ffs(pm) - 1;
              pm &= pm - 1u;
              const unsigned long long bk = __shfl_sync(0xFFFFFFFFu, key, src);
              if (bk > thr_key) {
                warp_topk_insert(top, bk, lane);
                thr_key = __shfl_sync(0xFFFFFFFFu, top, hp.k - 1);
              }
            } while (pm);
            if (thr_key != 0ull) thr = key_score(thr_key);
          }
        }
        for (uint32_t o = (uint32_t)lane * 16u; o < ((bhi - blo + 127u) >> 7) * 16u; o += 512u) sts_zero16(bits + o);
        __syncwarp();
      }
      if (lane < L) s_win = s_end - s_cur;
      s_cur = s_end;
      blo = bhi;
    }

    unsigned long long* out = hp.part_keys + (size_t)item.part * hp.k;
    if (lane < hp.k) out[lane] = top;
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xFFFFFFFFu, tot, o);
    if (lane == 0 && tot) atomicAdd(hp.totals + item.q, (unsigned long long)tot);
  }
}
