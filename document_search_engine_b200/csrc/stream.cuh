// k_score_stream — the hot kernel of libbm25f (included by bm25f.cu inside its anonymous namespace).
//
// Replaces Whoosh's matcher loop + BM25FScorer + TopCollector for one batch of lowered queries
// (reference call sites my_flask.py:208, :211, :304).
//
// Design (DESIGN.md §4 has the numbers):
//  * PERSISTENT grid, one CTA per SM, every WARP an independent worker that pulls work items
//    (query, document range) heaviest-first from a global counter.  No CTA barriers in the loop.
//  * The posting store is an array of 8-byte PAIRS {docid u32, impact f32} with
//    impact = tf / (tf + norm[field][length byte]) precomputed by bm25f_set_weighting, so the score
//    of a posting is one FMA:  acc = w * impact + acc  with  w = idf * (K1 + 1) * boost, and a row
//    of 32 postings is one coalesced 64-bit load per lane.
//  * Each warp owns `slot_bytes` of shared memory as dense accumulators for the sub-range of
//    documents it is sweeping: 4-byte score slots for a flat OR, 8-byte {groups matched, score}
//    slots for AND / AND-of-OR.  Slots are cleared with 128-bit stores after a sub-range that
//    touched them.
//  * The docid at the cursor of leaf l lives in LANE l of the warp, so "which leaves have a posting in
//    this sub-range" is one compare + ballot and a leaf that has none costs nothing; the rest of the
//    leaf's record (list origin, cursor, end, weight, group) sits in shared memory and is read with one
//    broadcast load per visit.  The partially consumed row of every leaf (its TAIL) is parked
//    in shared memory between visits; a visit that runs past its tail streams SUPER-ROWS (128
//    postings, four loads in flight per lane, the next super-row requested before the current one
//    is processed) straight into registers, with bulk L2 prefetches (cp.async.bulk.prefetch.L2,
//    SASS UBLKPF) `pf_dist` postings ahead.
//  * There is ONE copy of the visit code (no unrolling over leaves): the whole kernel is about
//    1k instructions, because 16 warps at 16 different places of a 90 KB kernel starve on
//    instruction fetch (measured: profiles/r01_notes.md).
//  * Matches are counted while accumulating.  A document is looked at for the top-k only when its
//    running score crosses the current k-th best score; the warp keeps the k best 64-bit keys
//    in registers (KR per lane, k <= 32 * KR) and inserts with shuffles.
#pragma once

constexpr int ST_HOT = 64;              // hot-list entries per warp
constexpr int ST_MAX_WARPS = 16;
constexpr uint32_t ST_PF_CHUNK = 512;   // postings per bulk L2 prefetch (2 KB per array)

constexpr int ST_MAX_LEAVES = 8;        // tails parked in shared memory per warp

struct StreamParams {
  const uint2* pairs;              // {docid, impact bits} per posting
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
  uint32_t pf_dist;                // L2 prefetch distance in postings (0: off)
  int k;
  // FINAL instantiation only (a weighting with a final() step, final.cuh):
  const double* final_add;         // [n_docs] date term of the document (NaN: no date)
  const double* final_blk;         // [n_docs / 32 + 2] largest date term of each aligned block of 32 documents (-inf: none)
  unsigned int* part_lo;           // [n_parts * k] low halves of the keys; part_keys holds the high halves
};

__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint2 ldg_pair(const uint2* p) {
  uint2 v;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t atoms_inc(uint32_t addr) {
  uint32_t v;
  asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts_zero16(uint32_t addr) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}

// First index in d[0, n) whose docid is >= target; all lanes call, 32 probes per round.
__device__ __forceinline__ uint32_t warp_lower_bound(const uint2* __restrict__ d, uint32_t n, uint32_t target, int lane) {
  uint32_t lo = 0, hi = n;
  while (hi - lo > 32u) {
    const uint32_t step = (hi - lo + 31u) >> 5;
    const uint32_t p = lo + (uint32_t)(lane + 1) * step - 1u;
    const bool lt = (p < hi) && (__ldg(&d[p].x) < target);
    const uint32_t c = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, lt));
    const uint32_t nlo = lo + c * step;
    hi = min(hi, lo + (c + 1u) * step - 1u);
    lo = nlo;
  }
  const uint32_t p = lo + (uint32_t)lane;
  const bool lt = (p < hi) && (__ldg(&d[p].x) < target);
  return lo + (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, lt));
}

// Per-warp constants of the sub-range being swept
struct SubCtx {
  uint32_t sub_hi;      // first docid after the sub-range
  uint32_t sbase;       // slot address of docid d = sbase + d * slot size
  uint32_t slots_addr;
  uint32_t hot_addr;
  uint32_t cnt_addr;
  float thr;            // score of the k-th best key so far (FLT_MIN until k hits exist)
};

// a document whose running score crossed the threshold: remember its slot (rare once k hits exist)
__device__ __forceinline__ void hot_push(const SubCtx& cx, uint32_t a, uint32_t shift) {
  const uint32_t h = atoms_inc(cx.cnt_addr);
  if (h < (uint32_t)ST_HOT) sts_u16(cx.hot_addr + h * 2u, (a - cx.slots_addr) >> shift);
}

// ---- flat OR: 4-byte score slots ------------------------------------------------------------
__device__ __forceinline__ void or_one(const SubCtx& cx, float w, uint32_t d, uint32_t ubits, unsigned& tot) {
  const uint32_t a = cx.sbase + (d << 2);
  const float old = lds_f32(a);
  const float nw = fmaf(w, __uint_as_float(ubits), old);
  sts_f32(a, nw);
  tot += (old == 0.0f) ? 1u : 0u;                       // first hit of the slot: a match
  if (nw >= cx.thr && old < cx.thr) hot_push(cx, a, 2u);
}
// four postings of one list per lane (distinct documents): loads first, then stores
__device__ __forceinline__ void or_four(const SubCtx& cx, float w, const uint2 (&q)[4], unsigned& tot) {
  uint32_t a[4];
  float old[4], nw[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) a[e] = cx.sbase + (q[e].x << 2);
#pragma unroll
  for (int e = 0; e < 4; ++e) old[e] = lds_f32(a[e]);
#pragma unroll
  for (int e = 0; e < 4; ++e) nw[e] = fmaf(w, __uint_as_float(q[e].y), old[e]);
#pragma unroll
  for (int e = 0; e < 4; ++e) sts_f32(a[e], nw[e]);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    tot += (old[e] == 0.0f) ? 1u : 0u;
    if (nw[e] >= cx.thr && old[e] < cx.thr) hot_push(cx, a[e], 2u);
  }
}

// ---- AND of OR-groups: 8-byte {groups matched, score} slots ---------------------------------
// A leaf of a NOT clause (group NEG_GROUP; such queries have at most NEG_GROUP - 1 positive groups) poisons
// the slot: leaves are visited negatives first, and no posting of a positive group ever finds "groups
// matched" equal to its rank in a poisoned slot.
constexpr uint32_t NEG_GROUP = 31u;
constexpr uint32_t SLOT_POISON = 0xFFFFFFF0u;

template <bool NEG>                                     // NEG: the kernel serves NOT clauses
__device__ __forceinline__ void and_one(const SubCtx& cx, float w, uint32_t g, bool lastg, uint32_t d, uint32_t ubits, unsigned& tot) {
  const uint32_t a = cx.sbase + (d << 3);
  if (NEG && g == NEG_GROUP) { sts_v2(a, SLOT_POISON, 0u); return; }
  const uint2 v = lds_v2(a);
  if (v.x - g <= 1u) {                                  // alive: all earlier groups matched
    const float old = __uint_as_float(v.y);
    const float nw = fmaf(w, __uint_as_float(ubits), old);
    sts_v2(a, g + 1u, __float_as_uint(nw));
    if (lastg) {
      const bool fresh = (v.x == g);                    // this hit completes the match
      tot += fresh ? 1u : 0u;
      if (nw >= cx.thr && (fresh || old < cx.thr)) hot_push(cx, a, 3u);
    }
  }
}
template <bool NEG>
__device__ __forceinline__ void and_four(const SubCtx& cx, float w, uint32_t g, bool lastg, const uint2 (&q)[4], unsigned& tot) {
  uint32_t a[4];
  uint2 v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) a[e] = cx.sbase + (q[e].x << 3);
  if (NEG && g == NEG_GROUP) {
#pragma unroll
    for (int e = 0; e < 4; ++e) sts_v2(a[e], SLOT_POISON, 0u);
    return;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = lds_v2(a[e]);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (v[e].x - g <= 1u) {
      const float old = __uint_as_float(v[e].y);
      const float nw = fmaf(w, __uint_as_float(q[e].y), old);
      sts_v2(a[e], g + 1u, __float_as_uint(nw));
      if (lastg) {
        const bool fresh = (v[e].x == g);
        tot += fresh ? 1u : 0u;
        if (nw >= cx.thr && (fresh || old < cx.thr)) hot_push(cx, a[e], 3u);
      }
    }
  }
}

// Requires: k <= 32 * KR, <= ST_MAX_LEAVES leaves, every leaf weight > 0, no after_key, no postings of
// deleted documents in the store (bm25f_create compacts them away).
template <int KR, bool FINAL>
// (KR >= 4 without final(): a register cap of 88 - the bound of a 736-thread CTA - instead of the 92 / 96 the compiler
// would take; it settles at 80 without spills, and one k_score_isect<4 / 8> CTA, 72 / 80 registers x 256 threads, then
// fits on the SM beside this kernel's 512 threads: config 4, top-100, 147.4 -> 128.3 ms per 50k-query step)
__global__ void __launch_bounds__((KR >= 4 && !FINAL) ? 736 : ST_MAX_WARPS * 32, 1) k_score_stream(StreamParams sp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const uint32_t slot_bytes = sp.slot_bytes;
  // a device-planned batch is launched with a full grid: CTAs beyond the item count leave their SM to the other kernels
  if (sp.n_items_dev && blockIdx.x * (uint32_t)nwarps >= __ldg(sp.n_items_dev)) return;
  // shared memory: [nwarps][slot_bytes] slots | [nwarps][ST_MAX_LEAVES][32] tails (8 B) | [nwarps][ST_HOT] hot (2 B) | [nwarps] hot counters |
  // [nwarps][ST_MAX_LEAVES] leaf records (32 B, 16-byte aligned)
  SubCtx cx;
  const uint32_t smem0 = smem_u32(smem_raw);
  cx.slots_addr = smem0 + (uint32_t)warp * slot_bytes;
  const uint32_t tails_addr = smem0 + (uint32_t)nwarps * slot_bytes + (uint32_t)warp * (ST_MAX_LEAVES * 256u) + (uint32_t)lane * 8u;
  cx.hot_addr = smem0 + (uint32_t)nwarps * (slot_bytes + ST_MAX_LEAVES * 256u) + (uint32_t)warp * (ST_HOT * 2);
  cx.cnt_addr = smem0 + (uint32_t)nwarps * (slot_bytes + ST_MAX_LEAVES * 256u + ST_HOT * 2) + (uint32_t)warp * 4u;
  // leaf records of the warp's item: {base lo, base hi, cursor, end | weight, group}, 32 bytes per leaf
  const uint32_t state_addr = ((smem0 + (uint32_t)nwarps * (slot_bytes + ST_MAX_LEAVES * 256u + ST_HOT * 2 + 4u) + 15u) & ~15u) +
                              (uint32_t)warp * (ST_MAX_LEAVES * 32u);

  for (uint32_t o = (uint32_t)lane * 16u; o < slot_bytes; o += 512u) sts_zero16(cx.slots_addr + o);
  if (lane == 0) sts_u32(cx.cnt_addr, 0u);
  __syncwarp();

  for (;;) {
    uint32_t item_idx = 0;
    if (lane == 0) item_idx = atomicAdd(sp.queue, 1u);
    item_idx = __shfl_sync(0xFFFFFFFFu, item_idx, 0);
    if (item_idx >= (sp.n_items_dev ? __ldg(sp.n_items_dev) : sp.n_items)) break;

    const ItemRec item = sp.items[(sp.items_off_dev ? __ldg(sp.items_off_dev) : 0u) + item_idx];
    const QueryRec q = sp.queries[item.q];
    const int L = (int)q.n_leaves;
    const uint32_t G = q.n_groups;
    const bool simple_or = (q.flags & QF_SIMPLE_OR) != 0;
    const uint32_t shift = simple_or ? 2u : 3u;
    const uint32_t SW = slot_bytes >> shift;            // documents per sub-range
    const uint32_t d_lo = item.tile_begin, d_hi = item.tile_end;

    // ---- leaf state ------------------------------------------------------------------------------
    // Index space of a leaf: absolute posting index minus its base, where base = off & ~31, so every
    // row is 256-byte aligned.  The list occupies [off & 31, s_end).
    // The records live in shared memory (one broadcast load per visit instead of six shuffles); the docid at
    // the cursor of leaf l stays in lane l, so "which leaves have a posting in this sub-range" is one ballot.
    uint32_t s_next = 0xFFFFFFFFu;
    if (lane < L) {
      const LeafRec lf = sp.leaves[q.leaf_begin + lane];
      const uint32_t a = (uint32_t)(lf.off & 31ull);
      const unsigned long long base = lf.off - a;
      sts_v4(state_addr + (uint32_t)lane * 32u, (uint32_t)base, (uint32_t)(base >> 32), a, a + lf.df);
      sts_v2(state_addr + (uint32_t)lane * 32u + 16u, __float_as_uint(lf.w), lf.group);
    }
    __syncwarp();
    for (int l = 0; l < L; ++l) {
      const uint4 st4 = lds_v4(state_addr + (uint32_t)l * 32u);
      const unsigned long long base = (unsigned long long)st4.x | ((unsigned long long)st4.y << 32);
      const uint32_t end = st4.w;
      uint32_t cur = st4.z;
      const uint2* __restrict__ pairs = sp.pairs + base;
      if (d_lo > 0u) cur += warp_lower_bound(pairs + cur, end - cur, d_lo, lane);
      // park the row that holds the cursor; remember the docid at the cursor
      const uint32_t idx = (cur & ~31u) + (uint32_t)lane;
      uint2 t = make_uint2(0xFFFFFFFFu, 0u);
      if (idx < end) t = ldg_pair(pairs + idx);
      sts_v2(tails_addr + (uint32_t)l * 256u, t.x, t.y);
      const uint32_t nd = __shfl_sync(0xFFFFFFFFu, t.x, cur & 31u);
      if (lane == 0) sts_u32(state_addr + (uint32_t)l * 32u + 8u, cur);
      if (lane == l) s_next = (cur < end) ? nd : 0xFFFFFFFFu;
      if (sp.pf_dist) {
        // chunks [cur, cur + pf_dist + chunk), one per lane
        const uint32_t c0 = (cur & ~(ST_PF_CHUNK - 1u)) + (uint32_t)lane * ST_PF_CHUNK;
        if (c0 < end && c0 <= cur + sp.pf_dist) bulk_prefetch_l2(pairs + c0, min(ST_PF_CHUNK, (end - c0 + 1u) & ~1u) * 8u);
      }
    }

    __syncwarp();
    unsigned long long top[KR];               // lane i, row j: the (32 j + i)-th best key of this item so far
#pragma unroll
    for (int j = 0; j < KR; ++j) top[j] = 0ull;
    unsigned long long thr_key = 0ull;
    // FLT_MIN until k hits exist: every first hit is hot.  FINAL: final() can lift any match over the k-th
    // best, so nothing is ever hot and every sub-range with a match is scanned (below).
    cx.thr = FINAL ? __int_as_float(0x7f800000) : 1.17549435e-38f;
    unsigned int tot = 0;
    uint32_t topl[FINAL ? KR : 1];            // FINAL: low halves of the keys (thr_key / thr_lo: the k-th best)
#pragma unroll
    for (int j = 0; j < (FINAL ? KR : 1); ++j) topl[j] = 0u;
    uint32_t thr_lo = 0u;
    double thr_v = -INFINITY, add_min = -INFINITY;   // FINAL: the k-th best final value; the least date term that can beat it

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
      const unsigned int tot_sub = tot;

      while (todo) {
        const int l = __ffs(todo) - 1;          // ascending leaf order = ascending group rank
        todo &= todo - 1u;
        const uint4 st4 = lds_v4(state_addr + (uint32_t)l * 32u);
        const uint2 wg = lds_v2(state_addr + (uint32_t)l * 32u + 16u);
        const unsigned long long base = (unsigned long long)st4.x | ((unsigned long long)st4.y << 32);
        const uint32_t end = st4.w;
        uint32_t cur = st4.z;
        const float w = __uint_as_float(wg.x);
        const uint32_t g = wg.y;
        const bool lastg = (g + 1u == G);
        const uint2* __restrict__ pairs = sp.pairs + base;
        const uint32_t tail = tails_addr + (uint32_t)l * 256u;

        uint2 r = lds_v2(tail);                 // the row that holds the cursor
        uint32_t off = cur & 31u;
        bool dirty = false;                     // r differs from the parked tail
        for (;;) {
          // ---- one row: the lanes at or after the cursor whose docid is inside the sub-range form
          // a contiguous run (lists are sorted)
          const bool act = ((uint32_t)lane >= off) && (r.x < cx.sub_hi);
          const unsigned mk = __ballot_sync(0xFFFFFFFFu, act);
          if (act) {
            if (simple_or) or_one(cx, w, r.x, r.y, tot);
            else and_one<true>(cx, w, g, lastg, r.x, r.y, tot);
          }
          const uint32_t n = (uint32_t)__popc(mk);
          cur += n;
          if (n == 0u || (cur & 31u) != 0u || cur >= end) break;   // the sub-range (or the list) ends in this row
          dirty = true;
          off = 0u;
          // ---- the row is exhausted and the cursor is row-aligned: stream super-rows while whole
          // ones fit, each requested one super-row before it is processed
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
                // the sub-range ends inside this super-row: the row it ends in becomes the tail
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
            if (stop) break;                    // r holds the row of the cursor
            if (cur >= end) { r = make_uint2(0xFFFFFFFFu, 0u); break; }
          }
          // ---- next single row
          const uint32_t idx = cur + (uint32_t)lane;
          r = make_uint2(0xFFFFFFFFu, 0u);
          if (idx < end) r = ldg_pair(pairs + idx);
        }
        if (dirty) sts_v2(tail, r.x, r.y);
        const uint32_t nd = __shfl_sync(0xFFFFFFFFu, r.x, cur & 31u);
        if (lane == 0) sts_u32(state_addr + (uint32_t)l * 32u + 8u, cur);
        if (lane == l) s_next = (cur < end) ? nd : 0xFFFFFFFFu;
      }

      // ---- sub-range epilogue ------------------------------------------------------------------
      __syncwarp();
      if constexpr (FINAL) {
        if (__any_sync(0xFFFFFFFFu, tot != tot_sub)) {          // the sub-range has matches: final() of each
          // ... of each that can still beat the k-th best.  A dated document's value is below (1 + add) / 1e9,
          // an undated one's below 1: blocks of 32 documents whose largest date term is too small are skipped
          // with one coalesced load per 32 blocks, the others pay one gather per match.
          const uint32_t n = cx.sub_hi - sub_lo;
          const uint32_t n_it = (n + 31u) >> 5;
          for (uint32_t c0 = 0; c0 < n_it; c0 += 32u) {
            const uint32_t it = c0 + (uint32_t)lane;
            bool pass = false;
            if (it < n_it) {
              const uint32_t d0 = sub_lo + (it << 5);
              pass = (thr_v < 1.0) || (__ldg(sp.final_blk + (d0 >> 5)) >= add_min) ||
                     ((d0 & 31u) != 0u && __ldg(sp.final_blk + (d0 >> 5) + 1) >= add_min);
            }
            unsigned im = __ballot_sync(0xFFFFFFFFu, pass);
            while (im) {
              const uint32_t j = ((c0 + (uint32_t)(__ffs(im) - 1)) << 5) + (uint32_t)lane;
              im &= im - 1u;
              unsigned long long kh = 0ull;
              uint32_t kl = 0u;
              if (j < n) {
                float sc;
                bool ok;
                if (simple_or) {
                  sc = lds_f32(cx.slots_addr + (j << 2));
                  ok = sc != 0.0f;
                } else {
                  const uint2 v = lds_v2(cx.slots_addr + (j << 3));
                  sc = __uint_as_float(v.y);
                  ok = v.x == G;
                }
                if (ok) {
                  const double add = __ldg(sp.final_add + sub_lo + j);
                  if (isnan(add) ? (thr_v < 1.0) : (add >= add_min)) {
                    kh = orderable_f64(final_value(sc, add));
                    kl = 0xFFFFFFFFu - (sp.doc_base + sub_lo + j);
                  }
                }
              }
              unsigned pm = __ballot_sync(0xFFFFFFFFu, key2_wanted(kh, kl, thr_key, thr_lo, q.after_key, q.after_lo));
              while (pm) {
                const int src = __ffs(pm) - 1;
                pm &= pm - 1u;
                const unsigned long long bh = __shfl_sync(0xFFFFFFFFu, kh, src);
                const uint32_t bl = __shfl_sync(0xFFFFFFFFu, kl, src);
                if (key2_gt(bh, bl, thr_key, thr_lo)) {
                  warp_topk2_insert_rows<KR>(top, topl, bh, bl, lane);
                  warp_topk2_kth<KR>(top, topl, sp.k, thr_key, thr_lo);
                  if (thr_key != 0ull) {
                    thr_v = orderable_f64_value(thr_key);
                    add_min = thr_v * 1e9 - 2.0;
                  }
                }
              }
            }
          }
        }
      }
      if constexpr (!FINAL) {
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
    if constexpr (FINAL) {
      unsigned int* out_lo = sp.part_lo + (size_t)item.part * sp.k;
#pragma unroll
      for (int j = 0; j < KR; ++j)
        if (32 * j + lane < sp.k) out_lo[32 * j + lane] = topl[j];
    }
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xFFFFFFFFu, tot, o);
    if (lane == 0 && tot) atomicAdd(sp.totals + item.q, (unsigned long long)tot);
  }
}
