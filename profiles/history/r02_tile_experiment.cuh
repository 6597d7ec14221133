// k_score_tile — flat ORs, CTA-cooperative (included by bm25f.cu after stream.cuh).
//
// Replaces Whoosh's UnionMatcher loop + BM25FScorer + TopCollector for Or([Term, ...]) / single-term
// queries (reference call sites my_flask.py:208, :211, :304).
//
// Why it exists: the warp-private stream kernel spends 2.3 warp-instructions per posting, almost all of it
// per-(sub-range, leaf) bookkeeping, because a warp's accumulators hold only ~3k documents
// (profiles/r01_notes.md).  Here ONE CTA owns a query and the whole CTA sweeps it tile by tile:
//
//  * the accumulators of a TILE of `tile_docs` documents live in shared memory as 8-byte slots
//    {tag, score}; the tag is the CTA's running tile number, so a slot whose tag differs is "empty" and
//    nothing is ever cleared (a first touch is also exactly a new match: totals are counted for free);
//  * every leaf's posting list is STREAMED through its own FIFO ring in shared memory, in segments of
//    TL_SEG postings (1 KB), by a PRODUCER warp (lane l serves leaf l) with 1-D bulk copies
//    (cp.async.bulk + mbarrier complete_tx: the TMA engine, SASS UBLKCP).  The rings are filled as fast as
//    the consumers free segments, regardless of tile boundaries: the whole ring (`ring_postings` x 8 bytes
//    per CTA, shared out between the leaves in proportion to their postings) is the prefetch depth, DRAM
//    latency is off the consumers' critical path and costs them no registers;
//  * the CONSUMER warps follow a static SCHEDULE that the CTA builds per item from a boundary table
//    (k_tile_item_bounds): the VISITS (tile, leaf) in order, every visit cut at the ring's segment
//    boundaries into PIECES of up to four rows of 32 postings, and the pieces dealt round-robin to the
//    warps.  A piece is one conflict-free 64-bit shared load per posting, then LDS.64 slot / FFMA /
//    STS.64 slot.  Postings of one list are distinct documents, so no atomics; two leaves of a tile may
//    hold the same document, so a piece starts only when every piece of the earlier visits is done - a
//    shared count of finished pieces, no CTA barrier: a warp without a piece in a visit just walks on.
//    The summation order is fixed = leaf order;
//  * a document is looked at for the top-k only when its running score crosses the k-th best score so
//    far ("hot" list, as in the stream kernel); the k best 64-bit keys of the item live in shared memory
//    and are updated by the one warp that runs the tile's epilogue piece.
//
// A visit is ~(tile_docs / 2944) times longer than in the stream kernel, its fixed cost is paid by the
// warps that have a piece in it, and nobody waits at a barrier.
#pragma once

constexpr int TL_MAX_LEAVES = 32;
constexpr int TL_MAX_SLOTS = 64;           // 1 KB ring blocks per CTA
constexpr int TL_MAX_CWARPS = 31;          // consumer warps (+ 1 producer warp <= 1024 threads)
constexpr int TL_HOT = 128;                // hot-list entries per tile
constexpr uint32_t TL_VCAP = 512;          // schedule entries (visits + epilogues) of one item: (leaves + 2) per tile
constexpr uint32_t TL_KEYS = 512;          // key buffer of the first tile's cooperative scan
constexpr uint32_t TV_VISIT = 0u, TV_SCAN = 1u, TV_EPI = 2u;
constexpr uint32_t TL_SEG = 128;           // postings per ring block (1 KB) = per piece; a SEGMENT (one bulk copy, one
constexpr uint32_t TL_SEG_LOG = 7;         // barrier) is 1, 2 or 4 blocks, per leaf, by how fast the leaf is consumed

struct TileParams {
  const uint2* pairs;              // {docid, impact bits} per posting
  const LeafRec* leaves;
  const QueryRec* queries;
  const ItemRec* items;            // tile_begin / tile_end hold the item's document range [lo, hi)
  const uint32_t* item_boff;       // [n_items] first entry of the item's boundary table
  uint32_t* bounds;                // per item [(nt + 1)][L]: index inside leaf l's list of its first posting
                                   // with docid >= lo + j * tile_docs (row nt: >= hi)
  unsigned long long* part_keys;   // [n_parts * k]
  unsigned long long* totals;      // [Q]
  unsigned int* queue;             // work counter, zeroed before the launch
  uint32_t n_items;
  uint32_t tile_docs;              // documents per tile (even, <= 65536)
  uint32_t ring_slots;             // ring segments per CTA (<= TL_MAX_SLOTS); a query may have at most this many leaves
  uint32_t doc_base;
  int k;
};

// One warp per item: every entry of the item's boundary table is one binary search over a posting list.
__global__ void __launch_bounds__(128) k_tile_item_bounds(TileParams tp) {
  const int lane = threadIdx.x & 31;
  const uint32_t it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (it >= tp.n_items) return;
  const ItemRec item = tp.items[it];
  const QueryRec q = tp.queries[item.q];
  const uint32_t L = q.n_leaves;
  const uint32_t lo = item.tile_begin, hi = item.tile_end, T = tp.tile_docs;
  const uint32_t nt = (hi - lo + T - 1u) / T;
  uint32_t* __restrict__ out = tp.bounds + tp.item_boff[it];
  const uint32_t n = (nt + 1u) * L;
  for (uint32_t e = (uint32_t)lane; e < n; e += 32u) {
    const uint32_t j = e / L, l = e - j * L;
    const LeafRec lf = tp.leaves[q.leaf_begin + l];
    const uint32_t target = (j == nt) ? hi : lo + j * T;
    const uint2* __restrict__ d = tp.pairs + lf.off;
    uint32_t a = 0u, b = lf.df;
    if (target == 0u) b = 0u;
    while (a < b) {
      const uint32_t mid = (a + b) >> 1;
      if (__ldg(&d[mid].x) < target) a = mid + 1u; else b = mid;
    }
    out[e] = a;
  }
}

// A leaf of the item being swept (shared memory, 48 bytes)
struct __align__(16) TileLeaf {
  unsigned long long base;   // absolute posting index of ring position 0 (even: bulk copies are 16-byte aligned)
  float w;                   // leaf weight
  uint32_t shift;            // ring-relative index of list element x is x - shift
  uint32_t ring_addr;        // shared-memory address of the leaf's ring
  uint32_t cap_mask;         // ring capacity in postings - 1 (capacity = segments * TL_SEG, a power of two)
  uint32_t slot0_lg;         // first block (= first barrier) of the ring | log2(segments) << 8 | log2(blocks per segment) << 16 |
                             // (ring-relative index of the leaf's first posting: 0 or 1) << 24
  uint32_t rel_end;          // ring-relative end of the leaf's postings in this item
  uint32_t par_lo, par_hi;   // bit i: parity of the phases the "full" barrier of the ring's block i had completed when the item began
  uint32_t seen;             // 1 + the segment of the leaf a consumer last saw landed (another piece of it need not wait again)
  uint32_t pad1;
};

__host__ __device__ inline size_t tile_smem_bytes(uint32_t tile_docs, uint32_t ring_slots, uint32_t cwarps) {
  // slots | rings | key buffer of the first tile's scan | schedule
  (void)cwarps;
  return (size_t)tile_docs * 8 + (size_t)ring_slots * TL_SEG * 8 + (size_t)TL_KEYS * 8 + (size_t)TL_VCAP * 16;
}

__device__ __forceinline__ void tl_cbar(uint32_t nct) {
  asm volatile("bar.sync 1, %0;" ::"r"(nct) : "memory");
}
__device__ __forceinline__ uint32_t lds_acquire_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// one more piece of the item's schedule is finished (its shared-memory writes are visible to whoever sees the count)
__device__ __forceinline__ void tl_piece_done(uint32_t done_addr, int lane) {
  __syncwarp();
  if (lane == 0) asm volatile("red.release.cta.shared.add.u32 [%0], 1;" ::"r"(done_addr) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity)
        : "memory");
  } while (!ok);
}

// Tile epilogue, run by the warp that finished the tile's last piece: the documents on the hot list (or, if it
// overflowed, the lists of the first tile's scan pieces, or every slot of the tile) are offered to the item's k
// best keys, which live in shared memory; the threshold for the next tile is published.
template <int KR>
__device__ __forceinline__ void tl_epilogue(uint32_t acc_addr, uint32_t hot_addr, uint32_t nhot_addr, uint32_t top_addr,
                                            uint32_t thr_addr, uint32_t keybuf_addr, unsigned long long* thrkey, uint32_t tag,
                                            uint32_t gdoc0, uint32_t Tn, uint32_t n_scanned, int k, int lane) {
  const uint32_t nhot = lds_u32(nhot_addr);
  if (nhot == 0u) return;
  unsigned long long top[KR];
#pragma unroll
  for (int r = 0; r < KR; ++r) {
    const uint2 kv = lds_v2(top_addr + ((32u * r + (uint32_t)lane) << 3));
    top[r] = ((unsigned long long)kv.y << 32) | kv.x;
  }
  unsigned long long thr_key = *thrkey;
  const float thr = lds_f32(thr_addr);
  const bool listed = nhot <= (uint32_t)TL_HOT;
  const bool scanned = !listed && n_scanned != 0u;
  const uint32_t n = listed ? nhot : scanned ? n_scanned : Tn;
  for (uint32_t j0 = 0; j0 < n; j0 += 32u) {
    const uint32_t jj = j0 + (uint32_t)lane;
    unsigned long long key = 0ull;
    if (jj < n) {
      if (scanned) {
        const uint2 kv = lds_v2(keybuf_addr + (jj << 3));
        key = ((unsigned long long)kv.y << 32) | kv.x;
      } else {
        const uint32_t slot = listed ? lds_u16(hot_addr + jj * 2u) : jj;
        const uint2 v = lds_v2(acc_addr + (slot << 3));
        if (v.x == tag && __uint_as_float(v.y) >= thr) key = make_key(__uint_as_float(v.y), gdoc0 + slot);
      }
    }
    unsigned pm = __ballot_sync(0xFFFFFFFFu, key > thr_key);
    while (pm) {
      const int src = __ffs(pm) - 1;
      pm &= pm - 1u;
      const unsigned long long bk = __shfl_sync(0xFFFFFFFFu, key, src);
      if (bk > thr_key) {
        warp_topk_insert_rows<KR>(top, bk, lane);
        thr_key = warp_topk_kth<KR>(top, k);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < KR; ++r) sts_v2(top_addr + ((32u * r + (uint32_t)lane) << 3), (uint32_t)top[r], (uint32_t)(top[r] >> 32));
  if (lane == 0) {
    *thrkey = thr_key;
    if (thr_key != 0ull) sts_f32(thr_addr, key_score(thr_key));
    sts_u32(nhot_addr, 0u);
  }
}

// Requires: flat OR (one group, no NOT clause), every leaf weight > 0, k <= 32 * KR, <= min(TL_MAX_LEAVES,
// ring_slots) leaves, no after_key, tiles of the item * (leaves + 2) <= TL_VCAP, no postings of deleted
// documents in the store.
template <int KR>
__global__ void __launch_bounds__(1024, 1) k_score_tile(TileParams tp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long s_full[TL_MAX_SLOTS];
  __shared__ __align__(8) unsigned long long s_wbar[32];   // per consumer warp: "the schedule entry of your next piece is open"
  __shared__ uint32_t s_free[TL_MAX_SLOTS];      // per segment (indexed by its first block): times it was handed back this item
  __shared__ uint32_t s_used[TL_MAX_SLOTS];      // ... postings of its current content that the consumers have finished
  __shared__ TileLeaf s_leaf[TL_MAX_LEAVES];
  __shared__ uint32_t s_par[2];                  // per ring slot: parity of the phases its two barriers have completed so far
  __shared__ unsigned short s_hot[TL_HOT];
  __shared__ unsigned long long s_top[32 * KR];   // the item's k best keys so far, descending
  __shared__ unsigned long long s_thrkey;
  __shared__ uint32_t s_nhot;
  __shared__ uint32_t s_done;               // finished pieces of the item's schedule
  __shared__ uint32_t s_wsum[2 * 32];       // schedule build: per-warp sums of the two scans
  __shared__ uint32_t s_nvis;
  __shared__ float s_thr;
  __shared__ uint32_t s_item;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int NC = (int)(blockDim.x >> 5) - 1;            // consumer warps; the last warp produces
  const uint32_t NCT = (uint32_t)NC * 32u;
  const uint32_t T = tp.tile_docs, NSLOT = tp.ring_slots;

  const uint32_t acc_addr = smem_u32(smem_raw);
  const uint32_t ring_addr = acc_addr + T * 8u;
  const uint32_t keybuf_addr = ring_addr + NSLOT * TL_SEG * 8u;
  const uint32_t vdesc_addr = keybuf_addr + TL_KEYS * 8u;
  const uint32_t leaf_addr = smem_u32(s_leaf);
  const uint32_t full_addr = smem_u32(s_full), free_addr = smem_u32(s_free), used_addr = smem_u32(s_used);
  const uint32_t hot_addr = smem_u32(s_hot);
  const uint32_t thr_addr = smem_u32(&s_thr);
  const uint32_t nhot_addr = smem_u32(&s_nhot);
  const uint32_t done_addr = smem_u32(&s_done);
  const uint32_t top_addr = smem_u32(s_top);
  const uint32_t wbar_addr = smem_u32(s_wbar);

  for (uint32_t o = (uint32_t)tid * 16u; o < T * 8u; o += blockDim.x * 16u) sts_zero16(acc_addr + o);
  if (tid == 0) {
    s_nhot = 0u;
    s_par[0] = 0u; s_par[1] = 0u;
    s_item = atomicAdd(tp.queue, 1u);
  }
  // The "full" barriers (a segment's bulk copy has landed) are set up once and run on from item to item: one
  // parity bit per barrier (s_par) carries their state over.  (Re-initialising them per item with
  // mbarrier.inval + init left stale phases behind on this driver.)  Handing a segment back to the producer is
  // a plain shared-memory count: mbarrier.test_wait costs ~150 cycles a probe, a shared load 30.
  if ((uint32_t)tid < NSLOT) mbar_init(&s_full[tid], 1);
  if (tid < 32) mbar_init(&s_wbar[tid], 1);
  if (tid < 64) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  uint32_t wpar = 0u;                       // consumers: phase parity of my s_wbar
  uint32_t tag_base = 0u;                   // consumers: tiles swept by this CTA so far (tile j of an item has tag tag_base + j + 1; slots start at 0)

  for (;;) {
    const uint32_t item_idx = s_item;
    if (item_idx >= tp.n_items) break;
    const ItemRec item = tp.items[item_idx];
    const QueryRec q = tp.queries[item.q];
    const int L = (int)q.n_leaves;
    const uint32_t d_lo = item.tile_begin, d_hi = item.tile_end;
    const uint32_t nt = (d_hi - d_lo + T - 1u) / T;
    const uint32_t* __restrict__ gb = tp.bounds + tp.item_boff[item_idx];
    if (warp == NC) {
      // ---- leaf records and ring shares (lane l = leaf l) ----------------------------------------
      uint32_t n_l = 0u, nseg_total = 0u, shift = 0u, nseg = 0u, rel_end = 0u, bbeg_odd = 0u;
      unsigned long long base = 0ull;
      float w = 0.0f;
      if (lane < L) {
        const LeafRec lf = tp.leaves[q.leaf_begin + lane];
        const uint32_t bbeg = __ldg(gb + lane), bend = __ldg(gb + (size_t)nt * L + lane);
        const uint32_t odd = (uint32_t)((lf.off + bbeg) & 1ull);
        n_l = bend - bbeg;
        bbeg_odd = odd;
        shift = bbeg - odd;                                   // may wrap below zero (bbeg = 0, odd = 1): arithmetic is mod 2^32
        base = lf.off + bbeg - odd;
        w = lf.w;
        rel_end = bend - shift;                               // ring-relative end of the leaf's postings in this item
      }
      // Segment size by how many postings a tile takes from the leaf (a dense leaf is copied 4 KB at a time: the
      // producer's cost is per copy); every leaf with postings gets one segment, shrunk while they do not all fit;
      // then the leaf with the most postings per ring block doubles its ring while blocks are left (rings are
      // powers of two so that positions wrap with a mask).
      uint32_t sblk_log = 0u;                                 // log2(blocks per segment)
      if (n_l) {
        const uint32_t per_tile = n_l / max(1u, nt);
        sblk_log = per_tile >= 1024u ? 2u : per_tile >= 256u ? 1u : 0u;
      }
      // two segments to start with (one being consumed, one in flight), halved while they do not all fit
      uint32_t nblk = n_l ? (2u << sblk_log) : 0u;
      for (;;) {
        uint32_t used = nblk;
        for (int o = 16; o > 0; o >>= 1) used += __shfl_xor_sync(0xFFFFFFFFu, used, o);
        if (used <= NSLOT) break;
        const uint32_t big = __reduce_max_sync(0xFFFFFFFFu, nblk);       // > 1: there are at most NSLOT leaves
        const int who = __ffs(__ballot_sync(0xFFFFFFFFu, nblk == big)) - 1;
        if (lane == who) { nblk >>= 1; if ((nblk >> sblk_log) == 0u) --sblk_log; }
      }
      nseg_total = n_l ? (rel_end + (TL_SEG << sblk_log) - 1u) >> (TL_SEG_LOG + sblk_log) : 0u;
      {
        uint32_t used = nblk;
        for (int o = 16; o > 0; o >>= 1) used += __shfl_xor_sync(0xFFFFFFFFu, used, o);
        uint32_t avail = NSLOT - used;
        for (;;) {
          const bool can = nblk != 0u && (nblk >> sblk_log) < nseg_total && nblk <= avail;
          const uint32_t score = can ? max(1u, n_l / nblk) : 0u;
          const uint32_t best = __reduce_max_sync(0xFFFFFFFFu, score);
          if (best == 0u) break;
          const int who = __ffs(__ballot_sync(0xFFFFFFFFu, score == best)) - 1;
          const uint32_t add = __shfl_sync(0xFFFFFFFFu, nblk, who);
          if (lane == who) nblk <<= 1;
          avail -= add;
        }
      }
      nseg = nblk >> sblk_log;
      uint32_t slot0 = nblk;                                  // exclusive prefix sum over the lanes
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, slot0, o);
        if (lane >= o) slot0 += v;
      }
      slot0 -= nblk;
      const uint32_t lg = nseg ? (uint32_t)(31 - __clz(nseg)) : 0u;
      // parity bits of this lane's barriers (blocks slot0 .. slot0 + nblk - 1 < 64), shifted so that bit i is ring block i
      const unsigned long long par_all = (unsigned long long)s_par[0] | ((unsigned long long)s_par[1] << 32);
      const unsigned long long my_par = par_all >> slot0;
      if (lane < L) {
        TileLeaf tl;
        tl.base = base;
        tl.w = w;
        tl.shift = shift;
        tl.ring_addr = ring_addr + slot0 * TL_SEG * 8u;
        tl.cap_mask = nblk * TL_SEG - 1u;
        tl.slot0_lg = slot0 | (lg << 8) | (sblk_log << 16) | ((bbeg_odd & 1u) << 24);
        tl.rel_end = rel_end;
        tl.par_lo = (uint32_t)my_par;
        tl.par_hi = (uint32_t)(my_par >> 32);
        tl.seen = 0u; tl.pad1 = 0u;
        s_leaf[lane] = tl;
      }
      for (uint32_t i = 0; i < nblk; ++i) { s_free[slot0 + i] = 0u; s_used[slot0 + i] = 0u; }
      __syncthreads();                      // records, table and barriers are ready; everybody has read s_item

      // =============================== PRODUCER ===============================================
      // Lane l streams leaf l: segment k goes to ring position k mod nseg as soon as the consumer that finished
      // that position's previous segment has handed it back.  Nothing blocks (test_wait, not try_wait): the lanes
      // poll side by side, so no lane can starve another; a lane issues every segment it can per round.
      uint32_t k = 0u;
      const uint32_t my_ring = ring_addr + slot0 * TL_SEG * 8u;
      const uint32_t seg_log = TL_SEG_LOG + sblk_log;       // log2(postings per segment)
      uint32_t idle = 0u;
      for (;;) {
        bool progressed = false;
        while (k < nseg_total) {
          const uint32_t sp = k & (nseg - 1u);              // ring segment
          const uint32_t bidx = sp << sblk_log;             // its first block = its barrier
          if (lds_acquire_u32(free_addr + (slot0 + bidx) * 4u) < (k >> lg)) break;     // its previous content is not finished yet
          // the last segment stops at the (even-rounded) end of the range
          const uint32_t seg_beg = k << seg_log;
          const uint32_t n = min(1u << seg_log, (rel_end - seg_beg + 1u) & ~1u);
          const uint32_t bytes = n * 8u;
          const uint32_t fb = full_addr + (slot0 + bidx) * 8u;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           my_ring + (bidx << (TL_SEG_LOG + 3u))),
                       "l"(tp.pairs + base + seg_beg), "r"(bytes), "r"(fb)
                       : "memory");
          ++k;
          progressed = true;
        }
        if (__all_sync(0xFFFFFFFFu, k >= nseg_total)) break;
        if (__any_sync(0xFFFFFFFFu, progressed)) idle = 0u;
        else if (++idle > 2u) __nanosleep(64);
#ifdef BM25F_TILE_CHECK
        if (idle > 3000000u) {
          if (k < nseg_total) printf("[tile stuck] producer item %u leaf %d/%d: issued %u of %u segments, ring %u slots at %u\n",
                                     item_idx, lane, L, k, nseg_total, nseg, slot0);
          break;
        }
#endif
      }
      {
        // "full" phases completed by this item: ring segment sp of a leaf held the segments sp, sp + nseg, ...
        unsigned long long flips = 0ull;
        for (uint32_t sp = 0; sp < nseg && sp < nseg_total; ++sp) {
          const uint32_t uses = (nseg_total - sp + nseg - 1u) >> lg;
          if (uses & 1u) flips |= 1ull << (slot0 + (sp << sblk_log));
        }
        uint32_t f0 = (uint32_t)flips, f1 = (uint32_t)(flips >> 32);
        for (int o = 16; o > 0; o >>= 1) {
          f0 ^= __shfl_xor_sync(0xFFFFFFFFu, f0, o);
          f1 ^= __shfl_xor_sync(0xFFFFFFFFu, f1, o);
        }
        if (lane == 0) { s_par[0] ^= f0; s_par[1] ^= f1; }
      }
      if (lane == 0) s_item = atomicAdd(tp.queue, 1u);   // the next item (everybody read the current one before the barrier above)
    } else {
      // ---- the item's schedule: visits (tile, leaf) with postings, after tile 0 the pieces of its cooperative
      // scan, after every tile with postings its epilogue; in order, compacted, with the number of pieces before
      // each entry.  One thread per (tile, schedule slot).
      {
        const uint32_t VS = (uint32_t)L + 2u;
        const uint32_t nslots = nt * VS;
        const uint32_t S_scan = min((uint32_t)NC, TL_KEYS / (uint32_t)tp.k);
        uint32_t carry_u = 0u, carry_v = 0u;              // pieces / entries before the chunk (uniform)
        for (uint32_t c0 = 0; c0 < nslots; c0 += NCT) {
          const uint32_t t = c0 + (uint32_t)tid;
          uint32_t np = 0u, ra = 0u, rb = 0u, meta = 0u;
          if (t < nslots) {
            const uint32_t j = t / VS, l = t - j * VS;
            if (l < (uint32_t)L) {
              const uint32_t s0 = __ldg(gb + (size_t)j * L + l), e0 = __ldg(gb + (size_t)(j + 1u) * L + l);
              if (s0 < e0) {
                const uint32_t bbeg = __ldg(gb + l);
                const uint32_t shift = bbeg - (uint32_t)((tp.leaves[q.leaf_begin + l].off + bbeg) & 1ull);
                ra = s0 - shift;
                rb = e0 - shift;
                np = ((rb - 1u) >> TL_SEG_LOG) - (ra >> TL_SEG_LOG) + 1u;
              }
              // the tile's last entry runs the epilogue: a visit is last if no later leaf has postings here and
              // the tile has no scan entry
              bool later = (j == 0u);
              for (int ll = (int)l + 1; ll < L && !later; ++ll) later = __ldg(gb + (size_t)j * L + ll) < __ldg(gb + (size_t)(j + 1u) * L + ll);
              meta = l | (TV_VISIT << 6) | (j << 8) | ((np && !later) ? (1u << 29) : 0u);
            } else {
              bool any = false;
              for (int ll = 0; ll < L; ++ll) any = any || (__ldg(gb + (size_t)j * L + ll) < __ldg(gb + (size_t)(j + 1u) * L + ll));
              if (l == (uint32_t)L) {                       // cooperative scan of the item's first tile
                np = (any && j == 0u) ? S_scan : 0u;
                meta = l | (TV_SCAN << 6) | (j << 8) | (1u << 29);
                rb = np;
              } else {
                np = any ? 1u : 0u;                         // the epilogue: one more piece, run by the tile's last finisher
                meta = 0xFFFFFFFFu;
              }
            }
          }
          // two exclusive scans over the chunk: pieces and non-empty entries
          uint32_t iu = np, iv = (np && meta != 0xFFFFFFFFu) ? 1u : 0u;
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, iu, o), b = __shfl_up_sync(0xFFFFFFFFu, iv, o);
            if (lane >= o) { iu += a; iv += b; }
          }
          if (lane == 31) { s_wsum[warp] = iu; s_wsum[32 + warp] = iv; }
          tl_cbar(NCT);
          uint32_t pu = carry_u, pv = carry_v, tu = 0u, tv = 0u;
          for (int ww = 0; ww < NC; ++ww) {
            const uint32_t a = s_wsum[ww], b = s_wsum[32 + ww];
            if (ww < warp) { pu += a; pv += b; }
            tu += a; tv += b;
          }
          if (np && meta != 0xFFFFFFFFu) {
            const uint32_t cum = pu + iu - np, vi = pv + iv - 1u;
            sts_v4(vdesc_addr + vi * 16u, ra, rb, meta | ((cum % (uint32_t)NC) << 24), cum);
          }
          carry_u += tu;
          carry_v += tv;
          tl_cbar(NCT);                                   // s_wsum is reused by the next chunk
        }
        if (tid == 0) {
          s_nvis = carry_v;
          s_done = 0u;
          s_nhot = 0u;
          s_thr = 1.17549435e-38f;          // FLT_MIN until k hits exist: every first hit is hot
          s_thrkey = 0ull;
        }
        for (uint32_t i = (uint32_t)tid; i < 32u * KR; i += NCT) s_top[i] = 0ull;
      }
      __syncthreads();                      // schedule, leaf records and barriers are ready; everybody has read s_item

      // ================================= CONSUMERS ============================================
      const uint32_t nvis = s_nvis;
      const uint32_t S_scan = min((uint32_t)NC, TL_KEYS / (uint32_t)tp.k);
      unsigned int tot = 0;
      for (uint32_t vi = 0; vi < nvis; ++vi) {
        const uint4 D = lds_v4(vdesc_addr + vi * 16u);      // ra, rb, meta, pieces before this entry
        const uint32_t kind = (D.z >> 6) & 3u;
        const uint32_t np = kind == TV_VISIT ? ((D.y - 1u) >> TL_SEG_LOG) - (D.x >> TL_SEG_LOG) + 1u : D.y;
        // my first piece here: pieces are dealt to the warps in turn over the whole schedule
        uint32_t pc = (uint32_t)warp - ((D.z >> 24) & 31u);
        if ((int)pc < 0) pc += (uint32_t)NC;
        if (pc >= np) continue;
        const uint32_t j = (D.z >> 8) & 0xFFFFu;
        const uint32_t t0 = d_lo + j * T;
        const uint32_t tag = tag_base + j + 1u;
        const bool tile_last = (D.z >> 29) & 1u;              // the tile's last entry: whoever finishes its last piece runs the epilogue
        // An entry is OPEN when every piece of the earlier entries is finished (another leaf may hold the same
        // documents).  The warp that finishes an entry's last piece opens the next one by arriving on the barrier of
        // every warp that has a piece in it; the waiting warps sleep in hardware instead of polling a counter.
        bool opened = (vi == 0u);
        bool closed_it = false;
        uint4 Dn = make_uint4(0u, 0u, 0u, 0u);                // the next entry
        if (vi + 1u < nvis) Dn = lds_v4(vdesc_addr + (vi + 1u) * 16u);
        if (kind == TV_VISIT) {
          const uint32_t l = D.z & 63u;
          const uint4 la = lds_v4(leaf_addr + l * 48u);            // base lo, base hi, w, shift
          const uint4 lb = lds_v4(leaf_addr + l * 48u + 16u);      // ring_addr, cap_mask, slot0_lg, rel_end
          const uint2 lp = lds_v2(leaf_addr + l * 48u + 32u);      // parities of the ring's barriers at the item's start
          const uint32_t seen_addr = leaf_addr + l * 48u + 40u;    // 1 + the segment last seen landed
          const float w = __uint_as_float(la.z);
          const uint32_t ra = D.x, rb = D.y;                        // ring-relative posting range of the visit
          const uint32_t lring = lb.x, cmask = lb.y;
          const uint32_t lg = (lb.z >> 8) & 255u, nsegm1 = (1u << lg) - 1u;   // ring segments
          const uint32_t sblk_log = (lb.z >> 16) & 255u;                      // blocks per segment
          const uint32_t rel_begin = lb.z >> 24;
          const uint32_t slot0 = lb.z & 255u;
          const unsigned long long par0 = (unsigned long long)lp.x | ((unsigned long long)lp.y << 32);
          const uint32_t sbase = acc_addr - (t0 << 3);
          const uint32_t k_first = ra >> TL_SEG_LOG;
          for (; pc < np; pc += (uint32_t)NC) {
            // ---- before my turn: everything that does not touch the tile's slots.  The chain of pieces is what
            // bounds the kernel (a piece may start only when every piece of the earlier visits is finished), so
            // what a piece does inside its turn is kept to: load slots, add, store, count.
            const uint32_t kb = k_first + pc;                 // ring-relative block of the piece
            const uint32_t sk = kb >> sblk_log;               // its segment
            const uint32_t bidx = (sk & nsegm1) << sblk_log;  // the segment's first block in the ring = its barrier
            if (lds_acquire_u32(seen_addr) != sk + 1u) {      // a barrier wait costs ~90 cycles even when the copy has landed; copies
                                                              // may land out of order, so only this very segment counts
              // A phase parity only tells two consecutive uses of a ring position apart: first make sure the position's
              // previous segments have all been handed back (then this segment's copy is the one in flight or landed).
              const uint32_t turn = sk >> lg;
              while (lds_acquire_u32(free_addr + (slot0 + bidx) * 4u) < turn) __nanosleep(100);
              mbar_wait_addr(full_addr + (slot0 + bidx) * 8u, (turn & 1u) ^ (uint32_t)((par0 >> bidx) & 1ull));
              if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(seen_addr), "r"(sk + 1u) : "memory");
            }
            const uint32_t i0 = (kb << TL_SEG_LOG) + (uint32_t)lane;
            uint2 p[4], v[4];
            uint32_t a[4];
            bool ok[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              const uint32_t i = i0 + 32u * e2;
              ok[e2] = i >= ra && i < rb;
              p[e2] = make_uint2(0u, 0u);
              if (ok[e2]) p[e2] = lds_v2(lring + ((i & cmask) << 3));
            }
#ifdef BM25F_TILE_CHECK
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              if (ok[e2] && (p[e2].x - t0) >= min(T, d_hi - t0)) {
                printf("[tile check] item %u q %u leaf %u/%d tile %u t0 %u: docid %u outside; sk %u i %u ra %u rb %u lg %u slot0 %u cmask %x shift %u warp %d lane %d\n",
                       item_idx, item.q, l, L, j, t0, p[e2].x, sk, i0 + 32u * e2, ra, rb, lg, slot0, cmask, la.w, warp, lane);
                ok[e2] = false;
              }
            }
#endif
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) a[e2] = sbase + (p[e2].x << 3);
            // ---- my turn: the entry is open
            if (!opened) {
              mbar_wait_addr(wbar_addr + (uint32_t)warp * 8u, wpar);
              wpar ^= 1u;
              opened = true;
            }
            const float thr = lds_f32(thr_addr);              // one threshold per tile: it only changes in the epilogue
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              v[e2] = make_uint2(0u, 0u);
              if (ok[e2]) v[e2] = lds_v2(a[e2]);
            }
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              if (ok[e2]) {
                const bool fresh = v[e2].x != tag;
                const float old = fresh ? 0.0f : __uint_as_float(v[e2].y);
                const float nw = fmaf(w, __uint_as_float(p[e2].y), old);
                sts_v2(a[e2], tag, __float_as_uint(nw));
                tot += fresh ? 1u : 0u;
                if (nw >= thr && old < thr) {
                  const uint32_t h = atoms_inc(nhot_addr);
                  if (h < (uint32_t)TL_HOT) sts_u16(hot_addr + h * 2u, (a[e2] - acc_addr) >> 3);
                }
              }
            }
            __syncwarp();
            uint32_t before = 0u;
            if (lane == 0) asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(before) : "r"(done_addr) : "memory");
            // ---- after my turn: count the piece's postings on its segment; whoever completes the segment hands it
            // back to the producer
            if (lane == 0) {
              const uint32_t mine = min(rb, (kb + 1u) << TL_SEG_LOG) - max(ra, kb << TL_SEG_LOG);
              const uint32_t seg_n = min(lb.w, (sk + 1u) << (TL_SEG_LOG + sblk_log)) - max(rel_begin, sk << (TL_SEG_LOG + sblk_log));
              const uint32_t ua = used_addr + (slot0 + bidx) * 4u;
              uint32_t old;
              asm volatile("atom.relaxed.cta.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(ua), "r"(mine) : "memory");
              if (old + mine == seg_n) {
                sts_u32(ua, 0u);
                asm volatile("red.release.cta.shared.add.u32 [%0], 1;" ::"r"(free_addr + (slot0 + bidx) * 4u) : "memory");
              }
            }
            if (__shfl_sync(0xFFFFFFFFu, before, 0) + 1u == D.w + np) closed_it = true;
          }
        } else {
          // ---- scan piece of the item's first tile: the tile's slots are final once every visit piece is finished.
          // With no threshold yet and more hot documents than the list holds (the usual first tile), every scan
          // piece takes a share of the tile's slots and leaves its k best keys in the key buffer for the epilogue.
          if (!opened) {
            mbar_wait_addr(wbar_addr + (uint32_t)warp * 8u, wpar);
            wpar ^= 1u;
          }
          const uint32_t Tn = min(T, d_hi - t0);
          if (lds_u32(nhot_addr) > (uint32_t)TL_HOT) {
            unsigned long long top[KR];
#pragma unroll
            for (int r = 0; r < KR; ++r) top[r] = 0ull;
            unsigned long long tk = 0ull;
            const uint32_t share = (Tn + S_scan - 1u) / S_scan;
            const uint32_t s_end = min(Tn, (pc + 1u) * share);
            for (uint32_t s0 = pc * share; s0 < s_end; s0 += 32u) {
              const uint32_t sl = s0 + (uint32_t)lane;
              unsigned long long key = 0ull;
              if (sl < s_end) {
                const uint2 v = lds_v2(acc_addr + (sl << 3));
                if (v.x == tag) key = make_key(__uint_as_float(v.y), tp.doc_base + t0 + sl);
              }
              unsigned pm = __ballot_sync(0xFFFFFFFFu, key > tk);
              while (pm) {
                const int src = __ffs(pm) - 1;
                pm &= pm - 1u;
                const unsigned long long bk = __shfl_sync(0xFFFFFFFFu, key, src);
                if (bk > tk) {
                  warp_topk_insert_rows<KR>(top, bk, lane);
                  tk = warp_topk_kth<KR>(top, tp.k);
                }
              }
            }
#pragma unroll
            for (int r = 0; r < KR; ++r)
              if (32 * r + lane < tp.k) {
                const uint32_t at = keybuf_addr + ((pc * (uint32_t)tp.k + 32u * r + (uint32_t)lane) << 3);
                sts_v2(at, (uint32_t)top[r], (uint32_t)(top[r] >> 32));
              }
          }
          __syncwarp();
          uint32_t before = 0u;
          if (lane == 0) asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(before) : "r"(done_addr) : "memory");
          if (__shfl_sync(0xFFFFFFFFu, before, 0) + 1u == D.w + np) closed_it = true;
        }
        if (closed_it) {
          if (tile_last) {
            // every piece of the tile is finished and nobody starts the next tile before it is opened below
            tl_epilogue<KR>(acc_addr, hot_addr, nhot_addr, top_addr, thr_addr, keybuf_addr, &s_thrkey, tag, tp.doc_base + t0,
                            min(T, d_hi - t0), j == 0u ? S_scan * (uint32_t)tp.k : 0u, tp.k, lane);
            tl_piece_done(done_addr, lane);                   // the epilogue counts as a piece (the schedule's counts include it)
          }
          if (vi + 1u < nvis) {
            // open the next entry: its first min(pieces, NC) pieces go to that many different warps
            const uint32_t nkind = (Dn.z >> 6) & 3u;
            const uint32_t nnp = nkind == TV_VISIT ? ((Dn.y - 1u) >> TL_SEG_LOG) - (Dn.x >> TL_SEG_LOG) + 1u : Dn.y;
            if ((uint32_t)lane < min(nnp, (uint32_t)NC)) {
              uint32_t wn = ((Dn.z >> 24) & 31u) + (uint32_t)lane;
              if (wn >= (uint32_t)NC) wn -= (uint32_t)NC;
              asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(wbar_addr + wn * 8u) : "memory");
            }
          }
        }
      }
      tag_base += nt;
      // ---- item epilogue ------------------------------------------------------------------------
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xFFFFFFFFu, tot, o);
      if (lane == 0 && tot) atomicAdd(tp.totals + item.q, (unsigned long long)tot);
      tl_cbar(NCT);                         // every piece is finished: the top list is final
      {
        unsigned long long* out = tp.part_keys + (size_t)item.part * tp.k;
        for (uint32_t i = (uint32_t)tid; i < (uint32_t)tp.k; i += NCT) out[i] = s_top[i];
      }
    }
    __syncthreads();                        // the item is done; s_item holds the next one
  }
}
