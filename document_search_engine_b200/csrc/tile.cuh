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
//  * a PRODUCER warp walks the item's (tile, leaf) visits, whose posting ranges come from a boundary
//    table computed by k_tile_item_bounds, and stages the {docid, impact} pairs into a ring of
//    shared-memory stages with 1-D bulk copies (cp.async.bulk + mbarrier complete_tx: the TMA engine,
//    SASS UBLKCP), running up to `stages` chunks ahead of the consumers: DRAM latency is off the
//    consumers' critical path and costs them no registers;
//  * the CONSUMER warps take the rows (32 postings) of a staged chunk round-robin: one conflict-free
//    64-bit shared load per posting, then LDS.64 slot / FFMA / STS.64 slot.  Postings of one list are
//    distinct documents, so no atomics; a named barrier separates the leaves of a tile (the same
//    document may be in both), which also keeps the summation order fixed = leaf order: deterministic
//    and identical to the other kernels;
//  * a document is looked at for the top-k only when its running score crosses the k-th best score so
//    far ("hot" list, as in the stream kernel); warp 0 keeps the k best 64-bit keys in registers.
//
// A visit is ~(tile_docs / 2944) times longer than in the stream kernel and its fixed cost is shared by
// all warps of the CTA.
#pragma once

constexpr int TL_MAX_LEAVES = 32;
constexpr int TL_MAX_STAGES = 8;
constexpr int TL_MAX_CWARPS = 31;          // consumer warps (+ 1 producer warp <= 1024 threads)
constexpr int TL_HOT = 128;                // hot-list entries per tile
constexpr uint32_t TL_BCAP = 2048;         // boundary-table entries of one item kept in shared memory
constexpr uint32_t TF_LEAF_END = 1u, TF_TILE_END = 2u, TF_ITEM_END = 4u;

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
  uint32_t chunk;                  // postings per stage (multiple of 32)
  uint32_t stages;                 // ring depth (2..TL_MAX_STAGES)
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

struct __align__(16) TileMeta {   // 32 B: one 128-bit and one 64-bit shared load
  uint32_t t0;        // first document of the tile
  uint32_t n;         // postings staged (even; includes up to one posting of alignment padding on each side)
  uint32_t vbeg;      // the leaf's postings inside the stage: [vbeg, vend)
  uint32_t vend;
  float w;            // leaf weight
  uint32_t flags;
  uint32_t pad0, pad1;
};

__host__ __device__ inline size_t tile_smem_bytes(uint32_t tile_docs, uint32_t chunk, uint32_t stages, uint32_t cwarps) {
  // slots | ring | key buffer of the overflow scan | boundary table
  return (size_t)tile_docs * 8 + (size_t)stages * chunk * 8 + (size_t)cwarps * 32 * 8 + (size_t)TL_BCAP * 4;
}

__device__ __forceinline__ void tl_cbar(uint32_t nct) {
  asm volatile("bar.sync 1, %0;" ::"r"(nct) : "memory");
}

// Requires: flat OR (one group, no NOT clause), every leaf weight > 0, k <= 32 * KR, <= TL_MAX_LEAVES leaves,
// no after_key, (tiles of the item + 1) * leaves <= TL_BCAP, no postings of deleted documents in the store.
template <int KR>
__global__ void __launch_bounds__(1024, 1) k_score_tile(TileParams tp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ TileMeta s_meta[TL_MAX_STAGES];
  __shared__ __align__(8) unsigned long long s_full[TL_MAX_STAGES];
  __shared__ __align__(8) unsigned long long s_empty[TL_MAX_STAGES];
  __shared__ unsigned long long s_leaf_off[TL_MAX_LEAVES];
  __shared__ float s_leaf_w[TL_MAX_LEAVES];
  __shared__ unsigned short s_hot[TL_HOT];
  __shared__ uint32_t s_nhot[2];            // by tile parity: a slow warp still reads the count of tile t while
                                            // fast warps already push for tile t + 1
  __shared__ uint32_t s_nkeys;
  __shared__ float s_thr;
  __shared__ uint32_t s_item;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int NC = (int)(blockDim.x >> 5) - 1;            // consumer warps; the last warp produces
  const uint32_t NCT = (uint32_t)NC * 32u;
  const uint32_t T = tp.tile_docs, CH = tp.chunk, NS = tp.stages;

  const uint32_t acc_addr = smem_u32(smem_raw);
  const uint32_t ring_addr = acc_addr + T * 8u;
  const uint32_t keybuf_addr = ring_addr + NS * CH * 8u;
  uint32_t* s_bounds = reinterpret_cast<uint32_t*>(smem_raw + (size_t)T * 8 + (size_t)NS * CH * 8 + (size_t)NCT * 8);
  const uint32_t hot_addr = smem_u32(s_hot);
  const uint32_t thr_addr = smem_u32(&s_thr);
  const uint32_t nkeys_addr = smem_u32(&s_nkeys);

  for (uint32_t o = (uint32_t)tid * 16u; o < T * 8u; o += blockDim.x * 16u) sts_zero16(acc_addr + o);
  if (tid == 0) {
    for (uint32_t i = 0; i < NS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], (uint32_t)NC); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_nhot[0] = 0u; s_nhot[1] = 0u; s_nkeys = 0u;
    s_item = atomicAdd(tp.queue, 1u);
  }
  __syncthreads();

  uint32_t stage = 0, phase = 0;            // ring position: advances identically in the producer and the consumers
  uint32_t tag = 1u;                        // consumers: number of the tile being accumulated (slots start at 0)
  uint32_t rot = 0u;                        // consumers: which warp takes the first row of the next chunk

  for (;;) {
    const uint32_t item_idx = s_item;
    if (item_idx >= tp.n_items) break;
    const ItemRec item = tp.items[item_idx];
    const QueryRec q = tp.queries[item.q];
    const int L = (int)q.n_leaves;
    const uint32_t d_lo = item.tile_begin, d_hi = item.tile_end;
    const uint32_t nt = (d_hi - d_lo + T - 1u) / T;
    if (tid < L) {
      const LeafRec lf = tp.leaves[q.leaf_begin + tid];
      s_leaf_off[tid] = lf.off;
      s_leaf_w[tid] = lf.w;
    }
    {
      const uint32_t* __restrict__ src = tp.bounds + tp.item_boff[item_idx];
      const uint32_t n = (nt + 1u) * (uint32_t)L;
      for (uint32_t i = (uint32_t)tid; i < n; i += blockDim.x) s_bounds[i] = __ldg(src + i);
    }
    __syncthreads();                        // records and table are in shared memory; everybody has read s_item

    if (warp == NC) {
      // =============================== PRODUCER ===============================================
      if (lane == 0) {
        for (uint32_t j = 0; j < nt; ++j) {
          const uint32_t* blo = s_bounds + (size_t)j * L;
          const uint32_t* bhi = blo + L;
          int last = -1;
          for (int l = 0; l < L; ++l)
            if (blo[l] < bhi[l]) last = l;
          if (last < 0) continue;           // no posting in this tile: the consumers never see it
          const uint32_t t0 = d_lo + j * T;
          for (int l = 0; l <= last; ++l) {
            const uint32_t s = blo[l], e = bhi[l];
            if (s >= e) continue;
            const unsigned long long a = s_leaf_off[l] + s, b = s_leaf_off[l] + e;
            const float w = s_leaf_w[l];
            // bulk copies want 16-byte aligned addresses and sizes: stages start at an even posting index
            for (unsigned long long c0 = a & ~1ull; c0 < b; c0 += CH) {
              const unsigned long long rem = (b - c0 + 1ull) & ~1ull;
              const uint32_t n = (uint32_t)(rem < (unsigned long long)CH ? rem : (unsigned long long)CH);
              const bool lastc = c0 + CH >= b;
              mbar_wait(&s_empty[stage], phase ^ 1u);
              TileMeta m;
              m.t0 = t0;
              m.n = n;
              m.vbeg = (uint32_t)(a > c0 ? a - c0 : 0ull);
              m.vend = (uint32_t)(b - c0 < (unsigned long long)n ? b - c0 : (unsigned long long)n);
              m.w = w;
              m.flags = lastc ? (TF_LEAF_END | (l == last ? TF_TILE_END : 0u)) : 0u;
              m.pad0 = 0u; m.pad1 = 0u;
              s_meta[stage] = m;
              mbar_arrive_expect_tx(&s_full[stage], n * 8u);
              bulk_g2s(smem_raw + (size_t)T * 8 + (size_t)stage * CH * 8, tp.pairs + c0, n * 8u, &s_full[stage]);
              if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
          }
        }
        // terminator: the consumers leave the item when they see it
        mbar_wait(&s_empty[stage], phase ^ 1u);
        TileMeta m = {0u, 0u, 0u, 0u, 0.0f, TF_ITEM_END, 0u, 0u};
        s_meta[stage] = m;
        mbar_arrive(&s_full[stage]);
        if (++stage == NS) { stage = 0; phase ^= 1u; }
        s_item = atomicAdd(tp.queue, 1u);   // the next item (everybody read the current one before the barrier above)
      }
      stage = __shfl_sync(0xFFFFFFFFu, stage, 0);
      phase = __shfl_sync(0xFFFFFFFFu, phase, 0);
    } else {
      // ================================= CONSUMERS ============================================
      unsigned long long top[KR];           // warp 0 only: lane i, row j holds the (32 j + i)-th best key of the item
#pragma unroll
      for (int j = 0; j < KR; ++j) top[j] = 0ull;
      unsigned long long thr_key = 0ull;
      float thr = 1.17549435e-38f;          // FLT_MIN until k hits exist: every first hit is hot
      unsigned int tot = 0;
      if (tid == 0) s_thr = thr;            // read back only after a barrier
      for (;;) {
        mbar_wait(&s_full[stage], phase);
        const uint4 m0 = lds_v4(smem_u32(&s_meta[stage]));          // t0, n, vbeg, vend
        const uint2 m1 = lds_v2(smem_u32(&s_meta[stage]) + 16u);    // w, flags
        const uint32_t t0 = m0.x, flags = m1.y;
        if (m0.y) {
          const uint32_t rows = (m0.y + 31u) >> 5;
          const uint32_t vbeg = m0.z, vend = m0.w;
          const float w = __uint_as_float(m1.x);
          const uint32_t sbase = acc_addr - (t0 << 3);
          const uint32_t rbase = ring_addr + stage * CH * 8u + (uint32_t)lane * 8u;
          const uint32_t nhot_addr = smem_u32(&s_nhot[tag & 1u]);
          uint32_t r = (uint32_t)warp >= rot ? (uint32_t)warp - rot : (uint32_t)warp + (uint32_t)NC - rot;
          for (; r < rows; r += 2u * (uint32_t)NC) {
            // two rows per round: both stage loads, both slot loads, then the two updates (distinct documents)
            const uint32_t r1 = r + (uint32_t)NC;
            const uint32_t i0 = (r << 5) + (uint32_t)lane, i1 = (r1 << 5) + (uint32_t)lane;
            const bool ok0 = i0 >= vbeg && i0 < vend;
            const bool ok1 = r1 < rows && i1 >= vbeg && i1 < vend;
            uint2 p0 = make_uint2(0u, 0u), p1 = make_uint2(0u, 0u), v0 = make_uint2(0u, 0u), v1 = make_uint2(0u, 0u);
            if (ok0) p0 = lds_v2(rbase + (r << 8));
            if (ok1) p1 = lds_v2(rbase + (r1 << 8));
            const uint32_t a0 = sbase + (p0.x << 3), a1 = sbase + (p1.x << 3);
            if (ok0) v0 = lds_v2(a0);
            if (ok1) v1 = lds_v2(a1);
            if (ok0) {
              const bool fresh = v0.x != tag;
              const float old = fresh ? 0.0f : __uint_as_float(v0.y);
              const float nw = fmaf(w, __uint_as_float(p0.y), old);
              sts_v2(a0, tag, __float_as_uint(nw));
              tot += fresh ? 1u : 0u;
              if (nw >= thr && old < thr) {
                const uint32_t h = atoms_inc(nhot_addr);
                if (h < (uint32_t)TL_HOT) sts_u16(hot_addr + h * 2u, (a0 - acc_addr) >> 3);
              }
            }
            if (ok1) {
              const bool fresh = v1.x != tag;
              const float old = fresh ? 0.0f : __uint_as_float(v1.y);
              const float nw = fmaf(w, __uint_as_float(p1.y), old);
              sts_v2(a1, tag, __float_as_uint(nw));
              tot += fresh ? 1u : 0u;
              if (nw >= thr && old < thr) {
                const uint32_t h = atoms_inc(nhot_addr);
                if (h < (uint32_t)TL_HOT) sts_u16(hot_addr + h * 2u, (a1 - acc_addr) >> 3);
              }
            }
          }
          if (++rot == (uint32_t)NC) rot = 0u;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[stage]);     // this warp no longer reads the stage
        if (++stage == NS) { stage = 0; phase ^= 1u; }
        if (flags & TF_LEAF_END) tl_cbar(NCT);            // the next leaf may touch the same slots
        if (flags & TF_TILE_END) {
          // ---- tile epilogue: every warp is past the barrier, the slots of this tile are final ----------
          const uint32_t nhot_addr = smem_u32(&s_nhot[tag & 1u]);
          const uint32_t nhot = lds_u32(nhot_addr);       // frozen until tile + 2
          if (nhot) {
            if (nhot <= (uint32_t)TL_HOT) {
              if (warp == 0) {
                for (uint32_t j0 = 0; j0 < nhot; j0 += 32u) {
                  const uint32_t j = j0 + (uint32_t)lane;
                  unsigned long long key = 0ull;
                  if (j < nhot) {
                    const uint32_t slot = lds_u16(hot_addr + j * 2u);
                    const uint2 v = lds_v2(acc_addr + (slot << 3));
                    key = make_key(__uint_as_float(v.y), tp.doc_base + t0 + slot);
                  }
                  unsigned pm = __ballot_sync(0xFFFFFFFFu, key > thr_key);
                  while (pm) {
                    const int src = __ffs(pm) - 1;
                    pm &= pm - 1u;
                    const unsigned long long bk = __shfl_sync(0xFFFFFFFFu, key, src);
                    if (bk > thr_key) {
                      warp_topk_insert_rows<KR>(top, bk, lane);
                      thr_key = warp_topk_kth<KR>(top, tp.k);
                    }
                  }
                }
                if (lane == 0 && thr_key != 0ull) sts_f32(thr_addr, key_score(thr_key));
              }
              tl_cbar(NCT);                               // the threshold is published, the hot list is free
            } else {
              // No (tight) threshold yet: the hot list overflowed.  Scan the tile's slots, NCT at a time: everybody
              // offers its slot if it is a match at or above the threshold, warp 0 inserts the offered keys.
              const uint32_t Tn = min(T, d_hi - t0);
              const uint32_t ctid = (uint32_t)tid;
              for (uint32_t seg0 = 0; seg0 < Tn; seg0 += NCT) {
                const uint32_t s = seg0 + ctid;
                unsigned long long key = 0ull;
                if (s < Tn) {
                  const uint2 v = lds_v2(acc_addr + (s << 3));
                  if (v.x == tag && __uint_as_float(v.y) >= thr) key = make_key(__uint_as_float(v.y), tp.doc_base + t0 + s);
                }
                const unsigned mk = __ballot_sync(0xFFFFFFFFu, key != 0ull);
                if (mk) {
                  uint32_t base = 0u;
                  if (lane == 0) base = atomicAdd(&s_nkeys, (uint32_t)__popc(mk));
                  base = __shfl_sync(0xFFFFFFFFu, base, 0);
                  if (key != 0ull) {
                    const uint32_t at = keybuf_addr + ((base + (uint32_t)__popc(mk & ((1u << lane) - 1u))) << 3);
                    sts_v2(at, (uint32_t)key, (uint32_t)(key >> 32));
                  }
                }
                tl_cbar(NCT);
                if (warp == 0) {
                  const uint32_t n = lds_u32(nkeys_addr);
                  for (uint32_t j0 = 0; j0 < n; j0 += 32u) {
                    const uint32_t j = j0 + (uint32_t)lane;
                    unsigned long long key2 = 0ull;
                    if (j < n) {
                      const uint2 kv = lds_v2(keybuf_addr + (j << 3));
                      key2 = ((unsigned long long)kv.y << 32) | kv.x;
                    }
                    unsigned pm = __ballot_sync(0xFFFFFFFFu, key2 > thr_key);
                    while (pm) {
                      const int src = __ffs(pm) - 1;
                      pm &= pm - 1u;
                      const unsigned long long bk = __shfl_sync(0xFFFFFFFFu, key2, src);
                      if (bk > thr_key) {
                        warp_topk_insert_rows<KR>(top, bk, lane);
                        thr_key = warp_topk_kth<KR>(top, tp.k);
                      }
                    }
                  }
                  if (lane == 0) {
                    sts_u32(nkeys_addr, 0u);
                    if (thr_key != 0ull) sts_f32(thr_addr, key_score(thr_key));
                  }
                }
                tl_cbar(NCT);
                thr = lds_f32(thr_addr);
              }
            }
            thr = lds_f32(thr_addr);
            if (tid == 0) sts_u32(nhot_addr, 0u);         // next used by tile + 2, after the barriers of tile + 1
          }
          ++tag;
        }
        if (flags & TF_ITEM_END) break;
      }
      // ---- item epilogue ------------------------------------------------------------------------
      if (warp == 0) {
        unsigned long long* out = tp.part_keys + (size_t)item.part * tp.k;
#pragma unroll
        for (int j = 0; j < KR; ++j)
          if (32 * j + lane < tp.k) out[32 * j + lane] = top[j];
      }
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xFFFFFFFFu, tot, o);
      if (lane == 0 && tot) atomicAdd(tp.totals + item.q, (unsigned long long)tot);
    }
    __syncthreads();                        // the item is done; s_item holds the next one
  }
}
