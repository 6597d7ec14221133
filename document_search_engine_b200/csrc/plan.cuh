// Batch planning on the device.
//
// The host planner (prepare_impl in bm25f.cu) resolves every leaf's posting list, orders a query's groups, routes the
// query to a kernel class and cuts it into work items; at 10k queries a step this is ~1 ms of host work, more than the
// GPU needs to score the batch on a 1/8 document shard.  These three kernels do the same work for the batches the
// warp kernels serve alone (at most ST_MAX_LEAVES leaves a query, positive weights, no NOT clause, no paging bound, no
// final() step, k <= FAST_MAX_K): the host only copies the caller's five arrays to the device.
//
//   k_plan_queries  one thread per query: leaf records (groups smallest-first, the host planner's order), query record,
//                   kernel class, number of items, weight bucket; counts items per (class, bucket)
//   k_plan_scan     one CTA: first partial list of every query (exclusive scan of the item counts), position of every
//                   (class, bucket) in the item array (heaviest bucket first: longest-processing-time order), item
//                   counts per class for the persistent kernels (StreamParams::n_items_dev and friends)
//   k_plan_items    one thread per query: writes the query's items at its bucket's cursor
//
// The routing rules, the split formulas and the item geometry are the host planner's: keep the two in step
// (tests/test_gpu_plan.py compares the results of both on the same batches).
#pragma once

constexpr int PL_NB = 256;                    // weight buckets per class
constexpr int PL_CLASSES = 3;                 // 0: warp streams, 1: warp teams, 2: candidate-driven
constexpr int PL_CTR_ITEMS = 0;               // ctr[0..2]: items per class
constexpr int PL_CTR_OFF = 4;                 // ctr[4..6]: first item of the class in the item array
constexpr int PL_CTR_PARTS = 7;               // ctr[7]: partial lists of the batch
constexpr int PL_CTR_BUCKETS = 8;             // ctr[8 + class * PL_NB + bucket]: count, then cursor
constexpr int PL_CTR_WORDS = PL_CTR_BUCKETS + PL_CLASSES * PL_NB;
constexpr int PL_POST_WORDS = 4;              // u64 after the counters: postings of the batch, then per class
constexpr size_t PL_CTR_BYTES = (size_t)PL_CTR_WORDS * 4 + PL_POST_WORDS * 8;
constexpr int PL_MAX_LEAVES = 8;              // = ST_MAX_LEAVES

struct PlanParams {
  const unsigned long long* term_offsets;     // [n_terms + 1] (Every(field) lists included)
  const uint8_t* term_field;                  // [n_terms]
  const uint32_t* q_off;                      // the caller's batch arrays (bm25f_query_batch), on the device
  const uint8_t* q_ng;
  const uint32_t* leaf_term;
  const float* leaf_w;
  const uint8_t* leaf_g;
  LeafRec* leaves;
  QueryRec* queries;
  ItemRec* items;
  uint4* tmp;                                 // per query: class (3: none), items, bucket, flags
  unsigned int* ctr;
  uint32_t Q, n_docs, n_fields, n_real_terms, max_split;
  uint32_t wsplit, is_split, tl_split, is_or_limit, is_ratio, st_slot_bytes, tl_slot_bytes;
  int k;
};

__device__ __forceinline__ int plan_bucket(unsigned long long x) {
  if (x < 4) return (int)x;
  const int lg = 63 - __clzll((long long)x);
  return min(PL_NB - 1, lg * 4 + (int)((x >> (lg - 2)) & 3));
}

__global__ void __launch_bounds__(128) k_plan_queries(PlanParams pp) {
  const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= pp.Q) return;
  const uint32_t a = pp.q_off[qi], e = pp.q_off[qi + 1], nl = e - a, G = pp.q_ng[qi];
  QueryRec qr;
  qr.leaf_begin = a;
  qr.n_leaves = 0;
  qr.n_groups = 0;
  qr.flags = 0;
  qr.after_lo = 0;
  qr.after_key = 0ull;
  qr.part_begin = 0;
  qr.n_parts = 0;
  unsigned long long loff[PL_MAX_LEAVES];
  uint32_t ldf[PL_MAX_LEAVES], lgrp[PL_MAX_LEAVES], lfield[PL_MAX_LEAVES];
  unsigned long long gsize[PL_MAX_LEAVES];
  uint32_t seen = 0;
#pragma unroll
  for (int g = 0; g < PL_MAX_LEAVES; ++g) gsize[g] = 0;
  for (uint32_t i = 0; i < nl; ++i) {
    uint32_t term = pp.leaf_term[a + i];
    if (term >= BM25F_TERM_EVERY_BASE && term != BM25F_TERM_UNKNOWN) {
      const uint32_t f = term - BM25F_TERM_EVERY_BASE;
      term = f < pp.n_fields ? pp.n_real_terms + f : BM25F_TERM_UNKNOWN;
    }
    const uint32_t g = pp.leaf_g[a + i];
    lgrp[i] = g;
    ldf[i] = 0;
    loff[i] = 0;
    lfield[i] = 0;
    seen |= 1u << g;
    if (term != BM25F_TERM_UNKNOWN) {
      const unsigned long long o = pp.term_offsets[term];
      loff[i] = o;
      ldf[i] = (uint32_t)(pp.term_offsets[term + 1] - o);
      lfield[i] = pp.term_field[term];
      gsize[g] += ldf[i];
    }
  }
  bool dead = (G == 0 || nl == 0 || seen != (G >= 32 ? 0xFFFFFFFFu : (1u << G) - 1u));
  for (uint32_t g = 0; g < G && !dead; ++g)
    if (gsize[g] == 0) dead = true;              // an empty group: the AND matches nothing (W10)
  uint32_t nlq = 0;
  unsigned long long P = 0, g0 = 0;
  if (!dead) {
    // groups smallest-first, ties in input order (the host planner's stable sort)
    for (uint32_t r = 0; r < G; ++r) {
      uint32_t gsel = 0;
      for (uint32_t g = 0; g < G; ++g) {
        uint32_t rank = 0;
        for (uint32_t o = 0; o < G; ++o) rank += (gsize[o] < gsize[g] || (gsize[o] == gsize[g] && o < g)) ? 1u : 0u;
        if (rank == r) gsel = g;
      }
      if (r == 0) g0 = gsize[gsel];
      for (uint32_t i = 0; i < nl; ++i) {
        if (lgrp[i] != gsel || ldf[i] == 0) continue;
        LeafRec lf;
        lf.off = loff[i];
        lf.df = ldf[i];
        lf.w = pp.leaf_w[a + i];
        lf.norm_off = lfield[i] * 256u;
        lf.group = r;
        lf.qleaf0 = a;
        lf.qnl = 0;
        pp.leaves[a + nlq] = lf;
        P += ldf[i];
        ++nlq;
      }
    }
    for (uint32_t i = 0; i < nlq; ++i) pp.leaves[a + i].qnl = nlq;
  }
  for (uint32_t i = a + nlq; i < e; ++i) {       // unused slots stay harmless
    LeafRec lf;
    lf.off = 0;
    lf.df = 0;
    lf.w = 0.0f;
    lf.norm_off = 0;
    lf.group = 0;
    lf.qleaf0 = i;
    lf.qnl = 1;
    pp.leaves[i] = lf;
  }
  uint4 t = make_uint4(3u, 0u, 0u, 0u);
  if (!dead) {
    qr.n_leaves = (uint16_t)nlq;
    qr.n_groups = (uint16_t)G;
    qr.flags = G == 1 ? QF_SIMPLE_OR : 0u;
    const bool simple = G == 1;
    const unsigned long long n_cand = simple ? P * (unsigned long long)(nlq > 1 ? nlq - 1 : 1) : g0 * (unsigned long long)(nlq - 1);
    const bool use_isect = simple ? n_cand < (unsigned long long)pp.is_or_limit : g0 * (unsigned long long)(nlq - 1) * pp.is_ratio < P;
    const bool use_team = !use_isect && pp.k <= 32 && !simple;
    unsigned long long nsplit, weight;
    uint32_t cls;
    if (use_isect) {
      cls = 2;
      nsplit = min(max(1ull, (unsigned long long)pp.n_docs / 256), max(1ull, (n_cand + pp.is_split) / (2ull * pp.is_split)));
      nsplit = min(nsplit, (unsigned long long)pp.max_split);
      weight = n_cand / nsplit + 64;
    } else if (use_team) {
      cls = 1;
      const unsigned long long sw = pp.tl_slot_bytes / 8u;
      const unsigned long long nsl = ((unsigned long long)pp.n_docs + sw - 1) / sw;
      const unsigned long long work = P + nsl * (16ull * nlq + 24ull);
      nsplit = min(max(1ull, nsl / 32), max(1ull, (work + pp.tl_split / 2) / pp.tl_split));
      nsplit = min(nsplit, (unsigned long long)pp.max_split);
      weight = work / nsplit;
    } else {
      cls = 0;
      const unsigned long long sw = pp.st_slot_bytes / (simple ? 4u : 8u);
      const unsigned long long nsub = ((unsigned long long)pp.n_docs + sw - 1) / sw;
      const unsigned long long work = P + nsub * (16ull * nlq + 24ull);
      nsplit = min(max(1ull, (unsigned long long)pp.n_docs / 1024), max(1ull, (work + pp.wsplit / 2) / pp.wsplit));
      nsplit = min(nsplit, (unsigned long long)pp.max_split);
      weight = work / nsplit;
    }
    qr.n_parts = (uint32_t)nsplit;
    t = make_uint4(cls, (uint32_t)nsplit, (uint32_t)(PL_NB - 1 - plan_bucket(weight)), qr.flags);
    atomicAdd(pp.ctr + PL_CTR_BUCKETS + cls * PL_NB + t.z, (unsigned int)nsplit);
    unsigned long long* post = reinterpret_cast<unsigned long long*>(pp.ctr + PL_CTR_WORDS);
    atomicAdd(post, P);
    atomicAdd(post + 1 + cls, P);
  }
  pp.queries[qi] = qr;
  pp.tmp[qi] = t;
}

__global__ void __launch_bounds__(1024) k_plan_scan(PlanParams pp) {
  __shared__ unsigned int s_warp[32];
  __shared__ unsigned int s_cnt[PL_CLASSES * PL_NB];
  __shared__ unsigned int s_cls[PL_CLASSES];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // (1) first partial list of every query
  const uint32_t per = (pp.Q + 1023) / 1024;
  const uint32_t lo = min(pp.Q, tid * per), hi = min(pp.Q, lo + per);
  unsigned int sum = 0;
  for (uint32_t q = lo; q < hi; ++q) sum += pp.tmp[q].y;
  unsigned int incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if ((int)lane >= d) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  if (tid < PL_CLASSES * PL_NB) s_cnt[tid] = pp.ctr[PL_CTR_BUCKETS + tid];
  __syncthreads();
  if (warp == 0) {
    unsigned int w = s_warp[lane], wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned int v = __shfl_up_sync(0xFFFFFFFFu, wi, d);
      if ((int)lane >= d) wi += v;
    }
    s_warp[lane] = wi - w;                       // exclusive
    if (lane == 31) pp.ctr[PL_CTR_PARTS] = wi;
  }
  if (tid < PL_CLASSES) {
    unsigned int c = 0;
    for (int b = 0; b < PL_NB; ++b) c += s_cnt[tid * PL_NB + b];
    s_cls[tid] = c;
  }
  __syncthreads();
  unsigned int base = s_warp[warp] + incl - sum;
  for (uint32_t q = lo; q < hi; ++q) {
    pp.queries[q].part_begin = base;
    base += pp.tmp[q].y;
  }
  // (2) where every (class, bucket) starts in the item array; items per class
  if (tid < PL_CLASSES * PL_NB) {
    const uint32_t c = tid / PL_NB, b = tid % PL_NB;
    unsigned int pos = 0;
    for (uint32_t o = 0; o < c; ++o) pos += s_cls[o];
    for (uint32_t o = 0; o < b; ++o) pos += s_cnt[c * PL_NB + o];
    pp.ctr[PL_CTR_BUCKETS + tid] = pos;
    if (b == 0) {
      pp.ctr[PL_CTR_OFF + c] = pos;
      pp.ctr[PL_CTR_ITEMS + c] = s_cls[c];
    }
  }
}

__global__ void __launch_bounds__(128) k_plan_items(PlanParams pp) {
  const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= pp.Q) return;
  const uint4 t = pp.tmp[qi];
  if (t.x >= (uint32_t)PL_CLASSES) return;
  const uint32_t nsplit = t.y;
  const uint32_t pos = atomicAdd(pp.ctr + PL_CTR_BUCKETS + t.x * PL_NB + t.z, nsplit);
  const uint32_t part0 = pp.queries[qi].part_begin;
  const unsigned long long n_docs = pp.n_docs;
  unsigned long long sw = 1, nsl = n_docs;
  if (t.x == 1) {                                // warp teams: slice-aligned document ranges
    sw = pp.tl_slot_bytes / ((t.w & QF_SIMPLE_OR) ? 4u : 8u);
    nsl = (n_docs + sw - 1) / sw;
  }
  for (uint32_t s = 0; s < nsplit; ++s) {
    ItemRec it;
    it.q = qi;
    if (t.x == 1) {
      it.tile_begin = (uint32_t)((nsl * s / nsplit) * sw);
      it.tile_end = (uint32_t)min(n_docs, (nsl * (s + 1) / nsplit) * sw);
    } else {
      it.tile_begin = (uint32_t)(n_docs * s / nsplit);
      it.tile_end = (uint32_t)(n_docs * (s + 1) / nsplit);
    }
    it.part = part0 + s;
    pp.items[pos + s] = it;
  }
}
