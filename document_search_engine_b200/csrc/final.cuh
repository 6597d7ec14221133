// Helpers of the final() step (a weighting whose final(searcher, docnum, score) is applied to every match
// before the top-k, W14): the value itself and the 96-bit keys it is ranked by.  Included by bm25f.cu before
// the scoring kernels.
#pragma once

// ---- DateBM25F.final (my_whoosh.py:129-146), applied to every match before the top-k (W14) -------
//   s' = 1 - 1 / s;  a dated document:  s' = (s' + (date seconds + 1.0)) / 10**9   -- float64, IEEE division
// `add` is (date seconds + 1.0), precomputed in float64 on the host.
__device__ __forceinline__ double final_value(float score, double add) {
  const double t = 1.0 - 1.0 / (double)score;
  return isnan(add) ? t : (t + add) / 1e9;
}
__host__ __device__ __forceinline__ unsigned long long orderable_f64(double v) {
#ifdef __CUDA_ARCH__
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
#else
  unsigned long long u;
  memcpy(&u, &v, 8);
#endif
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double orderable_f64_value(unsigned long long u) {
  u = (u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double v;
  memcpy(&v, &u, 8);
  return v;
#endif
}
// 96-bit keys (high: orderable final value, low: ~global docnum), KR per lane, the same sorted list as
// warp_topk_insert_rows keeps for 64-bit keys
__device__ __forceinline__ bool key2_gt(unsigned long long ah, uint32_t al, unsigned long long bh, uint32_t bl) {
  return ah > bh || (ah == bh && al > bl);
}
// ... a candidate: above the k-th best so far and, when the query carries a paging bound (the last hit of the
// previous pass: search_page deep into a date-ordered listing, my_flask.py:211), strictly below that bound
__device__ __forceinline__ bool key2_wanted(unsigned long long kh, uint32_t kl, unsigned long long th, uint32_t tl,
                                            unsigned long long bh, uint32_t bl) {
  return key2_gt(kh, kl, th, tl) && (bh == 0ull || key2_gt(bh, bl, kh, kl));
}
template <int KR>
__device__ __forceinline__ void warp_topk2_insert_rows(unsigned long long (&th)[KR], uint32_t (&tl)[KR], unsigned long long kh,
                                                       uint32_t kl, int lane) {
  bool inserted = false;
  unsigned long long ch = 0ull;
  uint32_t cl = 0u;
#pragma unroll
  for (int j = 0; j < KR; ++j) {
    const unsigned long long last_h = __shfl_sync(0xFFFFFFFFu, th[j], 31);
    const uint32_t last_l = __shfl_sync(0xFFFFFFFFu, tl[j], 31);
    const unsigned long long up_h = __shfl_up_sync(0xFFFFFFFFu, th[j], 1);
    const uint32_t up_l = __shfl_up_sync(0xFFFFFFFFu, tl[j], 1);
    if (!inserted) {
      const int pos = __popc(__ballot_sync(0xFFFFFFFFu, key2_gt(th[j], tl[j], kh, kl)));
      if (pos < 32) {
        if (lane > pos) { th[j] = up_h; tl[j] = up_l; }
        if (lane == pos) { th[j] = kh; tl[j] = kl; }
        ch = last_h;
        cl = last_l;
        inserted = true;
      }
    } else {
      th[j] = (lane == 0) ? ch : up_h;
      tl[j] = (lane == 0) ? cl : up_l;
      ch = last_h;
      cl = last_l;
    }
  }
}
template <int KR>
__device__ __forceinline__ void warp_topk2_kth(const unsigned long long (&th)[KR], const uint32_t (&tl)[KR], int k,
                                               unsigned long long& vh, uint32_t& vl) {
  vh = 0ull;
  vl = 0u;
#pragma unroll
  for (int j = 0; j < KR; ++j) {
    const unsigned long long h = __shfl_sync(0xFFFFFFFFu, th[j], (k - 1) & 31);
    const uint32_t l = __shfl_sync(0xFFFFFFFFu, tl[j], (k - 1) & 31);
    if (j == ((k - 1) >> 5)) { vh = h; vl = l; }
  }
}

// Requires: k <= 32 * KR, <= 32 leaves, <= 32 groups, every leaf weight > 0, no after_key, no postings of
// deleted documents in the store.
