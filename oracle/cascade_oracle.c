/* cascade_oracle.c — CPU restatement of openMVG's cascade-hashing matcher.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as match_oracle.c): only tests/, __graft_entry__.smoke() and
 * bench.py's cpu legs may load this.  The product path never calls it.
 *
 * Restated from /root/reference/src/openMVG:
 *   matching/cascade_hasher.hpp:166-176   GetZeroMeanDescriptor (float column means)
 *   matching/cascade_hasher.hpp:179-247   CreateHashedDescriptions: 128-bit code = sign(P (d - mean)),
 *                                         6 bucket ids of 10 bits = sign(S_g (d - mean)), MSB first; buckets hold
 *                                         descriptor ids in ascending order
 *   matching/cascade_hasher.hpp:251-370   Match_HashedDescriptions: union of the 6 buckets, "<= NN candidates:
 *                                         skip", first-occurrence de-duplication, Hamming distance bins in
 *                                         insertion order, first 10 by (Hamming, insertion), exact L2, top-2 by
 *                                         (distance, id) (std::partial_sort of pair<dist,int>)
 *   matching_image_collection/Cascade_Hashing_Matcher_Regions.cpp:78-105  one zero-mean vector for the whole
 *                                         collection: mean over the USED images of the per-image means (empty
 *                                         images contribute a zero row)
 *   :150-189                              queries = image J, database = image I, ratio test
 *                                         float(d1) < ratio^2 * float(d2)  (matching_filters.hpp:38-60)
 * The projections (std::mt19937 + std::normal_distribution, cascade_hasher.hpp:142-162) are INPUTS here, as in
 * the C ABI: the caller generates them with the very same std:: facilities.
 *
 * Parity status: the matching stage is exact integer work and is pinned bit-for-bit against the compiled
 * reference when both are fed the same hash codes; the hashing stage is float mat-vec whose summation order the
 * reference leaves to Eigen's vectorised GEMV, so a projection within rounding of zero can flip a bit: parity of
 * the whole pipeline against the reference is statistical (tests/test_oracle_match.py measures it), as SURVEY
 * M9 states.  The order fixed here (and in the CUDA kernel): k = 0..127 sequentially, separate multiply and
 * add, round-to-nearest.  Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CH_DIM 128
#define CH_GROUPS 6
#define CH_BITS 10
#define CH_TOP 10

/* per-image float column means: float(sum) / float(n)  (integer sums are exact in float: 255*n < 2^24 for n < 65793;
 * larger images would need Eigen's exact summation order and are outside the pinned range) */
void oracle_cascade_image_mean(const uint8_t *desc, uint32_t n, float *mean /*[128]*/)
{
  for (int k = 0; k < CH_DIM; ++k) {
    uint64_t s = 0;
    for (uint32_t r = 0; r < n; ++r) s += desc[(size_t)r * CH_DIM + k];
    mean[k] = n ? (float)s / (float)n : 0.0f;
  }
}

/* zero-mean descriptor of the collection: sequential float sum over the used images' means, / float(n_used) */
void oracle_cascade_zero_mean(const float *image_means /*[n_used][128]*/, uint32_t n_used, float *zm /*[128]*/)
{
  for (int k = 0; k < CH_DIM; ++k) {
    float s = 0.0f;
    for (uint32_t i = 0; i < n_used; ++i) s = s + image_means[(size_t)i * CH_DIM + k];
    zm[k] = n_used ? s / (float)n_used : 0.0f;
  }
}

/* codes[n][4] (bit j of the code = bit (j & 31) of word j >> 5), bids[n][6] */
void oracle_cascade_hash(const uint8_t *desc, uint32_t n, const float *zm, const float *primary /*[128][128]*/,
                         const float *secondary /*[6][10][128]*/, uint32_t *codes, uint16_t *bids)
{
  for (uint32_t r = 0; r < n; ++r) {
    float d[CH_DIM];
    for (int k = 0; k < CH_DIM; ++k) d[k] = (float)desc[(size_t)r * CH_DIM + k] - zm[k];
    uint32_t *code = codes + 4 * (size_t)r;
    code[0] = code[1] = code[2] = code[3] = 0;
    for (int j = 0; j < CH_DIM; ++j) {
      float acc = 0.0f;
      for (int k = 0; k < CH_DIM; ++k) { const float p = primary[j * CH_DIM + k] * d[k]; acc = acc + p; }
      if (acc > 0.0f) code[j >> 5] |= 1u << (j & 31);
    }
    for (int g = 0; g < CH_GROUPS; ++g) {
      uint16_t id = 0;
      for (int b = 0; b < CH_BITS; ++b) {
        const float *row = secondary + ((size_t)g * CH_BITS + b) * CH_DIM;
        float acc = 0.0f;
        for (int k = 0; k < CH_DIM; ++k) { const float p = row[k] * d[k]; acc = acc + p; }
        id = (uint16_t)((id << 1) + (acc > 0.0f ? 1 : 0));
      }
      bids[CH_GROUPS * (size_t)r + g] = id;
    }
  }
}

static int l2_u8(const uint8_t *a, const uint8_t *b)
{
  int s = 0;
  for (int k = 0; k < CH_DIM; ++k) { const int d = (int)a[k] - (int)b[k]; s += d * d; }
  return s;
}

/* queries = image J (desc_j, codes_j, bids_j), database = image I.  Emits (i, j) for every query that passes the
 * ratio test, in ascending j (the reference then sorts by (i, j); the caller does that).  Returns the count. */
int64_t oracle_cascade_match_pair(const uint8_t *desc_i, uint32_t n_i, const uint32_t *codes_i, const uint16_t *bids_i,
                                  const uint8_t *desc_j, uint32_t n_j, const uint32_t *codes_j, const uint16_t *bids_j,
                                  float ratio, uint32_t *out_ij)
{
  const float fratio = ratio * ratio;
  /* buckets of I: counting sort per group, ascending id inside a bucket */
  const int NB = 1 << CH_BITS;
  uint32_t *start = (uint32_t *)calloc((size_t)CH_GROUPS * (NB + 1), sizeof(uint32_t));
  uint32_t *items = (uint32_t *)malloc((size_t)CH_GROUPS * (n_i ? n_i : 1) * sizeof(uint32_t));
  for (int g = 0; g < CH_GROUPS; ++g) {
    uint32_t *st = start + (size_t)g * (NB + 1);
    for (uint32_t r = 0; r < n_i; ++r) st[bids_i[CH_GROUPS * (size_t)r + g] + 1]++;
    for (int b = 0; b < NB; ++b) st[b + 1] += st[b];
    uint32_t *cur = (uint32_t *)malloc(NB * sizeof(uint32_t));
    memcpy(cur, st, NB * sizeof(uint32_t));
    for (uint32_t r = 0; r < n_i; ++r) items[(size_t)g * n_i + cur[bids_i[CH_GROUPS * (size_t)r + g]]++] = r;
    free(cur);
  }
  uint32_t *cand = (uint32_t *)malloc((size_t)CH_GROUPS * (n_i ? n_i : 1) * sizeof(uint32_t));
  uint8_t *used = (uint8_t *)calloc(n_i ? n_i : 1, 1);
  uint32_t *bins = (uint32_t *)malloc((size_t)(CH_DIM + 1) * (n_i ? n_i : 1) * sizeof(uint32_t));   /* [hamming][k] */
  uint32_t nbin[CH_DIM + 1];
  int64_t n_out = 0;
  for (uint32_t q = 0; q < n_j; ++q) {
    uint32_t nc = 0;
    for (int g = 0; g < CH_GROUPS; ++g) {
      const uint32_t *st = start + (size_t)g * (NB + 1); const uint16_t b = bids_j[CH_GROUPS * (size_t)q + g];
      for (uint32_t t = st[b]; t < st[b + 1]; ++t) { const uint32_t id = items[(size_t)g * n_i + t]; cand[nc++] = id; used[id] = 0; }
    }
    if (nc <= 2) continue;                                    /* "not at least NN candidates" (:301-304) */
    memset(nbin, 0, sizeof nbin);
    for (uint32_t t = 0; t < nc; ++t) {
      const uint32_t id = cand[t];
      if (used[id]) continue;
      used[id] = 1;
      int h = 0;
      for (int w = 0; w < 4; ++w) h += __builtin_popcount(codes_j[4 * (size_t)q + w] ^ codes_i[4 * (size_t)id + w]);
      bins[(size_t)h * n_i + nbin[h]++] = id;
    }
    int nd = 0; int dist[CH_TOP]; uint32_t ids[CH_TOP];
    for (int h = 0; h <= CH_DIM && nd < CH_TOP; ++h)
      for (uint32_t k = 0; k < nbin[h] && nd < CH_TOP; ++k) {
        const uint32_t id = bins[(size_t)h * n_i + k];
        dist[nd] = l2_u8(desc_i + (size_t)id * CH_DIM, desc_j + (size_t)q * CH_DIM); ids[nd] = id; ++nd;
      }
    if (nd < 2) continue;
    /* top-2 of pair<distance, id> in lexicographic order (std::partial_sort, :352-355) */
    int b1 = -1, b2 = -1;
    for (int t = 0; t < nd; ++t) {
      if (b1 < 0 || dist[t] < dist[b1] || (dist[t] == dist[b1] && ids[t] < ids[b1])) { b2 = b1; b1 = t; }
      else if (b2 < 0 || dist[t] < dist[b2] || (dist[t] == dist[b2] && ids[t] < ids[b2])) b2 = t;
    }
    if ((float)dist[b1] < fratio * (float)dist[b2]) { out_ij[2 * n_out] = ids[b1]; out_ij[2 * n_out + 1] = q; ++n_out; }
  }
  free(start); free(items); free(cand); free(used); free(bins);
  return n_out;
}
