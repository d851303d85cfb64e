/*
 * match_oracle.c — CPU restatement of openMVG's BRUTE_FORCE_L2 putative matcher.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this.  The product path (openmvg_b200/csrc) never calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_match.py checks this file against
 *   (a) the reference's own known-answer tests (matching/metric_test.cpp:31-39: L2 of the two
 *       8-vectors = 168; matching/matching_test.cpp toy arrays), and
 *   (b) oracle/_ref/libref_match.so — the reference's own headers compiled where they lie —
 *       and the committed golden vectors under tests/golden/ generated from it.
 *
 * Reference lines restated (all relative to /root/reference/src/openMVG):
 *   matching/metric.hpp:55-93              L2<uint8_t>::operator(): sum of (a-b)^2 as int
 *   matching/matcher_brute_force.hpp:100-144 SearchNeighbours: false if NN > rows or nbQuery < 1
 *   matching/matcher_brute_force.hpp:163-200 per query: distance to every db row, 2 smallest
 *   stl/indexed_sort.hpp:48-63             partial_sort on the value only (ties: unspecified order)
 *   matching/matching_filters.hpp:38-60    keep i iff (float)d1 < fratio * (float)d2
 *   matching/regions_matcher.hpp:162-207   fratio = ratio*ratio (float); emit (idx in I, idx in J)
 *   matching_image_collection/Matcher_Regions.cpp:32-107  pair loop; empty results not inserted
 */
#include <stdint.h>
#include <stddef.h>
#include <limits.h>

#define OMVG_DESC_LEN 128

/* metric.hpp:55-93 — plain integer accumulation, result type int. */
int oracle_l2_u8(const uint8_t *a, const uint8_t *b, int n)
{
    int acc = 0;
    for (int k = 0; k < n; ++k) {
        const int d = (int)a[k] - (int)b[k];
        acc += d * d;
    }
    return acc;
}

/*
 * 2-NN of every query (rows of J) in the database (rows of I).
 * d1[q] = smallest distance, i1[q] = its row in I, d2[q] = second smallest WITH multiplicity
 * (matcher_brute_force.hpp:185-197 copies packet_vec[0], packet_vec[1] after a partial_sort).
 * Tie rule for i1 when the minimum occurs twice: lowest index.  The reference leaves it to
 * std::partial_sort; it cannot influence the ratio-filtered output for ratio <= 1 because a tie
 * d1 == d2 never passes d1 < fratio*d2.
 * Returns 0 on success, -1 if the reference's SearchNeighbours would return false.
 */
int oracle_top2(const uint8_t *db, uint32_t n_db, const uint8_t *q, uint32_t n_q,
                int32_t *d1, uint32_t *i1, int32_t *d2)
{
    if (db == NULL || n_db < 2 || n_q < 1) return -1;   /* :108-113  NN(=2) > rows  or nbQuery < 1 */
    #pragma omp parallel for schedule(static)
    for (int64_t qi = 0; qi < (int64_t)n_q; ++qi) {
        const uint8_t *qp = q + (size_t)qi * OMVG_DESC_LEN;
        int32_t b1 = INT_MAX, b2 = INT_MAX;
        uint32_t bi = 0;
        for (uint32_t r = 0; r < n_db; ++r) {
            const int32_t d = oracle_l2_u8(qp, db + (size_t)r * OMVG_DESC_LEN, OMVG_DESC_LEN);
            if (d < b1) { b2 = b1; b1 = d; bi = r; }
            else if (d < b2) { b2 = d; }
        }
        d1[qi] = b1; i1[qi] = bi; d2[qi] = b2;
    }
    return 0;
}

/* matching_filters.hpp:57 with NN=2: int < float*int evaluated in float. */
int oracle_ratio_keep(int32_t d1, int32_t d2, float fratio)
{
    volatile float rhs = fratio * (float)d2;   /* one IEEE single multiply, no FMA contraction */
    return (float)d1 < rhs;
}

/*
 * One image pair.  I = database, J = queries (Matcher_Regions.cpp:73,93).
 * out_ij receives (i_ = row in I, j_ = row in J) pairs in ascending j_ (regions_matcher.hpp:198-204).
 * scratch: 3*n_j 32-bit words.  Returns the number of kept matches (0 => the reference inserts
 * nothing for this pair, Matcher_Regions.cpp:99-102).
 */
int64_t oracle_match_pair(const uint8_t *desc_i, uint32_t n_i, const uint8_t *desc_j, uint32_t n_j,
                          float dist_ratio, uint32_t *out_ij, int32_t *scratch)
{
    if (n_i == 0 || n_j == 0) return 0;                 /* Matcher_Regions.cpp:65-69,85-90 */
    int32_t  *d1 = scratch;
    uint32_t *i1 = (uint32_t *)(scratch + n_j);
    int32_t  *d2 = scratch + 2 * (size_t)n_j;
    if (oracle_top2(desc_i, n_i, desc_j, n_j, d1, i1, d2) != 0) return 0;
    const float fratio = dist_ratio * dist_ratio;        /* numeric.h:56 Square<float> */
    int64_t n = 0;
    for (uint32_t q = 0; q < n_j; ++q) {
        if (oracle_ratio_keep(d1[q], d2[q], fratio)) {
            out_ij[2 * n + 0] = i1[q];
            out_ij[2 * n + 1] = q;
            ++n;
        }
    }
    return n;
}

/* FNV-1a over the (i,j) list exactly as the survey probe hashed the reference output. */
uint64_t oracle_fnv1a_ij(const uint32_t *ij, int64_t n_matches)
{
    uint64_t h = 1469598103934665603ull;
    for (int64_t k = 0; k < n_matches; ++k) {
        h = (h ^ ij[2 * k + 0]) * 1099511628211ull;
        h = (h ^ ij[2 * k + 1]) * 1099511628211ull;
    }
    return h;
}
