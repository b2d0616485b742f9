/* orb_oracle_bow.c -- CPU ORACLE (test infrastructure only; never linked into liborbfe.so) for the two "next" rows
 * of SURVEY.md section 8(f):
 *   DBoW2 vocabulary transform   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1262 (+ BowVector.cpp:34-84,
 *                                FeatureVector.cpp:32-48, FORB.cpp:79-99)
 *   MapPoint::ComputeDistinctiveDescriptors   src/MapPoint.cc:185-250
 * Parity unpinned: DBoW2 and MapPoint.cc need OpenCV 2.4 / Boost headers (not in this image), and the reference ships
 * no tests or golden vectors for them; the loops below follow the cited lines statement by statement.
 * The std::map containers are restated as sorted arrays with the maps' own lower_bound / insert steps.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "orb_oracle.h"

/* FORB::distance, FORB.cpp:79-99 (the bit-twiddling popcount there counts set bits of the XOR) */
static int bow_dist(const uint8_t *a, const uint8_t *b) {
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

/* transform(feature, word_id, weight, nid, levelsup), TemplatedVocabulary.h:1216-1260.  *nid is left at 0 when the
 * reference would leave it unset (leaf above nid_level). */
void orb_oracle_bow_descend(const uint8_t *node_desc, const int32_t *child_ptr, const int32_t *children, int depth_L,
                            const uint8_t *desc, int n, int levelsup, int32_t *leaf_out, int32_t *node_out) {
    const int nid_level = depth_L - levelsup;
    for (int f = 0; f < n; f++) {
        const uint8_t *feature = desc + (size_t)f * 32;
        int nid = 0; /* root when nid_level <= 0 (:1227) */
        int final_id = 0, current_level = 0;
        do {
            ++current_level;
            const int32_t *nodes = children + child_ptr[final_id];
            const int nn = child_ptr[final_id + 1] - child_ptr[final_id];
            final_id = nodes[0];
            double best_d = (double)bow_dist(feature, node_desc + (size_t)final_id * 32);
            for (int k = 1; k < nn; k++) {
                const int id = nodes[k];
                const double d = (double)bow_dist(feature, node_desc + (size_t)id * 32);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (child_ptr[final_id + 1] > child_ptr[final_id]); /* !isLeaf() */
        leaf_out[f] = final_id;
        node_out[f] = nid;
    }
}

/* transform(features, BowVector&, FeatureVector&, levelsup), TemplatedVocabulary.h:1126-1196.
 * weighting: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY; norm: 0 = scoring object does not normalise, 1 = L1, 2 = L2.
 * Outputs as in include/orbfe_bow.h. */
void orb_oracle_bow_transform(const uint8_t *node_desc, const int32_t *child_ptr, const int32_t *children, const int32_t *word_id,
                              const double *weight, int depth_L, int weighting, int norm, const uint8_t *desc, int n,
                              int levelsup, int *nwords_out, int32_t *bow_ids, double *bow_vals, int *nnodes_out,
                              int32_t *fv_ids, int32_t *fv_ptr, int32_t *fv_feats) {
    int32_t *leaf = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t *node = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    orb_oracle_bow_descend(node_desc, child_ptr, children, depth_L, desc, n, levelsup, leaf, node);
    int nw = 0; /* BowVector as a sorted array of (id, value) */
    /* FeatureVector as a sorted array of node ids, each with a growing list */
    int nn = 0;
    int32_t **lists = (int32_t **)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t *));
    int *lens = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    for (int i = 0; i < n; i++) {
        const int32_t id = word_id[leaf[i]];
        const double w = weight[leaf[i]];
        if (!(w > 0)) continue; /* stopped word */
        /* v.addWeight(id, w) / v.addIfNotExist(id, w): lower_bound, then += or insert */
        int pos = 0;
        while (pos < nw && bow_ids[pos] < id) pos++;
        if (pos < nw && bow_ids[pos] == id) {
            if (weighting == 0 || weighting == 1) bow_vals[pos] += w;
        } else {
            memmove(bow_ids + pos + 1, bow_ids + pos, sizeof(int32_t) * (size_t)(nw - pos));
            memmove(bow_vals + pos + 1, bow_vals + pos, sizeof(double) * (size_t)(nw - pos));
            bow_ids[pos] = id;
            bow_vals[pos] = w;
            nw++;
        }
        /* fv.addFeature(nid, i_feature) */
        const int32_t nid = node[i];
        int p = 0;
        while (p < nn && fv_ids[p] < nid) p++;
        if (!(p < nn && fv_ids[p] == nid)) {
            memmove(fv_ids + p + 1, fv_ids + p, sizeof(int32_t) * (size_t)(nn - p));
            memmove(lists + p + 1, lists + p, sizeof(int32_t *) * (size_t)(nn - p));
            memmove(lens + p + 1, lens + p, sizeof(int) * (size_t)(nn - p));
            fv_ids[p] = nid;
            lists[p] = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
            lens[p] = 0;
            nn++;
        }
        lists[p][lens[p]++] = i;
    }
    const int must = norm != 0;
    if ((weighting == 0 || weighting == 1) && nw > 0 && !must) { /* :1166-1172 */
        const double nd = (double)nw;
        for (int k = 0; k < nw; k++) bow_vals[k] /= nd;
    }
    if (must) { /* BowVector::normalize, BowVector.cpp:62-84 */
        double s = 0.0;
        if (norm == 1) {
            for (int k = 0; k < nw; k++) s += fabs(bow_vals[k]);
        } else {
            for (int k = 0; k < nw; k++) s += bow_vals[k] * bow_vals[k];
            s = sqrt(s);
        }
        if (s > 0.0)
            for (int k = 0; k < nw; k++) bow_vals[k] /= s;
    }
    *nwords_out = nw;
    fv_ptr[0] = 0;
    for (int p = 0; p < nn; p++) {
        memcpy(fv_feats + fv_ptr[p], lists[p], sizeof(int32_t) * (size_t)lens[p]);
        fv_ptr[p + 1] = fv_ptr[p] + lens[p];
        free(lists[p]);
    }
    *nnodes_out = nn;
    free(lists); free(lens); free(leaf); free(node);
}

static int cmp_int(const void *a, const void *b) { return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b); }

/* MapPoint::ComputeDistinctiveDescriptors, MapPoint.cc:213-243, for ngroups independent map points */
void orb_oracle_distinctive(const uint8_t *desc, const int32_t *group_ptr, int ngroups, int32_t *best_out) {
    for (int g = 0; g < ngroups; g++) {
        const int b = group_ptr[g], N = group_ptr[g + 1] - b;
        if (N <= 0) { best_out[g] = -1; continue; }
        float *D = (float *)malloc(sizeof(float) * (size_t)N * N);
        for (int i = 0; i < N; i++) {
            D[(size_t)i * N + i] = 0;
            for (int j = i + 1; j < N; j++) {
                const int dij = bow_dist(desc + (size_t)(b + i) * 32, desc + (size_t)(b + j) * 32);
                D[(size_t)i * N + j] = (float)dij;
                D[(size_t)j * N + i] = (float)dij;
            }
        }
        int BestMedian = 0x7FFFFFFF, BestIdx = 0;
        int *v = (int *)malloc(sizeof(int) * (size_t)N);
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < N; j++) v[j] = (int)D[(size_t)i * N + j]; /* vector<int> from float row */
            qsort(v, (size_t)N, sizeof(int), cmp_int);
            const int median = v[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best_out[g] = BestIdx;
        free(v); free(D);
    }
}
