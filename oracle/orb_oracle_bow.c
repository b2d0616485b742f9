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

/* L1Scoring::score, ScoringObject.cpp:23-67 (with the lower_bound jumps restated as linear advances to the first
 * element >= the other side's id: the same element) */
static double l1_score(const int32_t *i1, const double *v1, int n1, const int32_t *i2, const double *v2, int n2) {
    int a = 0, b = 0;
    double score = 0;
    while (a != n1 && b != n2) {
        const double vi = v1[a], wi = v2[b];
        if (i1[a] == i2[b]) {
            score += fabs(vi - wi) - fabs(vi) - fabs(wi);
            ++a; ++b;
        } else if (i1[a] < i2[b]) {
            while (a != n1 && i1[a] < i2[b]) ++a; /* v1.lower_bound(v2_it->first) */
        } else {
            while (b != n2 && i2[b] < i1[a]) ++b;
        }
    }
    score = -score / 2.0;
    return score;
}

/* KeyFrameDatabase::DetectLoopCandidates (mode 0, KeyFrameDatabase.cc:73-195) and DetectRelocalisationCandidates
 * (mode 1, :197-308) on arrays, written with a real inverted file (one list per word, keyframes appended in index
 * order = add() order, :40-46) and the reference's per-keyframe bookkeeping fields.  A keyframe's mRelocScore that
 * the reference would read without having written it in this query counts as 0.  Returns the number of candidates. */
int orb_oracle_bow_db_detect(int mode, int nq, const int32_t *q_ids, const double *q_vals, int nkf, const int32_t *kf_ptr,
                             const int32_t *db_ids, const double *db_vals, const uint8_t *connected, const int32_t *covis_ptr,
                             const int32_t *covis, float minScore, int32_t *cand_out, int32_t *common_out, float *score_out) {
    int nwords = 0;
    for (int k = 0; k < kf_ptr[nkf]; k++) if (db_ids[k] + 1 > nwords) nwords = db_ids[k] + 1;
    for (int k = 0; k < nq; k++) if (q_ids[k] + 1 > nwords) nwords = q_ids[k] + 1;
    /* mvInvertedFile: CSR built by add()ing keyframes 0..nkf-1 in order */
    int *inv_ptr = (int *)calloc((size_t)nwords + 2, sizeof(int));
    for (int k = 0; k < kf_ptr[nkf]; k++) inv_ptr[db_ids[k] + 1]++;
    for (int w = 0; w < nwords; w++) inv_ptr[w + 1] += inv_ptr[w];
    int *inv = (int *)malloc(sizeof(int) * (size_t)(kf_ptr[nkf] > 0 ? kf_ptr[nkf] : 1));
    int *fill = (int *)calloc((size_t)nwords + 1, sizeof(int));
    for (int kf = 0; kf < nkf; kf++)
        for (int k = kf_ptr[kf]; k < kf_ptr[kf + 1]; k++) inv[inv_ptr[db_ids[k]] + fill[db_ids[k]]++] = kf;
    free(fill);

    int *query = (int *)calloc((size_t)nkf, sizeof(int));   /* mnLoopQuery == this query */
    int *words = (int *)calloc((size_t)nkf, sizeof(int));   /* mnLoopWords / mnRelocWords */
    float *sc = (float *)calloc((size_t)nkf, sizeof(float)); /* mLoopScore / mRelocScore */
    int *sharing = (int *)malloc(sizeof(int) * (size_t)(nkf > 0 ? nkf : 1));
    int nsharing = 0;
    for (int qi = 0; qi < nq; qi++) {
        const int w = q_ids[qi];
        for (int p = inv_ptr[w]; p < inv_ptr[w + 1]; p++) {
            const int kfi = inv[p];
            if (!query[kfi]) {
                words[kfi] = 0;
                if (mode == 1 || !(connected && connected[kfi])) {
                    query[kfi] = 1;
                    sharing[nsharing++] = kfi;
                }
            }
            words[kfi]++;
        }
    }
    if (common_out) memcpy(common_out, words, sizeof(int) * (size_t)nkf);
    if (score_out) for (int k = 0; k < nkf; k++) score_out[k] = -1.0f;
    int ncand = 0;
    if (nsharing > 0) {
        int maxCommonWords = 0;
        for (int s = 0; s < nsharing; s++) if (words[sharing[s]] > maxCommonWords) maxCommonWords = words[sharing[s]];
        const int minCommonWords = maxCommonWords * 0.8f;
        float *lsc = (float *)malloc(sizeof(float) * (size_t)nsharing);
        int *lkf = (int *)malloc(sizeof(int) * (size_t)nsharing);
        int nl = 0;
        for (int s = 0; s < nsharing; s++) {
            const int kfi = sharing[s];
            if (words[kfi] > minCommonWords) {
                const float si = (float)l1_score(q_ids, q_vals, nq, db_ids + kf_ptr[kfi], db_vals + kf_ptr[kfi], kf_ptr[kfi + 1] - kf_ptr[kfi]);
                sc[kfi] = si;
                if (score_out) score_out[kfi] = si;
                if (mode == 1 || si >= minScore) { lsc[nl] = si; lkf[nl] = kfi; nl++; }
            }
        }
        if (nl > 0) {
            float *acc = (float *)malloc(sizeof(float) * (size_t)nl);
            int *bestkf = (int *)malloc(sizeof(int) * (size_t)nl);
            float bestAccScore = mode == 0 ? minScore : 0;
            for (int e = 0; e < nl; e++) {
                const int kfi = lkf[e];
                float bestScore = lsc[e], accScore = lsc[e];
                int pBest = kfi;
                for (int c = covis_ptr[kfi]; c < covis_ptr[kfi + 1]; c++) {
                    const int kf2 = covis[c];
                    if (mode == 0) {
                        if (query[kf2] && words[kf2] > minCommonWords) {
                            accScore += sc[kf2];
                            if (sc[kf2] > bestScore) { pBest = kf2; bestScore = sc[kf2]; }
                        }
                    } else {
                        if (!query[kf2]) continue;
                        accScore += sc[kf2];
                        if (sc[kf2] > bestScore) { pBest = kf2; bestScore = sc[kf2]; }
                    }
                }
                acc[e] = accScore;
                bestkf[e] = pBest;
                if (accScore > bestAccScore) bestAccScore = accScore;
            }
            const float minScoreToRetain = 0.75f * bestAccScore;
            uint8_t *added = (uint8_t *)calloc((size_t)nkf, 1);
            for (int e = 0; e < nl; e++) {
                if (acc[e] > minScoreToRetain && !added[bestkf[e]]) {
                    cand_out[ncand++] = bestkf[e];
                    added[bestkf[e]] = 1;
                }
            }
            free(added); free(acc); free(bestkf);
        }
        free(lsc); free(lkf);
    }
    free(inv_ptr); free(inv); free(query); free(words); free(sc); free(sharing);
    return ncand;
}
