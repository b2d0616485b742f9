/* orbfe_bow.h -- C-ABI of the two "next" rows of SURVEY.md section 8(f) that reuse the 256-bit Hamming primitive:
 *
 *   N2  DBoW2 vocabulary-tree transform: descriptors -> BowVector + FeatureVector
 *       (reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1262, FORB.cpp:79-99 (distance),
 *        BowVector.cpp:34-84, FeatureVector.cpp:32-48; callers Frame.cc:280-287, KeyFrame.cc:56-65)
 *   N3  KeyFrameDatabase::DetectLoopCandidates / DetectRelocalisationCandidates on arrays
 *       (reference src/KeyFrameDatabase.cc:73-308, L1Scoring::score Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-67)
 *   N4  MapPoint::ComputeDistinctiveDescriptors, batched over map points (reference src/MapPoint.cc:185-250)
 *
 * Plain pointers and sizes; host arrays unless a parameter is named d_*.  Return values: OrbfeStatus (orbfe.h).
 */
#ifndef ORBFE_BOW_H
#define ORBFE_BOW_H

#include <stdint.h>

#include "orbfe.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrbfeVocabulary OrbfeVocabulary;

/* DBoW2::WeightingType / the LNorm a scoring object asks for (ORB-SLAM: TF_IDF + L1, ORBVocabulary.h). */
enum { ORBFE_BOW_TF_IDF = 0, ORBFE_BOW_TF = 1, ORBFE_BOW_IDF = 2, ORBFE_BOW_BINARY = 3 };
enum { ORBFE_BOW_NORM_NONE = 0, ORBFE_BOW_NORM_L1 = 1, ORBFE_BOW_NORM_L2 = 2 };

/* The vocabulary tree as flat arrays (what TemplatedVocabulary::m_nodes holds after load()):
 *   node 0 is the root; the children of node i are children[child_ptr[i] .. child_ptr[i+1]) in the order of
 *   m_nodes[i].children (the descent keeps the FIRST child of minimum distance); a node without children is a leaf
 *   (a word) with word_id[i] >= 0 and weight[i]; node_desc holds 32 bytes per node (the root's are ignored).
 * depth_L = m_L.  The arrays are copied; the node table lives in device memory of `device`. */
OrbfeVocabulary *orbfe_vocabulary_create(int device, int nnodes, int depth_L, const uint8_t *node_desc, const int32_t *child_ptr,
                                         const int32_t *children, const int32_t *word_id, const double *weight, int weighting,
                                         int norm);
void orbfe_vocabulary_destroy(OrbfeVocabulary *v);

/* transform(feature, word_id, weight, &nid, levelsup) for n descriptors (TemplatedVocabulary.h:1216-1260):
 * leaf_out[i] = node id of the leaf reached, node_out[i] = id of the ancestor at level depth_L - levelsup (0 = root when
 * that level is <= 0, and also when the leaf is shallower than that level, where the reference leaves *nid unset).
 * Device-pointer form: enqueued on `stream` (NULL = the vocabulary's stream), not synchronised -- it chains directly
 * after orbfe_extract_batch_device on the descriptors that call produced. */
int orbfe_bow_descend_device(OrbfeVocabulary *v, const uint8_t *d_desc, int n, int levelsup, int32_t *d_leaf_out,
                             int32_t *d_node_out, void *stream);
int orbfe_bow_descend(OrbfeVocabulary *v, const uint8_t *desc, int n, int levelsup, int32_t *leaf_out, int32_t *node_out);

/* transform(features, BowVector&, FeatureVector&, levelsup) (TemplatedVocabulary.h:1126-1196): the descent on the
 * device, the two std::map builds restated on sorted arrays on the host.
 *   BowVector:     *nwords_out entries (word id ascending) in bow_ids / bow_vals            (capacity n each)
 *   FeatureVector: *nnodes_out entries (node id ascending) in fv_ids, rows fv_ptr[k]..fv_ptr[k+1] of fv_feats
 *                  (feature indices ascending; capacities n, n+1, n) -- the CSR form orbfe_search_by_bow takes. */
int orbfe_bow_transform(OrbfeVocabulary *v, const uint8_t *desc, int n, int levelsup, int *nwords_out, int32_t *bow_ids,
                        double *bow_vals, int *nnodes_out, int32_t *fv_ids, int32_t *fv_ptr, int32_t *fv_feats);

/* MapPoint::ComputeDistinctiveDescriptors for ngroups map points at once: group g owns the descriptors
 * desc[group_ptr[g] .. group_ptr[g+1]) (its observations, in the order of the reference's vDescriptors);
 * best_out[g] = index inside the group of the descriptor with the least median distance to the group
 * (median = sorted[(int)(0.5*(N-1))], first minimum wins, MapPoint.cc:228-243), -1 for an empty group.
 * All N x N distances and the medians are computed on the device (one warp per map point). */
int orbfe_distinctive_descriptors(OrbfeMatcher *m, const uint8_t *desc, const int32_t *group_ptr, int ngroups, int32_t *best_out);

/* KeyFrameDatabase::DetectLoopCandidates (mode 0, KeyFrameDatabase.cc:73-195) / DetectRelocalisationCandidates (mode 1,
 * :197-308) with the database as arrays.  Keyframe k (k = 0..nkf-1, in the order the keyframes were add()ed: that is
 * the order inside every inverted-file list) owns the BowVector db_ids/db_vals[kf_ptr[k] .. kf_ptr[k+1]) (word ids
 * ascending); the query BowVector is q_ids/q_vals (ascending).  connected[k] != 0 marks the query keyframe's
 * GetConnectedKeyFrames() (mode 0 only, may be NULL); covis/covis_ptr hold GetBestCovisibilityKeyFrames(10) of every
 * keyframe in the order that call returns them.  min_score: the minScore argument (mode 0 only).
 * One device thread per keyframe walks the two sorted word lists (shared-word count, first shared word, and the L1
 * score accumulated in ascending word order exactly as L1Scoring::score does); thresholds, the covisibility
 * accumulation and the candidate list are replayed on the host in the reference's list order.
 * cand_out (capacity nkf) receives *ncand_out keyframe indices in the order of the returned vector.  common_out /
 * score_out (capacity nkf each, may be NULL) receive mnLoopWords / the float score of every keyframe that shares a word
 * (score = -1 where the reference does not compute one). */
int orbfe_bow_db_detect(OrbfeMatcher *m, int mode, int nq, const int32_t *q_ids, const double *q_vals, int nkf,
                        const int32_t *kf_ptr, const int32_t *db_ids, const double *db_vals, const uint8_t *connected,
                        const int32_t *covis_ptr, const int32_t *covis, float min_score, int *ncand_out, int32_t *cand_out,
                        int32_t *common_out, float *score_out);

#ifdef __cplusplus
}
#endif
#endif /* ORBFE_BOW_H */
