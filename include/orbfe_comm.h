/*
 * orbfe_comm.h -- multi-GPU entry points of liborbfe.so (one process per GPU, NCCL over NVLink 5 / NVSwitch).
 *
 * The reference is single-camera and has no communication layer at all; these calls are what SURVEY.md 8(b)/(e) asks the
 * replacement to export for the two BASELINE configurations that exchange data between GPUs:
 *   config 4  8-camera rig, one camera per GPU: every GPU extracts its own camera's frame (ORBextractor::operator(),
 *             src/ORBextractor.cc:718-779), then needs every other camera's keypoints + 32-byte descriptors for cross-camera
 *             ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:598-713; call pattern Tracking.cc:497);
 *   config 5  loop-closure brute force: one query (2000 descriptors) against a keyframe database row-sharded over the GPUs
 *             (ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1794-1810, per pair; consumers LoopClosing.cc:233-370).
 *
 * Two ways to do the config-4 exchange:
 *   orbfe_allgather_desc          plain ncclAllGather of the extractor's device output blocks (no host hop);
 *   OrbfeRigExchange              the exchange FUSED into the extractor: the descriptor kernel stores every keypoint and
 *                                 descriptor straight into each peer's gather buffer over NVLink (peer pointers obtained with
 *                                 CUDA IPC), the last thread block of the kernel publishes an epoch flag to every peer, and
 *                                 the consumer side is one tiny wait kernel -- no collective launch, no extra pass over the data.
 *
 * NCCL is loaded at run time (dlopen "libnccl.so.2": the copy already in the process if the host program has one, e.g.
 * torch's, else the system one), so liborbfe.so has no link-time dependency on it and single-GPU users never need it.
 * Conventions as in orbfe.h: 0 = ok, negative OrbfeStatus, orbfe_last_error() has the text; device pointers + a
 * cudaStream_t (NULL = the communicator's own stream); nothing is synchronised unless stated.
 */
#ifndef ORBFE_COMM_H
#define ORBFE_COMM_H

#include "orbfe.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORBFE_COMM_ID_BYTES 128   /* sizeof(ncclUniqueId) */
#define ORBFE_MAX_RANKS 16

typedef struct OrbfeComm OrbfeComm;
typedef struct OrbfeRigExchange OrbfeRigExchange;

/* rank 0 creates the id (ncclGetUniqueId) and ships it to the other ranks by any out-of-band means (MPI, a file, torch.distributed) */
int orbfe_comm_unique_id(uint8_t id[ORBFE_COMM_ID_BYTES]);
/* ncclCommInitRank on `device`; collective over all `world` ranks */
int orbfe_comm_create(const uint8_t id[ORBFE_COMM_ID_BYTES], int world, int rank, int device, OrbfeComm **out);
int orbfe_comm_destroy(OrbfeComm *c);
int orbfe_comm_world(const OrbfeComm *c);
int orbfe_comm_rank(const OrbfeComm *c);
/* version of the NCCL library actually loaded (e.g. 22809), 0 if none */
int orbfe_comm_nccl_version(void);
int orbfe_comm_sync(OrbfeComm *c);                       /* cudaStreamSynchronize of the communicator's own stream */
int orbfe_comm_barrier(OrbfeComm *c, void *stream);      /* a 4-byte all-reduce */

/* config 4, plain: all-gather of `nslots` frames per rank.  d_kps: nslots x cap keypoints, d_desc: nslots x cap x 32 bytes,
 * d_counts: nslots ints (what orbfe_extract_batch_device wrote).  Outputs: world x nslots x cap (...), rank-major.
 * Three ncclAllGather calls in one group (NCCL fuses them into one launch). */
int orbfe_allgather_desc(OrbfeComm *c, const OrbfeKeyPoint *d_kps, const uint8_t *d_desc, const int *d_counts, int cap, int nslots,
                         OrbfeKeyPoint *d_all_kps, uint8_t *d_all_desc, int *d_all_counts, void *stream);
/* config 5 building blocks: broadcast `bytes` from `root`; all-gather of `bytes_per_rank` from every rank */
int orbfe_comm_broadcast(OrbfeComm *c, void *d_buf, size_t bytes, int root, void *stream);
int orbfe_comm_allgather(OrbfeComm *c, const void *d_send, void *d_recv, size_t bytes_per_rank, void *stream);

/* config 5: the sharded sweep.  The database is row-sharded by keyframe: this rank holds `ngroups_local` groups of
 * `group_size` descriptors (orbfe_shard_range gives the split).  The query (nq x 32 bytes, valid on rank `root`) is
 * broadcast, every rank sweeps its shard (knn2 kernel: best / second-best distance and best index per keyframe and
 * query), and the per-keyframe results are all-gathered so that every rank ends up with `ngroups_total` x nq results
 * in global keyframe order.  Shards are padded to ceil(ngroups_total / world) groups for the collective; d_scratch must
 * hold (world + 1) * ceil(ngroups_total / world) * nq * 8 bytes (unused when world == 1). */
int orbfe_shard_range(int n_items, int world, int rank, int *lo, int *hi);
int orbfe_knn2_sweep_sharded(OrbfeComm *c, OrbfeMatcher *m, uint8_t *d_query, int nq, int root, const uint8_t *d_db_shard,
                             int ngroups_total, int group_size, uint16_t *d_best_all, int32_t *d_best_idx_all,
                             uint16_t *d_second_all, void *d_scratch, void *stream);

/* ---- config 4, fused: extractor -> peers ------------------------------------------------------------------------
 * Every rank creates one exchange object (collective): it allocates the local gather buffers (world x nslots x cap
 * keypoints / descriptors / counts, double-buffered), exchanges CUDA IPC handles through the communicator and opens the
 * peers' buffers.  Then, per exchange:
 *   orbfe_extract_batch_device_exchange(ex, ..., x, stream)   extract `nslots` frames; the descriptor kernel writes its
 *                                                           outputs into slot [rank] of EVERY rank's buffer and its last
 *                                                           thread block publishes the epoch to every rank;
 *   orbfe_rig_exchange_wait(x, stream)                      enqueue the wait for all ranks' data of this epoch;
 *   ... consume orbfe_rig_exchange_buffers() on `stream` (matchers) ...
 *   orbfe_rig_exchange_release(x, stream)                   tell the peers this rank is done reading the epoch (a buffer
 *                                                           half is only overwritten after every rank released it).
 * Waits give up after about two seconds and raise the device error flag read by orbfe_rig_exchange_check(). */
int orbfe_rig_exchange_create(OrbfeComm *c, int cap, int nslots, OrbfeRigExchange **out);
int orbfe_rig_exchange_destroy(OrbfeRigExchange *x);
int orbfe_extract_batch_device_exchange(OrbfeExtractor *ex, const uint8_t *d_imgs, int width, int height, size_t stride,
                                        size_t frame_stride, int batch, OrbfeRigExchange *x, void *stream);
/* config 4's consumer, fused as well: orbfe_search_for_initialization_device (include/orbfe_match.h) on the gathered arrays
 * of the epoch just produced; the matcher kernel polls the local epoch flags before its first read and its last thread
 * block publishes the release -- neither orbfe_rig_exchange_wait nor _release is needed around it. */
int orbfe_search_for_initialization_exchange(OrbfeMatcher *m, OrbfeRigExchange *x, int npairs, const int *d_f1_idx, const int *d_f2_idx,
                                             float *d_prev_matched, float min_x, float min_y, float max_x, float max_y, int window,
                                             float nnratio, int check_orientation, int *d_match12, int *d_nmatches, void *stream);
int orbfe_rig_exchange_wait(OrbfeRigExchange *x, void *stream);
int orbfe_rig_exchange_release(OrbfeRigExchange *x, void *stream);
/* gathered views of the epoch last waited for: world x nslots x cap keypoints / descriptors, world x nslots counts */
int orbfe_rig_exchange_buffers(OrbfeRigExchange *x, OrbfeKeyPoint **d_all_kps, uint8_t **d_all_desc, int **d_all_counts);
/* the same views for the epoch just PRODUCED by orbfe_extract_batch_device_exchange (before any wait): only this rank's own
 * slot [rank] is valid there without waiting (it was written locally, in stream order) */
int orbfe_rig_exchange_buffers_produced(OrbfeRigExchange *x, OrbfeKeyPoint **d_all_kps, uint8_t **d_all_desc, int **d_all_counts);
/* synchronises `stream` and reports a timed-out wait (ORBFE_ERR_INTERNAL) */
int orbfe_rig_exchange_check(OrbfeRigExchange *x, void *stream);
/* bytes this rank pushed to its peers over NVLink in the last exchange (capacity-based upper bound: nslots x cap x 60 x (world-1)) */
size_t orbfe_rig_exchange_bytes(const OrbfeRigExchange *x);

#ifdef __cplusplus
}
#endif
#endif /* ORBFE_COMM_H */
