"""GPU parity of the SURVEY section 8(f) rows N2 (DBoW2 vocabulary transform) and N4 (batched
MapPoint::ComputeDistinctiveDescriptors): CUDA path through the C-ABI vs the oracle, bit-exact (ids and the float64
BowVector values, which are produced by the same additions in the same order)."""
import numpy as np
import pytest

import oracle as O
import orb_slam_b200 as fe
from orb_slam_b200 import bow as B
from orb_slam_b200.synth import random_descriptors, noisy_copies, random_vocabulary, random_keyframe_db

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,L,ragged,n", [(10, 4, False, 2000), (7, 3, True, 777), (40, 2, False, 500), (10, 5, True, 1500)])
def test_bow_transform_matches_oracle(gpu_required, k, L, ragged, n):
    voc = random_vocabulary(k, L, seed=k + L, ragged=ragged)
    words = voc["node_desc"][voc["word_id"] >= 0]
    rng = np.random.default_rng(5)
    desc = noisy_copies(words[rng.integers(0, len(words), n)], 0.1, 9)
    desc[::50] = random_descriptors(len(desc[::50]), 4)  # a few unrelated descriptors
    for levelsup, weighting, norm in [(4, B.TF_IDF, B.NORM_L1), (2, B.TF, B.NORM_NONE), (0, B.IDF, B.NORM_L2), (L + 1, B.BINARY, B.NORM_L1)]:
        v = B.Vocabulary(voc, weighting, norm)
        leaf, node = v.descend(desc, levelsup)
        leaf_o, node_o = O.bow_descend(voc, desc, levelsup)
        assert np.array_equal(leaf, leaf_o) and np.array_equal(node, node_o)
        (ids, vals), (fids, fptr, ffeat) = v.transform(desc, levelsup)
        (ids_o, vals_o), (fids_o, fptr_o, ffeat_o) = O.bow_transform(voc, desc, levelsup, weighting, norm)
        assert np.array_equal(ids, ids_o) and np.array_equal(vals.view(np.uint64), vals_o.view(np.uint64))
        assert np.array_equal(fids, fids_o) and np.array_equal(fptr, fptr_o) and np.array_equal(ffeat, ffeat_o)
        assert len(ids) > 10
        v.close()


def test_bow_descend_device_chains_after_extraction(gpu_required):
    """Device-pointer form on the descriptors orbfe_extract_batch_device leaves in HBM (no host round trip)."""
    import torch
    from orb_slam_b200.synth import textured_frame
    voc = random_vocabulary(10, 4, seed=2)
    v = B.Vocabulary(voc)
    ex = fe.ORBextractor(1000, 1.2, 8)
    img = textured_frame(640, 480, seed=3)
    kps, desc = ex(img)
    dev = torch.device("cuda", 0)
    d_desc = torch.from_numpy(np.ascontiguousarray(desc)).to(dev)
    d_leaf = torch.zeros(len(desc), dtype=torch.int32, device=dev)
    d_node = torch.zeros(len(desc), dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    v.descend_device(d_desc.data_ptr(), len(desc), 4, d_leaf.data_ptr(), d_node.data_ptr(), s.cuda_stream)
    s.synchronize()
    leaf_o, node_o = O.bow_descend(voc, desc, 4)
    assert np.array_equal(d_leaf.cpu().numpy(), leaf_o) and np.array_equal(d_node.cpu().numpy(), node_o)
    ex.close()
    v.close()


def test_empty_and_error_cases(gpu_required):
    voc = random_vocabulary(3, 2, seed=0)
    v = B.Vocabulary(voc)
    (ids, vals), (fids, fptr, ffeat) = v.transform(np.zeros((0, 32), np.uint8))
    assert len(ids) == 0 and len(fids) == 0 and list(fptr) == [0]
    bad = dict(voc)
    bad["children"] = voc["children"].copy()
    bad["children"][0] = len(voc["word_id"]) + 5
    with pytest.raises(fe.OrbfeError):
        B.Vocabulary(bad)
    v.close()


def test_distinctive_descriptors_matches_oracle(gpu_required):
    rng = np.random.default_rng(12)
    groups, ptr = [], [0]
    for g in range(700):
        n = int(rng.integers(1, 60)) if g % 11 else int(rng.integers(0, 3))
        if g == 5:
            n = 400  # a long-lived map point
        base = random_descriptors(1, 1000 + g)
        groups.append(noisy_copies(np.repeat(base, n, axis=0), rng.uniform(0.01, 0.3), 5000 + g) if n else np.zeros((0, 32), np.uint8))
        ptr.append(ptr[-1] + n)
    desc = np.concatenate(groups)
    ptr = np.array(ptr, np.int32)
    m = fe.ORBmatcher(0.6, True)
    best = B.distinctive_descriptors(m, desc, ptr)
    best_o = O.distinctive_descriptors(desc, ptr)
    assert np.array_equal(best, best_o)
    assert (best >= 0).sum() > 600
    # empty batch
    assert len(B.distinctive_descriptors(m, np.zeros((0, 32), np.uint8), np.zeros(1, np.int32))) == 0
    m.close()


def test_keyframe_db_detect_matches_oracle(gpu_required):
    """N3: DetectLoopCandidates / DetectRelocalisationCandidates on arrays, CUDA scoring + host list logic vs the oracle's
    literal inverted-file walk: candidates in the same order, identical shared-word counts and float scores."""
    m = fe.ORBmatcher(0.75, True)
    for seed, nkf in ((0, 400), (1, 1500), (2, 60)):
        db = random_keyframe_db(nkf=nkf, nwords=8000, words_per_kf=300, seed=seed, loop_at=nkf // 5)
        for mode, ms in ((0, 0.0), (0, 0.03), (1, 0.0)):
            args = (db["q_ids"], db["q_vals"], db["kf_ptr"], db["db_ids"], db["db_vals"], db["connected"], db["covis_ptr"], db["covis"], ms)
            cand, common, score = B.db_detect(m, mode, *args)
            cand_o, common_o, score_o = O.bow_db_detect(mode, *args)
            assert np.array_equal(cand, cand_o), (seed, mode, ms)
            assert np.array_equal(common, common_o)
            assert np.array_equal(score.view(np.uint32), score_o.view(np.uint32))
            assert len(cand) > 0 and (nkf // 5 in cand or nkf // 5 + 1 in cand or nkf // 5 - 1 in cand)
    # a query that shares nothing
    db = random_keyframe_db(nkf=50, nwords=2000, words_per_kf=100, seed=5)
    cand, common, score = B.db_detect(m, 1, np.array([2001, 2002], np.int32), np.array([0.5, 0.5]), db["kf_ptr"], db["db_ids"], db["db_vals"],
                                      db["connected"], db["covis_ptr"], db["covis"], 0.0)
    assert len(cand) == 0 and common.sum() == 0
    m.close()
