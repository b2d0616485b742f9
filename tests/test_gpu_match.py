"""GPU parity of the Hamming kernels against ORBmatcher::DescriptorDistance as restated by the oracle."""
import numpy as np
import pytest

import oracle as O
import orb_slam_b200 as fe
from orb_slam_b200.synth import random_descriptors, noisy_copies

pytestmark = pytest.mark.gpu


def _np_hamming(a, b):
    return np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(axis=2)


def test_dense_and_csr(gpu_required):
    q = random_descriptors(300, 1)
    t = noisy_copies(random_descriptors(500, 2), 0.1, 3)
    t[:300] = noisy_copies(q, 0.08, 4)
    m = fe.ORBmatcher()
    D = m.hamming_dense(q, t)
    ref = _np_hamming(q, t)
    assert np.array_equal(D, ref)
    for i, j in [(0, 0), (5, 7), (299, 499)]:
        assert D[i, j] == O.hamming(q[i], t[j])
    rng = np.random.default_rng(0)
    rows = [np.sort(rng.choice(500, rng.integers(0, 40), replace=False)) for _ in range(300)]
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    cols = np.concatenate(rows).astype(np.int32)
    out = m.hamming_csr(q, t, row_ptr, cols)
    exp = np.concatenate([ref[i, r] for i, r in enumerate(rows)])
    assert np.array_equal(out, exp)
    # edge cases: identical, complement, empty rows
    z = np.zeros((1, 32), np.uint8)
    assert m.hamming_dense(z, z)[0, 0] == 0
    assert m.hamming_dense(z, ~z)[0, 0] == 256
    assert len(m.hamming_csr(q[:2], t, np.array([0, 0, 0], np.int32), np.zeros(0, np.int32))) == 0
    m.close()


def test_knn2_groups(gpu_required):
    q = random_descriptors(130, 7)
    db = np.concatenate([noisy_copies(q[np.random.default_rng(g).permutation(130)[:100]], 0.1, g) for g in range(9)])
    m = fe.ORBmatcher()
    best, idx, second = m.knn2_groups(q, db, 100)
    for g in range(9):
        bd, bi, sd = O.knn2(q, db[g * 100:(g + 1) * 100])
        assert np.array_equal(best[g], bd) and np.array_equal(idx[g], bi) and np.array_equal(second[g], np.minimum(sd, 65535))
    m.close()
