"""GPU parity of the windowed matchers (host candidate lists + device Hamming + host greedy replay, all through
the C-ABI of liborbfe.so) against the oracle's restatement of ORBmatcher.cc / Frame.cc."""
import numpy as np
import pytest

import oracle as O
import orb_slam_b200 as fe
from orb_slam_b200 import matching as M
from orb_slam_b200.synth import textured_frame, shifted_frame

pytestmark = pytest.mark.gpu

W, H = 640, 480
FX = FY = 500.0
CX, CY, DEPTH = W / 2.0, H / 2.0, 4.0


def _features(n_frames=3):
    ex = fe.ORBextractor(1000, 1.2, 8)
    base = textured_frame(W, H, seed=21)
    frames, shifts = [base], [(0, 0)]
    rng = np.random.default_rng(1)
    for i in range(1, n_frames):
        dx, dy = int(rng.integers(-5, 6)), int(rng.integers(-4, 5))
        frames.append(shifted_frame(frames[-1], dx, dy, seed=i))
        shifts.append((dx, dy))
    feats = [ex(f) for f in frames]
    ex.close()
    return feats, shifts


def _tcw(dx, dy):
    T = np.zeros((3, 4), np.float32)
    T[0, 0] = T[1, 1] = T[2, 2] = 1
    T[0, 3], T[1, 3] = dx * DEPTH / FX, dy * DEPTH / FY
    return T


def _world(k):
    w = np.empty((len(k), 3), np.float32)
    w[:, 0] = (k["x"] - np.float32(CX)) / np.float32(FX) * np.float32(DEPTH)
    w[:, 1] = (k["y"] - np.float32(CY)) / np.float32(FY) * np.float32(DEPTH)
    w[:, 2] = DEPTH
    return w


def test_search_by_projection_pairs(gpu_required):
    feats, shifts = _features(4)
    m = fe.ORBmatcher(0.9, True)
    rng = np.random.default_rng(3)
    curs, lasts, has, outl, world, T, pre = [], [], [], [], [], [], []
    for i in range(1, 4):
        kc, dc = feats[i]
        kl, dl = feats[i - 1]
        curs.append(M.FrameView(kc, dc, W, H))
        lasts.append(M.FrameView(kl, dl, W, H))
        has.append((rng.random(len(kl)) < 0.9).astype(np.uint8))
        outl.append((rng.random(len(kl)) < 0.05).astype(np.uint8))
        world.append(_world(kl))
        T.append(_tcw(*shifts[i]))
        occ = np.full(len(kc), -1, np.int32)
        occ[rng.random(len(kc)) < 0.03] = 5  # a few already-occupied slots
        pre.append(occ)
    nm, mp = M.search_by_projection_frames(m, curs, lasts, has, outl, world, T, FX, FY, CX, CY, 15.0, cur_mp=pre)
    total = 0
    for j in range(3):
        fc = O.OracleFrame(curs[j].kps, curs[j].desc, W, H)
        fl = O.OracleFrame(lasts[j].kps, lasts[j].desc, W, H)
        n_o, mp_o = O.search_by_projection_ff(fc, fl, has[j], outl[j], world[j], T[j], FX, FY, CX, CY, 15.0, True, cur_mp=pre[j])
        assert nm[j] == n_o
        assert np.array_equal(mp[j], mp_o)
        total += n_o
    assert total > 300  # the synthetic stream really has true matches
    m.close()


def test_window_search_and_initialization(gpu_required):
    feats, shifts = _features(2)
    (k1, d1), (k2, d2) = feats
    f1, f2 = M.FrameView(k1, d1, W, H), M.FrameView(k2, d2, W, H)
    o1, o2 = O.OracleFrame(k1, d1, W, H), O.OracleFrame(k2, d2, W, H)
    has = (np.random.default_rng(0).random(len(k1)) < 0.8).astype(np.uint8)
    for nnratio, ori, win, lo, hi in [(0.9, True, 50, -1, 2 ** 31 - 1), (0.6, False, 100, 2, 5), (0.9, True, 200, -1, 2 ** 31 - 1)]:
        m = fe.ORBmatcher(nnratio, ori)
        n, m21 = M.window_search(m, f1, f2, has, win, lo, hi)
        n_o, m21_o = O.window_search(o1, o2, has, win, lo, hi, nnratio=nnratio, check_orientation=ori)
        assert n == n_o and np.array_equal(m21, m21_o)
        m.close()
    m = fe.ORBmatcher(0.9, True)
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
    n, m12, prev_out = M.search_for_initialization(m, f1, f2, prev, 100)
    n_o, m12_o, prev_o = O.search_for_initialization(o1, o2, prev, 100, nnratio=0.9, check_orientation=True)
    assert n == n_o and np.array_equal(m12, m12_o) and np.array_equal(prev_out, prev_o)
    assert n > 20
    # second round with the updated prev-matched positions (Tracking::Initialize calls it repeatedly)
    n2, m12b, _ = M.search_for_initialization(m, f1, f2, prev_out, 100)
    n2_o, m12b_o, _ = O.search_for_initialization(o1, o2, prev_o, 100, nnratio=0.9, check_orientation=True)
    assert n2 == n2_o and np.array_equal(m12b, m12b_o)
    m.close()
