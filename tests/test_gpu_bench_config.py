"""GPU parity of SearchByProjection(Frame, Frame) (reference src/ORBmatcher.cc:1507-1620) at the BENCHMARKED configuration:
1920x1080, 2000 keypoints per frame, cap 2000, consecutive frames of a stream.  At this size the fused kernel loops
twice over the queries (2000 > 1024 threads) and its candidate entries may live in shared memory or in the per-pair
global scratch (match_kernels.cu `ent = T_total <= smem_entries ? s_ent : scratch`): both placements are forced here,
through the device-pointer entry point and through the host-view entry point, against the oracle."""
import os

import numpy as np
import pytest

import oracle as O
import orb_slam_b200 as fe
from orb_slam_b200 import matching as M
from orb_slam_b200.synth import textured_frame, shifted_frame

pytestmark = pytest.mark.gpu

W, H, NF = 1920, 1080, 2000
FX = FY = 1000.0
CX, CY, DEPTH = W / 2.0, H / 2.0, 5.0
NFRAMES = 9   # 8 consecutive pairs


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


@pytest.fixture(scope="module")
def stream(gpu_required):
    rng = np.random.default_rng(5)
    frames, shifts = [textured_frame(W, H, seed=77)], [(0, 0)]
    for i in range(1, NFRAMES):
        dx, dy = int(rng.integers(-6, 7)), int(rng.integers(-4, 5))
        frames.append(shifted_frame(frames[-1], dx, dy, seed=100 + i))
        shifts.append((dx, dy))
    ex = fe.ORBextractor(NF, 1.2, 8)
    kps, desc, cnt = ex.extract_batch(np.stack(frames))
    ex.close()
    assert list(cnt) == [NF] * NFRAMES
    # the extractor itself at this geometry: two of the frames against the oracle
    p = O.make_params(NF, 1.2, 8, 1, 20)
    for f in (0, NFRAMES - 1):
        rc, ok, od, _ = O.extract(p, frames[f])
        assert rc == 0 and np.array_equal(desc[f], od)
        for name in ("x", "y", "octave", "response"):
            assert np.array_equal(kps[f][name], ok[name]), name
    has = [(rng.random(NF) < 0.93).astype(np.uint8) for _ in range(NFRAMES)]
    outl = [(rng.random(NF) < 0.04).astype(np.uint8) for _ in range(NFRAMES)]
    pre = []
    for _ in range(NFRAMES):
        occ = np.full(NF, -1, np.int32)
        occ[rng.random(NF) < 0.03] = 11   # slots occupied on entry (ORBmatcher.cc:1562)
        pre.append(occ)
    return kps, desc, shifts, has, outl, pre


def _oracle(stream, th, ori):
    kps, desc, shifts, has, outl, pre = stream
    out = []
    for j in range(1, NFRAMES):
        fc = O.OracleFrame(kps[j], desc[j], W, H)
        fl = O.OracleFrame(kps[j - 1], desc[j - 1], W, H)
        out.append(O.search_by_projection_ff(fc, fl, has[j - 1], outl[j - 1], _world(kps[j - 1]), _tcw(*shifts[j]), FX, FY, CX, CY,
                                             th, ori, cur_mp=pre[j]))
    return out


@pytest.fixture(params=[0, 1], ids=["smem-entries", "global-scratch"])
def entry_placement(request):
    if request.param:
        os.environ["ORBFE_SBP_FORCE_SCRATCH"] = "1"
    yield request.param
    os.environ.pop("ORBFE_SBP_FORCE_SCRATCH", None)


@pytest.mark.parametrize("th,ori", [(15.0, True), (15.0, False), (40.0, True)])
def test_sbp_device_at_bench_config(stream, entry_placement, th, ori):
    """orbfe_search_by_projection_device: th=15 is Tracking.cc:565; th=40 multiplies the candidate lists by ~7 (the
    entries then exceed the shared-memory staging area by themselves)."""
    import torch
    kps, desc, shifts, has, outl, pre = stream
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    npairs = NFRAMES - 1
    world = np.stack([_world(kps[f]) for f in range(NFRAMES)])
    flags = np.stack([(has[f] & (1 - outl[f])).astype(np.uint8) for f in range(NFRAMES)])
    T = np.stack([_tcw(*shifts[j]) for j in range(1, NFRAMES)]).reshape(npairs, 12).astype(np.float32)
    d_kps, d_desc = t(kps.view(np.uint8).reshape(NFRAMES, NF, 28)), t(desc)
    d_cnt = t(np.full(NFRAMES, NF, np.int32))
    d_mp = t(np.stack(pre[1:]))
    d_nm = torch.zeros(npairs, dtype=torch.int32, device=dev)
    d_cur, d_last = t(np.arange(1, NFRAMES, dtype=np.int32)), t(np.arange(0, NFRAMES - 1, dtype=np.int32))
    d_world, d_flags, d_T = t(world), t(flags), t(T)
    torch.cuda.synchronize()
    m = fe.ORBmatcher(0.9, ori)
    M.search_by_projection_device(m, npairs, d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), NF, d_cur.data_ptr(),
                                  d_last.data_ptr(), d_world.data_ptr(), d_flags.data_ptr(), d_T.data_ptr(), W, H, 1.2, 8,
                                  FX, FY, CX, CY, th, d_mp.data_ptr(), d_nm.data_ptr())
    m.sync()
    mp, nm = d_mp.cpu().numpy(), d_nm.cpu().numpy()
    ref = _oracle(stream, th, ori)
    tot = 0
    for j in range(npairs):
        assert nm[j] == ref[j][0], (j, nm[j], ref[j][0])
        assert np.array_equal(mp[j], ref[j][1]), j
        tot += ref[j][0]
    assert tot > 8 * 800   # the stream has true matches at this size too
    m.close()


@pytest.mark.parametrize("replay", [0, 1], ids=["fused-kernel", "host-replay"])
def test_sbp_frames_at_bench_config(stream, entry_placement, replay):
    """orbfe_search_by_projection_frames (the call behind ORBmatcher::SearchByProjection(Frame&, const Frame&, float))
    with host views, eight pairs per call, through the fused kernel and through the host-replay path."""
    kps, desc, shifts, has, outl, pre = stream
    views = [M.FrameView(kps[f], desc[f], W, H) for f in range(NFRAMES)]
    m = fe.ORBmatcher(0.9, True)
    fe.lib().orbfe_matcher_force_host_replay(replay)
    try:
        nm, mp = M.search_by_projection_frames(m, views[1:], views[:-1], has[:-1], outl[:-1], [_world(kps[f]) for f in range(NFRAMES - 1)],
                                               [_tcw(*shifts[j]) for j in range(1, NFRAMES)], FX, FY, CX, CY, 15.0, cur_mp=pre[1:])
    finally:
        fe.lib().orbfe_matcher_force_host_replay(0)
    ref = _oracle(stream, 15.0, True)
    for j in range(NFRAMES - 1):
        assert nm[j] == ref[j][0] and np.array_equal(mp[j], ref[j][1]), j
    m.close()
