#!/usr/bin/env python3
"""bench.py -- Mkeypoints/s of the ORB front-end hot path (extract + match) on B200.

Workload (BASELINE.json configs[1]): a stream of synthetic 1920x1080 u8 frames, ORBextractor(2000, 1.2, 8,
FAST_SCORE, 20), frame-to-frame ORBmatcher(0.9, true).SearchByProjection(Current, Last, 15).
One "step" = one batch of `--batch` consecutive frames of the stream: extract all of them, then match every
frame against its predecessor.  Mkeypoints/s = keypoints extracted-and-matched / time.

  value : inputs already resident in HBM when the timed region starts (device API of liborbfe.so),
          results read back + matched.
  e2e   : the reference-facing C-ABI call with HOST buffers (orbfe_extract_batch + the matcher), H2D of the
          step's frames from pinned memory and D2H of keypoints/descriptors inside the timed region.

`--impl reference` times the CPU oracle port of the reference path (oracle/) on the host cores.
PyTorch is used only for plumbing: device buffers, pinned host memory, torch.distributed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NFEAT, NLEVELS, SCALE, FAST_TH = 1920, 1080, 2000, 8, 1.2, 20
FX = FY = 1000.0
CX, CY = W / 2.0, H / 2.0
DEPTH = 5.0
MATCH_TH = 15.0  # Tracking.cc:565
METRIC = "Mkeypoints/s extract+match"


def level_sizes():
    # float32 emulation of ORBextractor.cc:461-471,785-786
    inv = np.float32(1.0) / np.float64(np.float32(SCALE))
    inv = np.float32(inv)
    s = np.float32(1.0)
    out = []
    for l in range(NLEVELS):
        out.append((int(np.rint(np.float32(W) * s)), int(np.rint(np.float32(H) * s))))
        s = np.float32(s * inv)
    return out


def algorithmic_bytes():
    """per-frame algorithmic bytes of extract (reads, writes, gathers, outputs)."""
    ls = level_sizes()
    P = sum(w * h for w, h in ls)
    reads = sum(w * h for w, h in ls[:-1]) + 2 * P
    writes = (P - W * H) + P
    gathers = NFEAT * (749 + 512)
    outputs = NFEAT * 60
    return {"P": P, "total": reads + writes + gathers + outputs, "fast_read": P, "blur": 2 * P,
            "resize": sum(w * h for w, h in ls[:-1]) + (P - W * H)}


def make_stream(batch, seed):
    """`batch` consecutive frames: a few base textures, each followed by small translations of itself."""
    from orb_slam_b200.synth import textured_frame, shifted_frame
    frames = np.empty((batch, H, W), np.uint8)
    shifts = np.zeros((batch, 2), np.int32)  # shift of frame i relative to frame i-1 (dx, dy)
    rng = np.random.default_rng(seed)
    nbase = max(1, min(4, batch // 4))
    per = (batch + nbase - 1) // nbase
    i = 0
    for b in range(nbase):
        base = textured_frame(W, H, seed=seed * 100 + b)
        cur = base
        for k in range(per):
            if i >= batch:
                break
            if k > 0:
                dx, dy = int(rng.integers(-6, 7)), int(rng.integers(-4, 5))
                cur = shifted_frame(cur, dx, dy, seed=seed * 1000 + i)
                shifts[i] = (dx, dy)
            frames[i] = cur
            i += 1
    return frames, shifts


def tcw_for_shift(dx, dy):
    """Camera translation that moves every point at depth DEPTH by (dx, dy) pixels."""
    T = np.zeros((3, 4), np.float32)
    T[0, 0] = T[1, 1] = T[2, 2] = 1.0
    T[0, 3] = dx * DEPTH / FX
    T[1, 3] = dy * DEPTH / FY
    return T


def backproject(kps):
    w = np.empty((len(kps), 3), np.float32)
    w[:, 0] = (kps["x"] - np.float32(CX)) / np.float32(FX) * np.float32(DEPTH)
    w[:, 1] = (kps["y"] - np.float32(CY)) / np.float32(FY) * np.float32(DEPTH)
    w[:, 2] = DEPTH
    return w


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    # the sampler runs from before the warm-up (nvidia-smi needs ~1 s to start); only samples whose arrival time
    # falls inside [t_begin, t_end] (the timed region) are used

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []
        self.t_begin = None
        self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        inside = [ln for t, ln in self.lines if self.t_begin is not None and self.t_begin <= t <= (self.t_end or 1e30)]
        scope = "timed region"
        if not inside:
            inside, scope = [ln for _, ln in self.lines], "whole run (timed region shorter than the sampling period)"
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


# ------------------------------------------------------------------------------------------------
# CPU side (oracle port of the reference path): the only place bench.py executes oracle/
# ------------------------------------------------------------------------------------------------
def cpu_reference_pass(frames, shifts, threads):
    """Oracle extract + SearchByProjection over consecutive frames with `threads` host threads.
    The frame list is repeated until every thread has at least two extraction tasks (with fewer tasks than threads the
    machine would sit half idle and the baseline would be understated).  Returns (keypoints processed, seconds)."""
    import oracle as O
    from concurrent.futures import ThreadPoolExecutor
    n = len(frames)
    reps = max(1, -(-2 * threads // n))
    tasks = n * reps

    def extract_one(t):
        p = O.make_params(NFEAT, SCALE, NLEVELS, 1, FAST_TH)
        rc, k, d, _ = O.extract(p, frames[t % n])
        assert rc == 0
        return k, d

    def match_one(t, feats):
        i = t % n
        if i == 0:
            return 0
        (kl, dl), (kc, dc) = feats[t - 1], feats[t]
        fl = O.OracleFrame(kl, dl, W, H, SCALE, NLEVELS)
        fc = O.OracleFrame(kc, dc, W, H, SCALE, NLEVELS)
        nm, _ = O.search_by_projection_ff(fc, fl, np.ones(fl.n, np.uint8), np.zeros(fl.n, np.uint8), backproject(kl),
                                          tcw_for_shift(*shifts[i]), FX, FY, CX, CY, MATCH_TH, True)
        return nm

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda t: t, range(threads)))          # spin the worker threads up outside the timed region
        t0 = time.perf_counter()
        feats = list(ex.map(extract_one, range(tasks)))
        list(ex.map(lambda t: match_one(t, feats), range(tasks)))
        dt = time.perf_counter() - t0
    return sum(len(k) for k, _ in feats), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    nframes = max(2, min(cores, 64))
    frames, shifts = make_stream(nframes, seed=7)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_pass(frames[:2], shifts[:2], cores)
    tot_kp, tot_t = 0, 0.0
    for _ in range(args.steps):
        kp, dt = cpu_reference_pass(frames, shifts, cores)
        tot_kp += kp
        tot_t += dt
    val = tot_kp / tot_t / 1e6
    kp1, dt1 = cpu_reference_pass(frames[:2], shifts[:2], 1)   # the reference's own mode: one Tracking thread
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mkeypoints/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: 1920x1080 u8 stream, 2000 kp, 8 levels, scale 1.2, SearchByProjection th=15",
                       "frames_per_step": nframes},
            "cpu_baseline": {"value": val, "unit": "Mkeypoints/s", "cores": cores, "kind": "port",
                             "single_thread_value": kp1 / dt1 / 1e6,
                             "sample": "%d distinct frames, repeated to %d extract+match tasks per step, x %d steps, CPU oracle port on "
                                       "%d threads (reference cannot be compiled: needs OpenCV 2.4/ROS/Boost)"
                                       % (nframes, nframes * max(1, -(-2 * cores // nframes)), args.steps, cores)},
            "e2e": {"value": val, "unit": "Mkeypoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="orbfe", choices=["orbfe", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per step per GPU")
    ap.add_argument("--chunks", type=int, default=1, help="split a step into chunks: extract chunk k+1 overlaps match chunk k")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import orb_slam_b200 as fe
    from orb_slam_b200 import matching as M

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: liborbfe has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    B = args.batch
    frames_np, shifts = make_stream(B, seed=11 + rank)
    h_frames = torch.from_numpy(frames_np).pin_memory()                 # pinned host input (e2e path)
    d_frames = torch.from_numpy(frames_np).to(dev)                      # resident input (value path)
    h_kps = torch.empty((B, NFEAT, 28), dtype=torch.uint8).pin_memory()
    h_desc = torch.empty((B, NFEAT, 32), dtype=torch.uint8).pin_memory()
    h_cnt = torch.empty((B,), dtype=torch.int32).pin_memory()
    d_kps = torch.empty((B, NFEAT, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.empty((B, NFEAT, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.empty((B,), dtype=torch.int32, device=dev)
    kps_np = h_kps.numpy().view(fe.KP_DTYPE).reshape(B, NFEAT)
    desc_np = h_desc.numpy()
    cnt_np = h_cnt.numpy()

    ex = fe.ORBextractor(NFEAT, SCALE, NLEVELS, fe.FAST_SCORE, FAST_TH, device=local_rank)
    mt = fe.ORBmatcher(0.9, True, device=local_rank)
    ex.set_profiling(True)
    # a dedicated (non-default) stream: the library's kernels and torch's async copies share it
    stream = torch.cuda.Stream(device=dev)
    Tcws = [tcw_for_shift(*shifts[i]) for i in range(B)]
    prev = {"view": None, "kps": None}
    stage_acc, stage_n = {}, [0]
    kp_total = [0]
    launches = [0]
    host_t = {"extract_call": 0.0, "e2e_extract_call": 0.0, "match_call": 0.0, "views": 0.0, "n": 0}

    ones_u8, zeros_u8 = np.ones(NFEAT, np.uint8), np.zeros(NFEAT, np.uint8)

    def prepare_views(kps_a, desc_a, cnt_a):
        """Frame views + synthetic map points of one step's host results, plus a private copy of the step's last frame
        (it is the Last frame of the next step's first pair, and the pinned result buffers are reused)."""
        t0 = time.perf_counter()
        views = [M.FrameView(kps_a[i, :cnt_a[i]], desc_a[i, :cnt_a[i]], W, H, SCALE, NLEVELS) for i in range(B)]
        world = [backproject(f.kps) for f in views]
        tail = (M.FrameView(views[B - 1].kps.copy(), views[B - 1].desc.copy(), W, H, SCALE, NLEVELS), world[B - 1].copy())
        host_t["views"] += time.perf_counter() - t0
        return views, world, tail

    def match_step(prepared, prev_prepared, matcher):
        views, world_cur, tail = prepared
        t1 = time.perf_counter()
        pv, pw = prev_prepared[2] if prev_prepared is not None else tail
        lasts = [pv] + views[:-1]
        world = [pw] + world_cur[:-1]
        has = [ones_u8[:f.n] for f in lasts]
        outl = [zeros_u8[:f.n] for f in lasts]
        nm, _ = M.search_by_projection_frames(matcher, views, lasts, has, outl, world, Tcws, FX, FY, CX, CY, MATCH_TH)
        host_t["match_call"] += time.perf_counter() - t1
        host_t["n"] += 1
        return int(nm.sum())

    def collect_stages():
        for name, ms in ex.stage_times():
            stage_acc[name] = stage_acc.get(name, 0.0) + ms
        stage_n[0] += 1

    # ---- device-resident pipeline (value path).  Two buffer sets: while step t is matched and read back on
    # stream_m, step t+1 is already being extracted on stream_x.  Slot B of a set holds the previous step's last frame.
    def make_set():
        S = {}
        S["kps"] = torch.zeros((B + 1, NFEAT, 28), dtype=torch.uint8, device=dev)
        S["desc"] = torch.zeros((B + 1, NFEAT, 32), dtype=torch.uint8, device=dev)
        S["cnt"] = torch.zeros((B + 1,), dtype=torch.int32, device=dev)
        S["world"] = torch.zeros((B + 1, NFEAT, 3), dtype=torch.float32, device=dev)
        S["mp"] = torch.empty((B, NFEAT), dtype=torch.int32, device=dev)
        S["nm"] = torch.zeros((B,), dtype=torch.int32, device=dev)
        S["kxy"] = S["kps"].view(torch.float32).view(B + 1, NFEAT, 7)
        S["h_kps"] = torch.empty((B, NFEAT, 28), dtype=torch.uint8).pin_memory()
        S["h_desc"] = torch.empty((B, NFEAT, 32), dtype=torch.uint8).pin_memory()
        S["h_cnt"] = torch.zeros((B,), dtype=torch.int32).pin_memory()
        S["h_mp"] = torch.empty((B, NFEAT), dtype=torch.int32).pin_memory()
        S["h_nm"] = torch.zeros((B,), dtype=torch.int32).pin_memory()
        S["ev_x"] = torch.cuda.Event()
        S["ev_done"] = torch.cuda.Event()
        return S

    dsets = [make_set(), make_set()]
    d_flags = torch.ones((B + 1, NFEAT), dtype=torch.uint8, device=dev)   # every feature carries a map point
    d_T = torch.from_numpy(np.stack(Tcws).reshape(B, 12)).to(dev)
    d_cur = torch.arange(0, B, dtype=torch.int32, device=dev)
    d_last = torch.tensor([B] + list(range(0, B - 1)), dtype=torch.int32, device=dev)
    stream_m = torch.cuda.Stream(device=dev, priority=-1)   # the short matcher kernel goes first when SMs free up
    stream_b = stream_m
    dev_state = {"step": 0}

    def enqueue_device_step():
        st = dev_state["step"]
        S, Pv = dsets[st & 1], dsets[(st - 1) & 1]
        with torch.cuda.stream(stream):
            stream.wait_event(S["ev_done"])          # the set's previous results have left the device
            ex.extract_batch_device(d_frames.data_ptr(), W, H, W, W * H, B, S["kps"].data_ptr(), S["desc"].data_ptr(),
                                    S["cnt"].data_ptr(), stream.cuda_stream)
            # synthetic map points: back-project every keypoint at depth DEPTH (same float32 ops as backproject())
            S["world"][:B, :, 0] = (S["kxy"][:B, :, 0] - CX) / FX * DEPTH
            S["world"][:B, :, 1] = (S["kxy"][:B, :, 1] - CY) / FY * DEPTH
            S["world"][:B, :, 2] = DEPTH
            S["ev_x"].record(stream)
        with torch.cuda.stream(stream_m):
            stream_m.wait_event(S["ev_x"])
            # slot B <- last frame of the previous step (extracted earlier on `stream`, matched earlier on stream_m)
            S["kps"][B].copy_(Pv["kps"][B - 1]); S["desc"][B].copy_(Pv["desc"][B - 1])
            S["cnt"][B:B + 1].copy_(Pv["cnt"][B - 1:B]); S["world"][B].copy_(Pv["world"][B - 1])
            S["mp"].fill_(-1)
            M.search_by_projection_device(mt, B, S["kps"].data_ptr(), S["desc"].data_ptr(), S["cnt"].data_ptr(), NFEAT,
                                          d_cur.data_ptr(), d_last.data_ptr(), S["world"].data_ptr(), d_flags.data_ptr(),
                                          d_T.data_ptr(), W, H, SCALE, NLEVELS, FX, FY, CX, CY, MATCH_TH,
                                          S["mp"].data_ptr(), S["nm"].data_ptr(), stream_m.cuda_stream)
            S["h_cnt"].copy_(S["cnt"][:B], non_blocking=True)
            S["h_nm"].copy_(S["nm"], non_blocking=True)
            S["h_kps"].copy_(S["kps"][:B], non_blocking=True)
            S["h_desc"].copy_(S["desc"][:B], non_blocking=True)
            S["h_mp"].copy_(S["mp"], non_blocking=True)
            S["ev_done"].record(stream_m)
        launches[0] += ex.last_launches() + 1
        dev_state["step"] = st + 1
        return S

    def finish_device_step(S):
        S["ev_done"].synchronize()
        kp_total[0] += int(S["h_cnt"].numpy().sum())
        return int(S["h_nm"].numpy().sum())

    def run_device(steps):
        t0 = time.perf_counter()
        nm, pending = 0, None
        for _ in range(steps):
            S = enqueue_device_step()
            if pending is not None:
                nm += finish_device_step(pending)
            pending = S
        nm += finish_device_step(pending)
        stream.synchronize()
        collect_stages()
        host_t["extract_call"] += time.perf_counter() - t0
        return nm

    # ---- e2e: host buffers in, host results out.  A stream of frames is processed as a software pipeline over the
    # public C-ABI calls only: two extractor handles (each with its own device buffers and streams) alternate steps on
    # two host threads, so the upload of step t+1 overlaps the kernels of step t and there is no bubble at a call
    # boundary; frame views (the host-side Frame objects of the reference) are built on their own worker threads and
    # a last thread matches finished steps in order (orbfe_search_by_projection_frames on host views).
    # ctypes releases the GIL inside the calls.  Outputs rotate through four pinned buffer sets.
    from concurrent.futures import ThreadPoolExecutor
    NEX = max(1, int(os.environ.get("ORBFE_E2E_EXTRACTORS", "2")))   # extractor handles in flight
    NBUF = NEX + 2
    e2e_bufs = []
    for _ in range(NBUF):
        hk = torch.empty((B, NFEAT, 28), dtype=torch.uint8).pin_memory()
        hd = torch.empty((B, NFEAT, 32), dtype=torch.uint8).pin_memory()
        hc = torch.empty((B,), dtype=torch.int32).pin_memory()
        e2e_bufs.append((hk, hd, hc, hk.numpy().view(fe.KP_DTYPE).reshape(B, NFEAT), hd.numpy(), hc.numpy()))
    e2e_ex = [ex] + [fe.ORBextractor(NFEAT, SCALE, NLEVELS, fe.FAST_SCORE, FAST_TH, device=local_rank) for _ in range(NEX - 1)]
    e2e_ex_pool = [ThreadPoolExecutor(max_workers=1) for _ in range(NEX)]
    E2E_MODE = int(os.environ.get("ORBFE_E2E_BATCH_MODE", "0"))   # measured: chunked 28.7 vs phased 26.8 Mkp/s with two handles
    e2e_mt = (mt, fe.ORBmatcher(0.9, True, device=local_rank))
    e2e_match_pool = (ThreadPoolExecutor(max_workers=1), ThreadPoolExecutor(max_workers=1))
    e2e_views_pool = ThreadPoolExecutor(max_workers=2)

    def extract_step(st):
        t0 = time.perf_counter()
        x = e2e_ex[st % NEX]
        x.set_batch_mode(E2E_MODE)
        hk, hd, hc, _, _, c_np = e2e_bufs[st % NBUF]
        x.extract_batch_ptr(h_frames.data_ptr(), W, H, W, W * H, B, hk.data_ptr(), hd.data_ptr(), NFEAT, hc.data_ptr())
        host_t["e2e_extract_call"] += time.perf_counter() - t0
        return x.last_launches(), int(c_np.sum())

    def run_e2e(steps):
        nm = 0
        ex_futs, m_futs, v_futs = {}, {}, {}

        def finish_extract(st):
            nl, nk = ex_futs.pop(st).result()
            launches[0] += nl
            kp_total[0] += nk
            _, _, _, k_np, d_np, c_np = e2e_bufs[st % NBUF]
            vf = e2e_views_pool.submit(prepare_views, k_np, d_np, c_np)   # frame views: a stage of its own
            pf = v_futs.get(st - 1)
            v_futs[st] = vf
            v_futs.pop(st - 2, None)
            # two matcher handles alternate so that the queueing latency of one call hides behind the other
            m_futs[st] = e2e_match_pool[st & 1].submit(
                lambda f=vf, p=pf, mm=e2e_mt[st & 1]: match_step(f.result(), p.result() if p is not None else None, mm))

        for st in range(steps):
            if st - (NBUF - 1) in m_futs:             # buffer set st % NBUF was last used by step st - NBUF
                nm += m_futs.pop(st - (NBUF - 1)).result()
            ex_futs[st] = e2e_ex_pool[st % NEX].submit(extract_step, st)
            if st >= NEX - 1:
                finish_extract(st - (NEX - 1))
        for st in sorted(ex_futs):
            finish_extract(st)
        for st in sorted(m_futs):
            nm += m_futs[st].result()
        return nm

    def timed(run_fn, steps):
        kp_total[0] = 0
        launches[0] = 0
        c0 = [sum(x) for x in zip(*(m_.counters() for m_ in e2e_mt))]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        nm = run_fn(steps)
        stream.wait_stream(stream_b)
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        c1 = [sum(x) for x in zip(*(m_.counters() for m_ in e2e_mt))]
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        k = torch.tensor([kp_total[0], nm], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(k, op=dist.ReduceOp.SUM)
            dist.barrier()
        return {"ms": float(t[0]), "wall_ms": float(t[1]), "kp": float(k[0]), "matches": float(k[1]),
                "launches": launches[0] + (c1[2] - c0[2]), "mh2d": c1[0] - c0[0], "md2h": c1[1] - c0[1]}

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    run_device(args.warmup)
    ex.set_profiling(False)          # stage events only during the device-resident timed region
    run_e2e(args.warmup)
    ex.set_profiling(True)
    ex.stage_times()                 # flush
    stage_acc.clear()
    stage_n[0] = 0

    sampler.t_begin = time.perf_counter()
    r_dev = timed(run_device, args.steps)
    stages = {k: v / max(args.steps, 1) for k, v in stage_acc.items()}   # ms per step
    ex.set_profiling(False)
    r_e2e = timed(run_e2e, args.steps)
    sampler.t_end = time.perf_counter()
    clocks = sampler.stop() if rank == 0 else None

    # stage times once more with the extractor alone on the GPU (in the timed region above the matcher of the
    # previous step shares the SMs with whatever stage is running): a stable figure for kernel-to-kernel comparisons
    ex.set_profiling(True)
    ex.stage_times()
    S0 = dsets[0]
    iso_acc, iso_n = {}, 10
    for _ in range(iso_n):
        ex.extract_batch_device(d_frames.data_ptr(), W, H, W, W * H, B, S0["kps"].data_ptr(), S0["desc"].data_ptr(),
                                S0["cnt"].data_ptr(), stream.cuda_stream)
    stream.synchronize()
    for name, ms in ex.stage_times():
        iso_acc[name] = iso_acc.get(name, 0.0) + ms / iso_n
    ex.set_profiling(False)

    # latency of the reference's own usage pattern: one frame per call, host image in, host keypoints/descriptors out
    # (ORBextractor::operator(), src/ORBextractor.cc:718-779), nothing else on the GPU
    lat = []
    hk1, hd1, hc1 = e2e_bufs[0][0], e2e_bufs[0][1], e2e_bufs[0][2]
    for i in range(25):
        t0 = time.perf_counter()
        ex.extract_batch_ptr(h_frames.data_ptr() + (i % B) * W * H, W, H, W, W * H, 1, hk1.data_ptr(), hd1.data_ptr(), NFEAT,
                             hc1.data_ptr())
        lat.append((time.perf_counter() - t0) * 1e3)
    single_ms = float(np.median(lat[5:]))

    if rank == 0:
        ab = algorithmic_bytes()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # dominant kernel of the extract pipeline by measured stage time
        kernel_stages = {k: v for k, v in stages.items() if k not in ("ingest", "h2d", "d2h")}
        dom = max(kernel_stages, key=kernel_stages.get) if kernel_stages else None
        dom_bytes = {"fast_nms": ab["fast_read"], "blur7": ab["blur"], "pyramid": ab["resize"]}.get(dom, ab["total"]) * B
        dom_ms = kernel_stages.get(dom, 0.0) if dom else 0.0
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        ext_ms = sum(kernel_stages.values())
        traffic = None
        try:  # DRAM bytes of the dominant kernel from the committed ncu capture, scaled to this launch's frame count
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_fast_nms_traffic.json")))
            if dom == "fast_nms":
                traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["frames_in_launch"] * B
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                "note": "fast_nms is integer-ALU-pipe bound (ncu: 78.7% of ALU peak, 3% of DRAM peak), see profiles/README.md",
                "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": dom_ms,
                "extract_all_kernels": {"algorithmic_bytes": ab["total"] * B, "ms": ext_ms,
                                        "achieved": ab["total"] * B / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0},
                "stage_ms": stages, "stage_ms_extractor_alone": iso_acc}
        value = r_dev["kp"] / (r_dev["ms"] * 1e-3) / 1e6
        e2e_val = r_e2e["kp"] / (r_e2e["ms"] * 1e-3) / 1e6
        line = {"metric": METRIC, "value": value, "unit": "Mkeypoints/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r_dev["ms"] / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "configs[1]: 1920x1080 u8 stream, 2000 kp, 8 levels, scale 1.2, "
                                       "frame-to-frame SearchByProjection th=15",
                           "frames_per_step_per_gpu": B, "parallelism": "frames sharded one stream per GPU, no data-path collective",
                           "l2": "inputs+pyramids per step (%.0f MB) larger than L2 (126 MB)" % (B * (W * H + 2 * ab["P"]) / 1e6)},
                "e2e": {"value": e2e_val, "unit": "Mkeypoints/s", "ms_per_step": r_e2e["ms"] / args.steps,
                        # whole job: every rank moves the same amount
                        "h2d_bytes_per_step": world * (B * W * H + r_e2e["mh2d"] // args.steps),
                        "d2h_bytes_per_step": world * (B * (NFEAT * 60 + 4) + r_e2e["md2h"] // args.steps)},
                "gpu_launches": int(r_dev["launches"]),
                "matches_per_step": r_dev["matches"] / args.steps / world,
                "keypoints_per_step": r_dev["kp"] / args.steps,
                "wall_ms_per_step": r_dev["wall_ms"] / args.steps,
                "host_ms_per_step": {k: 1e3 * v / max(host_t["n"], 1) for k, v in host_t.items() if k != "n"},
                "single_frame_latency_ms": single_ms,
                "roofline": roof, "clocks": clocks}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            nfr = max(2, min(B, 64))
            kp, dt = cpu_reference_pass(frames_np[:nfr], shifts[:nfr], cores)
            ntasks = nfr * max(1, -(-2 * cores // nfr))
            kp1, dt1 = cpu_reference_pass(frames_np[:2], shifts[:2], 1)   # the reference's own mode: one Tracking thread
            line["cpu_baseline"] = {"value": kp / dt / 1e6, "unit": "Mkeypoints/s", "cores": cores, "kind": "port",
                                    "single_thread_value": kp1 / dt1 / 1e6,
                                    "sample": "%d of the step's frames repeated to %d extract+match tasks (about %.0f CPU-seconds), "
                                              "CPU oracle port on %d threads, %.1f s wall" % (nfr, ntasks, 0.11 * ntasks, cores, dt)}
        print(json.dumps(line))
    for x in e2e_ex:
        x.close()
    for m_ in e2e_mt:
        m_.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
