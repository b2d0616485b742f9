#!/usr/bin/env python3
"""The two BASELINE configurations with a real exchange step between GPUs, through the C-ABI of liborbfe.so
(include/orbfe_comm.h).  One process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multi_gpu.py --what rig,dbsweep

  rig      config 4: an N-camera rig at 1280x720, one camera per GPU, `--slots` time steps per exchange.  Every rank extracts
           its camera's frames (ORBextractor, 2000 kp, 8 levels), all ranks exchange keypoints + descriptors, every rank runs
           cross-camera ORBmatcher::SearchForInitialization (ORBmatcher.cc:598-713) of its camera against the next one.
           The exchange is measured two ways: `nccl` = orbfe_allgather_desc (ncclAllGather of the extractor's output blocks)
           and `fused` = OrbfeRigExchange (the descriptor kernel stores into every peer's buffer over NVLink).
  dbsweep  config 5: 1 query (2000 descriptors) x 10 000 keyframes x 2000 descriptors, database row-sharded over the ranks:
           ncclBroadcast of the query, local best/second sweep, all-gather of the per-keyframe results.

Also importable: bench.py --workload rig8 / dbsweep calls run_rig / run_dbsweep.  Rank 0 prints one JSON object per workload;
results are checked against the CPU oracle outside the timed regions.  torch is plumbing (device buffers, process group)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

RW, RH, RNF, RNL = 1280, 720, 2000, 8


def _setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return torch, dist, world, rank, local


def _max_over_ranks(torch, dist, dev, vals):
    t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=dev)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def rig_frames(rank, nslots, seed=40):
    """Camera `rank` of the rig: every time step t is one scene seen by all cameras with a per-camera offset."""
    from orb_slam_b200.synth import textured_frame, shifted_frame
    out = np.empty((nslots, RH, RW), np.uint8)
    for t in range(nslots):
        base = textured_frame(RW, RH, seed=seed + t)
        out[t] = shifted_frame(base, 12 * rank, 4 * rank, seed=1000 * rank + t) if rank else base
    return out


def run_rig(args, ctx=None):
    torch, dist, world, rank, local = ctx or _setup()
    import orb_slam_b200 as fe
    from orb_slam_b200 import matching as M, comm as CM
    dev = torch.device("cuda", local)
    T = args.slots
    frames = rig_frames(rank, T)
    d_frames = torch.from_numpy(frames).to(dev)
    ex = fe.ORBextractor(RNF, 1.2, RNL, fe.FAST_SCORE, 20, device=local)
    mt = fe.ORBmatcher(0.9, True, device=local)
    comm = CM.Comm.create(torch, dist, local)
    xch = CM.RigExchange(comm, RNF, T)
    stream = torch.cuda.Stream(device=dev)
    s = stream.cuda_stream
    # own outputs (nccl variant) and gathered arrays
    d_kps = torch.zeros((T, RNF, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.zeros((T, RNF, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros((T,), dtype=torch.int32, device=dev)
    g_kps = torch.zeros((world, T, RNF, 28), dtype=torch.uint8, device=dev)
    g_desc = torch.zeros((world, T, RNF, 32), dtype=torch.uint8, device=dev)
    g_cnt = torch.zeros((world, T), dtype=torch.int32, device=dev)
    nxt = (rank + 1) % world
    d_f1 = torch.arange(rank * T, rank * T + T, dtype=torch.int32, device=dev)
    d_f2 = torch.arange(nxt * T, nxt * T + T, dtype=torch.int32, device=dev)
    d_prev = torch.zeros((T, RNF, 2), dtype=torch.float32, device=dev)
    d_m12 = torch.full((T, RNF), -1, dtype=torch.int32, device=dev)
    d_nm = torch.zeros((T,), dtype=torch.int32, device=dev)

    def match(all_kps_ptr, all_desc_ptr, all_cnt_ptr, kps_view):
        # vbPrevMatched starts as the feature's own position (Tracking.cc:343-345)
        d_prev.copy_(kps_view[rank].view(torch.float32).view(T, RNF, 7)[:, :, 0:2])
        M.search_for_initialization_device(mt, T, all_kps_ptr, all_desc_ptr, all_cnt_ptr, RNF, d_f1.data_ptr(), d_f2.data_ptr(),
                                           d_prev.data_ptr(), RW, RH, 100, d_m12.data_ptr(), d_nm.data_ptr(), s)

    def step_nccl():
        with torch.cuda.stream(stream):
            ex.extract_batch_device(d_frames.data_ptr(), RW, RH, RW, RW * RH, T, d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), s)
            comm.allgather_desc(d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), RNF, T, g_kps.data_ptr(), g_desc.data_ptr(),
                                g_cnt.data_ptr(), s)
            match(g_kps.data_ptr(), g_desc.data_ptr(), g_cnt.data_ptr(), g_kps)

    own_view = {}

    def step_fused():
        # two library calls, no extra launch: the descriptor kernel stores into every rank's buffer and publishes the epoch;
        # the matcher kernel polls the epoch flags before its first read and releases the epoch when its last block is done
        with torch.cuda.stream(stream):
            xch.extract(ex, d_frames.data_ptr(), RW, RH, RW, RW * RH, s)
            a, _, _ = xch.buffers_of_current_epoch()
            if a not in own_view:   # this rank's own slot of each buffer half: written locally, stream-ordered (views cached: two halves)
                own_view[a] = _as_tensor(torch, a + rank * T * RNF * 28, (T, RNF, 28), dev).view(torch.float32).view(T, RNF, 7)[:, :, 0:2]
            d_prev.copy_(own_view[a])
            xch.search_for_initialization(mt, T, d_f1.data_ptr(), d_f2.data_ptr(), d_prev.data_ptr(), RW, RH, 100, d_m12.data_ptr(),
                                          d_nm.data_ptr(), s)

    # Throughput form of the fused step: the exchange's buffer halves alternate by epoch, so the extraction of time step t+1
    # (stream) may run while the matcher of step t (stream_m: one CTA per pair, one busy warp each) is still going.  Depth 2:
    # extract(t+2) rewrites the half of step t and is held back until THIS rank's matcher of step t is done -- its descriptor
    # kernel would otherwise spin on an acknowledgement that a kernel queued behind it has to produce.
    stream_m = torch.cuda.Stream(device=dev)
    ev_x = [torch.cuda.Event(), torch.cuda.Event()]
    ev_m = [torch.cuda.Event(), torch.cuda.Event()]
    pipe = {"k": 0}

    def step_fused_pipelined():
        k = pipe["k"]
        with torch.cuda.stream(stream):
            if k >= 2:
                stream.wait_event(ev_m[k & 1])
            xch.extract(ex, d_frames.data_ptr(), RW, RH, RW, RW * RH, s)
            ev_x[k & 1].record(stream)
            a, _, _ = xch.buffers_of_current_epoch()
            if a not in own_view:
                own_view[a] = _as_tensor(torch, a + rank * T * RNF * 28, (T, RNF, 28), dev).view(torch.float32).view(T, RNF, 7)[:, :, 0:2]
        with torch.cuda.stream(stream_m):
            stream_m.wait_event(ev_x[k & 1])
            d_prev.copy_(own_view[a])
            xch.search_for_initialization(mt, T, d_f1.data_ptr(), d_f2.data_ptr(), d_prev.data_ptr(), RW, RH, 100, d_m12.data_ptr(),
                                          d_nm.data_ptr(), stream_m.cuda_stream)
            ev_m[k & 1].record(stream_m)
        pipe["k"] = k + 1

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        stream.wait_stream(stream_m)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        stream.wait_stream(stream_m)
        e1.record(stream)
        stream.synchronize()
        return _max_over_ranks(torch, dist, dev, [e0.elapsed_time(e1) / steps])[0]

    # exchange alone (no extraction, no matching): ncclAllGather of ready blocks vs nothing to compare for fused (it has no separate pass)
    def only_allgather():
        with torch.cuda.stream(stream):
            comm.allgather_desc(d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), RNF, T, g_kps.data_ptr(), g_desc.data_ptr(), g_cnt.data_ptr(), s)

    def only_extract():
        with torch.cuda.stream(stream):
            ex.extract_batch_device(d_frames.data_ptr(), RW, RH, RW, RW * RH, T, d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), s)

    def only_extract_fused():
        with torch.cuda.stream(stream):
            xch.extract(ex, d_frames.data_ptr(), RW, RH, RW, RW * RH, s)
            xch.wait(s)
            xch.release(s)

    ms_nccl = timed(step_nccl, args.steps, args.warmup)
    nm_nccl = d_nm.cpu().numpy().copy()
    m12_nccl = d_m12.cpu().numpy().copy()
    gk_nccl, gd_nccl, gc_nccl = g_kps.cpu().numpy().copy(), g_desc.cpu().numpy().copy(), g_cnt.cpu().numpy().copy()
    ms_fused = timed(step_fused, args.steps, args.warmup)
    xch.check(s)
    nm_fused = d_nm.cpu().numpy().copy()
    m12_fused = d_m12.cpu().numpy().copy()
    a, b, c = xch.buffers()
    gk_f = _as_tensor(torch, a, (world, T, RNF, 28), dev).cpu().numpy()
    gd_f = _as_tensor(torch, b, (world, T, RNF, 32), dev).cpu().numpy()
    gc_f = _as_tensor(torch, c, (world, T), dev, torch.int32).cpu().numpy()
    ms_pipe = timed(step_fused_pipelined, args.steps, args.warmup)
    xch.check(s)
    same_pipe = bool(np.array_equal(nm_fused, d_nm.cpu().numpy()) and np.array_equal(m12_fused, d_m12.cpu().numpy()))
    ms_ag = timed(only_allgather, args.steps, args.warmup)
    ms_ex = timed(only_extract, args.steps, args.warmup)
    ms_exf = timed(only_extract_fused, args.steps, args.warmup)
    xch.check(s)
    # the descriptor kernel itself, plain vs with the remote stores + publish (CUDA events around the stage, extractor alone)
    stage = {}
    ex.set_profiling(True)
    for name, fn in (("plain", only_extract), ("exchange", only_extract_fused)):
        ex.stage_times()
        for _ in range(args.steps):
            fn()
        stream.synchronize()
        acc = {}
        for k, v in ex.stage_times():
            acc[k] = acc.get(k, 0.0) + v / args.steps
        stage[name] = acc
    ex.set_profiling(False)
    xch.check(s)
    desc_ms = _max_over_ranks(torch, dist, dev, [stage["plain"].get("describe", 0.0), stage["exchange"].get("describe", 0.0)])

    # ---- checks (outside the timed regions) ----
    same_gather = bool(np.array_equal(gc_nccl, gc_f) and np.array_equal(gk_nccl, gk_f) and np.array_equal(gd_nccl, gd_f))
    same_match = bool(np.array_equal(nm_nccl, nm_fused) and np.array_equal(m12_nccl, m12_fused))
    oracle_ok = None
    if rank == 0 and not args.no_parity:
        import oracle as O
        oracle_ok = True
        for t in range(min(T, 2)):
            k1 = gk_f[rank, t].reshape(-1).view(fe.KP_DTYPE)[:gc_f[rank, t]]
            k2 = gk_f[nxt, t].reshape(-1).view(fe.KP_DTYPE)[:gc_f[nxt, t]]
            o1 = O.OracleFrame(k1, gd_f[rank, t][:len(k1)], RW, RH)
            o2 = O.OracleFrame(k2, gd_f[nxt, t][:len(k2)], RW, RH)
            prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
            n_o, m_o, _ = O.search_for_initialization(o1, o2, prev, 100, nnratio=0.9, check_orientation=True)
            oracle_ok &= bool(n_o == nm_fused[t] and np.array_equal(m_o, m12_fused[t][:len(k1)]))
            if t == 0:   # the extractor behind the exchange, too
                rc, ok, od, _ = O.extract(O.make_params(RNF, 1.2, RNL, 1, 20), frames[0])
                oracle_ok &= bool(rc == 0 and np.array_equal(od, gd_f[rank, 0][:len(ok)]) and np.array_equal(ok["x"], k1["x"]))
    kp_step = float(world * T * RNF)
    flags = _max_over_ranks(torch, dist, dev, [0.0 if same_gather else 1.0, 0.0 if same_match else 1.0, 0.0 if same_pipe else 1.0])
    out = None
    if rank == 0:
        out = {"workload": "configs[3]: %d-camera rig 1280x720, one camera per GPU, %d time steps per exchange, 2000 kp, 8 levels; "
                           "cross-camera SearchForInitialization(window 100) of camera r against camera r+1" % (world, T),
               "n_gpus": world, "nccl_version": CM.nccl_version(),
               "ms_per_step": {"extract+ncclAllGather+match": ms_nccl, "extract(fused exchange)+match": ms_fused,
                               "extract(fused exchange) of step t+1 overlapping the match of step t": ms_pipe,
                               "ncclAllGather alone": ms_ag, "extract alone": ms_ex, "extract with fused exchange + wait + release": ms_exf},
               "exchange_cost_ms": {"nccl": ms_ag, "fused (extra time over a plain extract)": ms_exf - ms_ex,
                                    "fused (whole step vs whole nccl step)": ms_fused - ms_nccl},
               "describe_kernel_ms": {"plain": desc_ms[0], "with remote stores + publish": desc_ms[1]},
               "Mkeypoints_per_s": {"nccl": kp_step / (ms_nccl * 1e-3) / 1e6, "fused": kp_step / (ms_fused * 1e-3) / 1e6,
                                    "fused, pipelined over time steps": kp_step / (ms_pipe * 1e-3) / 1e6},
               "matches_identical_pipelined_vs_fused_all_ranks": flags[2] == 0.0,
               "nvlink_bytes_per_step_per_gpu": T * RNF * 60 * (world - 1),
               "matches_rank0": int(nm_fused.sum()), "gathered_identical_nccl_vs_fused_all_ranks": flags[0] == 0.0,
               "matches_identical_nccl_vs_fused_all_ranks": flags[1] == 0.0, "oracle_check_rank0": oracle_ok}
    xch.close()
    comm.close()
    ex.close()
    mt.close()
    return out


def _as_tensor(torch, addr, shape, dev, dtype=None):
    """A torch view of raw device memory owned by liborbfe (through the CUDA array interface)."""
    dtype = dtype or torch.uint8
    n = int(np.prod(shape))
    itemsize = torch.tensor([], dtype=dtype).element_size()

    class _Holder:
        pass
    h = _Holder()
    typestr = {torch.uint8: "|u1", torch.int32: "<i4", torch.float32: "<f4"}[dtype]
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(addr), False), "version": 2, "strides": None}
    return torch.as_tensor(h, device=dev).view(*shape)


def run_dbsweep(args, ctx=None):
    torch, dist, world, rank, local = ctx or _setup()
    import orb_slam_b200 as fe
    from orb_slam_b200 import comm as CM
    from orb_slam_b200.synth import random_descriptors
    dev = torch.device("cuda", local)
    nq, per, ng = 2000, 2000, args.groups
    comm = CM.Comm.create(torch, dist, local)
    lo, hi = CM.shard_range(ng, world, rank)
    gmax = -(-ng // world)
    # keyframe g's descriptors depend only on g (any rank can regenerate any keyframe for the check)
    def group(g):
        gen = torch.Generator(device=dev)
        gen.manual_seed(7000 + g)
        return torch.randint(0, 256, (per, 32), dtype=torch.uint8, device=dev, generator=gen)
    db = torch.empty(((hi - lo) * per, 32), dtype=torch.uint8, device=dev)
    for g in range(lo, hi):
        db[(g - lo) * per:(g - lo + 1) * per] = group(g)
    q = torch.zeros((nq, 32), dtype=torch.uint8, device=dev)
    q_host = random_descriptors(nq, 1)
    if rank == 0:
        q.copy_(torch.from_numpy(q_host))
    best = torch.zeros((ng, nq), dtype=torch.uint16, device=dev)
    idx = torch.zeros((ng, nq), dtype=torch.int32, device=dev)
    second = torch.zeros((ng, nq), dtype=torch.uint16, device=dev)
    scratch = torch.zeros(((world + 1) * gmax * nq * 8,), dtype=torch.uint8, device=dev)
    m = fe.ORBmatcher(device=local)
    stream = torch.cuda.Stream(device=dev)
    s = stream.cuda_stream

    def sweep():
        with torch.cuda.stream(stream):
            comm.knn2_sweep_sharded(m, q.data_ptr(), nq, 0, db.data_ptr(), ng, per, best.data_ptr(), idx.data_ptr(), second.data_ptr(),
                                    scratch.data_ptr(), s)
    for _ in range(args.warmup):
        sweep()
    stream.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        sweep()
    e1.record(stream)
    stream.synchronize()
    ms = _max_over_ranks(torch, dist, dev, [e0.elapsed_time(e1) / args.steps])[0]
    ok = None
    if not args.no_parity:
        import oracle as O
        ok = True
        for g in sorted(set([0, ng // 3, ng - 1, lo])):   # keyframes owned by different ranks, checked on every rank's gathered copy
            bd, bi, sd = O.knn2(q_host, group(g).cpu().numpy())
            ok &= bool(np.array_equal(best[g].cpu().numpy(), bd) and np.array_equal(idx[g].cpu().numpy(), bi)
                       and np.array_equal(second[g].cpu().numpy(), np.minimum(sd, 65535)))
    bad = _max_over_ranks(torch, dist, dev, [0.0 if (ok is None or ok) else 1.0])[0]
    pairs = float(nq) * ng * per
    out = None
    if rank == 0:
        out = {"workload": "configs[4]: 1 query x %d keyframes x %d descriptors (%d-descriptor query), 256-bit Hamming best/second per keyframe, "
                           "database row-sharded over %d GPU(s); ncclBroadcast(query) + local sweep + ncclAllGather(results)" % (ng, per, nq, world),
               "n_gpus": world, "nccl_version": CM.nccl_version(), "ms_per_query": ms, "Gpairs_per_s": pairs / (ms * 1e-3) / 1e9,
               "db_MB_total": ng * per * 32 / 1e6, "db_MB_per_gpu": (hi - lo) * per * 32 / 1e6,
               "nvlink_bytes_per_query_per_gpu": nq * 32 + (world - 1) * gmax * nq * 8,
               "word_ops_per_s_T": pairs * 8 / (ms * 1e-3) / 1e12, "oracle_check_all_ranks": None if args.no_parity else bad == 0.0}
    m.close()
    comm.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="rig,dbsweep")
    ap.add_argument("--slots", type=int, default=8, help="rig: time steps per exchange")
    ap.add_argument("--groups", type=int, default=10000, help="dbsweep: keyframes in the database")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    ctx = _setup()
    torch, dist, world, rank, local = ctx
    for w in args.what.split(","):
        out = {"rig": run_rig, "dbsweep": run_dbsweep}[w](args, ctx)
        if rank == 0:
            print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
