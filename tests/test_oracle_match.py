"""CPU tests: the matcher half of the oracle against brute-force Python restatements written directly from
the reference loops (ORBmatcher.cc, Frame.cc), on small random frames."""
import numpy as np

import oracle as O
from orb_slam_b200.synth import random_descriptors, noisy_copies

W, H = 640, 480


def _rand_frame(n, seed, desc=None):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, O.KP_DTYPE)
    k["octave"] = rng.integers(0, 4, n)
    k["x"] = rng.uniform(0, W, n).astype(np.float32)
    k["y"] = rng.uniform(0, H, n).astype(np.float32)
    k["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    d = random_descriptors(n, seed) if desc is None else desc
    return k, d


def _ham(a, b):
    return int(np.unpackbits(a ^ b).sum())


def test_hamming():
    d = random_descriptors(50, 3)
    for i in range(49):
        assert O.hamming(d[i], d[i + 1]) == _ham(d[i], d[i + 1])
    assert O.hamming(d[0], d[0]) == 0 and O.hamming(d[0], ~d[0]) == 256


def _grid_py(k):
    gx, gy = np.float32(64) / np.float32(W), np.float32(48) / np.float32(H)
    cells = {}
    for i in range(len(k)):
        vx, vy = np.float32(k["x"][i]) * gx, np.float32(k["y"][i]) * gy
        px = int(np.floor(vx + np.float32(0.5))) if vx >= 0 else int(np.ceil(vx - np.float32(0.5)))   # round(): half away
        py = int(np.floor(vy + np.float32(0.5))) if vy >= 0 else int(np.ceil(vy - np.float32(0.5)))
        if 0 <= px < 64 and 0 <= py < 48:
            cells.setdefault((px, py), []).append(i)
    return cells, gx, gy


def _area_py(k, cells, gx, gy, x, y, r, lo, hi):
    x, y, r = np.float32(x), np.float32(y), np.float32(r)
    x0 = max(0, int(np.floor((x - r) * gx)))
    x1 = min(63, int(np.ceil((x + r) * gx)))
    y0 = max(0, int(np.floor((y - r) * gy)))
    y1 = min(47, int(np.ceil((y + r) * gy)))
    if x0 >= 64 or x1 < 0 or y0 >= 48 or y1 < 0:
        return []
    out = []
    for ix in range(x0, x1 + 1):
        for iy in range(y0, y1 + 1):
            for i in cells.get((ix, iy), []):
                o = k["octave"][i]
                if not (lo == -1 and hi == -1):
                    if lo == hi:
                        if o != lo:
                            continue
                    elif o < lo or o > hi:
                        continue
                if abs(np.float32(k["x"][i]) - x) > r or abs(np.float32(k["y"][i]) - y) > r:
                    continue
                out.append(i)
    return out


def test_features_in_area_order_and_filters():
    k, d = _rand_frame(800, 1)
    f = O.OracleFrame(k, d, W, H)
    cells, gx, gy = _grid_py(k)
    rng = np.random.default_rng(2)
    for _ in range(200):
        x, y, r = rng.uniform(-20, W + 20), rng.uniform(-20, H + 20), rng.uniform(1, 120)
        lo = int(rng.integers(-1, 3))
        hi = lo if rng.random() < 0.4 else (-1 if lo == -1 else lo + int(rng.integers(0, 3)))
        got = f.features_in_area(np.float32(x), np.float32(y), np.float32(r), lo, hi).tolist()
        assert got == _area_py(k, cells, gx, gy, x, y, r, lo, hi)


def _three_max_py(counts):
    m = [0, 0, 0]
    ind = [-1, -1, -1]
    for i, s in enumerate(counts):
        if s > m[0]:
            m, ind = [s, m[0], m[1]], [i, ind[0], ind[1]]
        elif s > m[1]:
            m, ind = [m[0], s, m[1]], [ind[0], i, ind[1]]
        elif s > m[2]:
            m[2], ind[2] = s, i
    if m[1] < np.float32(0.1) * np.float32(m[0]):
        ind[1] = ind[2] = -1
    elif m[2] < np.float32(0.1) * np.float32(m[0]):
        ind[2] = -1
    return ind


def _bin(a1, a2):
    rot = np.float32(a1) - np.float32(a2)
    if rot < 0:
        rot = np.float32(rot + np.float32(360.0))
    v = np.float32(rot * (np.float32(1.0) / np.float32(30)))
    b = int(np.floor(v + np.float32(0.5)))
    return 0 if b == 30 else b


def test_window_search_against_python_loop():
    k1, d1 = _rand_frame(400, 5)
    k2 = k1.copy()
    rng = np.random.default_rng(6)
    k2["x"] = (k1["x"] + rng.uniform(-8, 8, 400)).astype(np.float32)
    k2["y"] = (k1["y"] + rng.uniform(-8, 8, 400)).astype(np.float32)
    k2["angle"] = ((k1["angle"] + rng.uniform(-10, 10, 400)) % 360).astype(np.float32)
    d2 = noisy_copies(d1, 0.06, 7)
    perm = rng.permutation(400)
    k2, d2 = k2[perm], d2[perm]
    f1, f2 = O.OracleFrame(k1, d1, W, H), O.OracleFrame(k2, d2, W, H)
    has = (rng.random(400) < 0.85).astype(np.uint8)
    nnr = np.float32(0.8)
    n, m21 = O.window_search(f1, f2, has, 30, nnratio=0.8, check_orientation=True)
    # python restatement of ORBmatcher.cc:409-516
    cells, gx, gy = _grid_py(k2)
    exp = np.full(400, -1, np.int64)
    hist = [[] for _ in range(30)]
    for i1 in range(400):
        if not has[i1]:
            continue
        cand = _area_py(k2, cells, gx, gy, k1["x"][i1], k1["y"][i1], 30, k1["octave"][i1], k1["octave"][i1])
        b1 = b2 = 2 ** 31 - 1
        bi = -1
        for i2 in cand:
            if exp[i2] >= 0:
                continue
            dd = _ham(d1[i1], d2[i2])
            if dd < b1:
                b2, b1, bi = b1, dd, i2
            elif dd < b2:
                b2 = dd
        if np.float32(b1) <= np.float32(b2) * nnr and b1 <= 100:
            exp[bi] = i1
            hist[_bin(k1["angle"][i1], k2["angle"][bi])].append(bi)
    keep = _three_max_py([len(h) for h in hist])
    for b in range(30):
        if b not in keep:
            for i2 in hist[b]:
                exp[i2] = -1
    assert np.array_equal(m21, exp) and n == int((exp >= 0).sum()) and n > 150


def test_three_maxima_edge_cases():
    import ctypes as C
    for counts in ([0] * 30, [5] + [0] * 29, [10, 1, 0] + [0] * 27, [10, 10, 10] + [0] * 27, list(range(30)), [3, 3, 2, 2] + [0] * 26):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        arr = np.array(counts, np.int32)
        O.lib().orb_oracle_three_maxima(arr.ctypes.data_as(C.c_void_p), 30, C.byref(a), C.byref(b), C.byref(c))
        assert [a.value, b.value, c.value] == _three_max_py(counts)


def test_knn2_first_minimum_and_second():
    q = random_descriptors(20, 1)
    db = random_descriptors(300, 2)
    db[17] = q[3]
    db[200] = q[3]  # duplicate minimum: the first one must win, second-best = 0
    bd, bi, sd = O.knn2(q, db)
    assert bi[3] == 17 and bd[3] == 0 and sd[3] == 0
    for i in range(20):
        ds = sorted((_ham(q[i], db[j]), j) for j in range(300))
        assert bd[i] == ds[0][0] and bi[i] == ds[0][1] and sd[i] == ds[1][0]


def test_search_by_bow_against_python_loop():
    """Oracle M9 (both overloads) vs a direct Python restatement of ORBmatcher.cc:155-284 / :715-850."""
    rng = np.random.default_rng(3)
    n1, n2 = 300, 280
    d1 = random_descriptors(n1, 5)
    perm = rng.permutation(n1)[:n2]
    d2 = noisy_copies(d1[perm], 0.05, 6)
    a1 = rng.uniform(0, 360, n1).astype(np.float32)
    a2 = ((a1[perm] + rng.normal(5, 3, n2)) % 360).astype(np.float32)
    node1 = (d1[:, 0].astype(np.int32) >> 4) * 5 + 1
    node2 = (d2[:, 0].astype(np.int32) >> 4) * 5 + 1
    node2[rng.random(n2) < 0.05] = 777

    def fv(nodes):
        order = np.argsort(nodes, kind="stable")
        ids, cnt = np.unique(nodes, return_counts=True)
        return ids.astype(np.int32), np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32), order.astype(np.int32)

    fv1, fv2 = fv(node1), fv(node2)
    v1 = (rng.random(n1) < 0.8).astype(np.uint8)
    v2 = (rng.random(n2) < 0.9).astype(np.uint8)
    nnr = np.float32(0.75)
    for variant in (0, 1):
        n, out = O.search_by_bow(variant, d1, v1, a1, fv1, d2, v2, a2, fv2, nnratio=0.75, check_orientation=True)
        exp = np.full(n2 if variant == 0 else n1, -1, np.int64)
        matched2 = np.zeros(n2, bool)
        hist = [[] for _ in range(30)]
        m1 = {int(i): fv1[2][fv1[1][k]:fv1[1][k + 1]] for k, i in enumerate(fv1[0])}
        m2 = {int(i): fv2[2][fv2[1][k]:fv2[1][k + 1]] for k, i in enumerate(fv2[0])}
        for node in sorted(set(m1) & set(m2)):
            for i1 in m1[node]:
                if not v1[i1]:
                    continue
                b1 = b2 = 2 ** 31 - 1
                bi = -1
                for i2 in m2[node]:
                    if variant == 0:
                        if exp[i2] >= 0:
                            continue
                    elif matched2[i2] or not v2[i2]:
                        continue
                    dd = _ham(d1[i1], d2[i2])
                    if dd < b1:
                        b2, b1, bi = b1, dd, i2
                    elif dd < b2:
                        b2 = dd
                ok = (b1 <= 50) if variant == 0 else (b1 < 50)
                if ok and np.float32(b1) < nnr * np.float32(b2):
                    if variant == 0:
                        exp[bi] = i1
                        hist[_bin(a1[i1], a2[bi])].append(bi)
                    else:
                        exp[i1] = bi
                        matched2[bi] = True
                        hist[_bin(a1[i1], a2[bi])].append(i1)
        keep = _three_max_py([len(h) for h in hist])
        for b in range(30):
            if b not in keep:
                for i in hist[b]:
                    exp[i] = -1
        assert np.array_equal(out, exp) and n == int((exp >= 0).sum()) and n > 60


def test_guided_search_rule0_against_python_loop():
    """Oracle guided search (rule 0, histogram filter) vs Python: the skeleton of ORBmatcher.cc:1547-1617."""
    k1, d1 = _rand_frame(300, 11)
    k2 = k1.copy()
    rng = np.random.default_rng(12)
    k2["x"] = (k1["x"] + rng.uniform(-3, 3, 300)).astype(np.float32)
    k2["y"] = (k1["y"] + rng.uniform(-3, 3, 300)).astype(np.float32)
    k2["angle"] = ((k1["angle"] + rng.normal(4, 3, 300)) % 360).astype(np.float32)
    d2 = noisy_copies(d1, 0.07, 13)
    f2 = O.OracleFrame(k2, d2, W, H)
    qr = np.full(300, 12.0, np.float32)
    lo, hi = (k1["octave"] - 1).astype(np.int32), (k1["octave"] + 1).astype(np.int32)
    occ = np.full(300, -1, np.int32)
    occ[::17] = 4
    n, so = O.guided_search(f2, k1["x"], k1["y"], qr, lo, hi, d1, k1["angle"], 0, 0.9, 100, 1, slot_owner=occ)
    cells, gx, gy = _grid_py(k2)
    exp = occ.astype(np.int64).copy()
    hist = [[] for _ in range(30)]
    for q in range(300):
        cand = _area_py(k2, cells, gx, gy, k1["x"][q], k1["y"][q], 12.0, lo[q], hi[q])
        b1, bi = 2 ** 31 - 1, -1
        for i2 in cand:
            if exp[i2] >= 0:
                continue
            dd = _ham(d1[q], d2[i2])
            if dd < b1:
                b1, bi = dd, i2
        if b1 <= 100:
            exp[bi] = q
            hist[_bin(k1["angle"][q], k2["angle"][bi])].append(bi)
    keep = _three_max_py([len(h) for h in hist])
    for b in range(30):
        if b not in keep:
            for i in hist[b]:
                exp[i] = -1
    assert np.array_equal(so, exp) and n > 150


def _pair(n, seed, jitter=4.0, dang=6.0, flip=0.06, levels=4):
    """Two small random frames: the second is a jittered, re-ordered, noisy copy of the first."""
    k1, d1 = _rand_frame(n, seed)
    k1["octave"] = np.random.default_rng(seed + 100).integers(0, levels, n)
    rng = np.random.default_rng(seed + 1)
    k2 = k1.copy()
    k2["x"] = (k1["x"] + rng.uniform(-jitter, jitter, n)).astype(np.float32)
    k2["y"] = (k1["y"] + rng.uniform(-jitter, jitter, n)).astype(np.float32)
    k2["angle"] = ((k1["angle"] + rng.normal(dang, 3, n)) % 360).astype(np.float32)
    d2 = noisy_copies(d1, flip, seed + 2)
    perm = rng.permutation(n)
    return k1, d1, k2[perm], d2[perm], rng


def test_search_for_initialization_against_python_loop():
    """ORBmatcher.cc:598-713: level-0 features only, candidates around the previously matched position, a better match
    steals an already matched F2 feature (vMatchedDistance), rotation filter, prev-matched update."""
    k1, d1, k2, d2, rng = _pair(500, 21, levels=2)
    f1, f2 = O.OracleFrame(k1, d1, W, H), O.OracleFrame(k2, d2, W, H)
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
    nnr = np.float32(0.9)
    for rounds in range(2):
        n, m12, prev_out = O.search_for_initialization(f1, f2, prev, 40, nnratio=0.9, check_orientation=True)
        cells, gx, gy = _grid_py(k2)
        e12 = np.full(len(k1), -1, np.int64)
        e21 = np.full(len(k2), -1, np.int64)
        mdist = np.full(len(k2), 2 ** 31 - 1, np.int64)
        hist = [[] for _ in range(30)]
        nm = 0
        for i1 in range(len(k1)):
            if k1["octave"][i1] > 0:
                continue
            cand = _area_py(k2, cells, gx, gy, prev[i1, 0], prev[i1, 1], 40, 0, 0)
            b1 = b2 = 2 ** 31 - 1
            bi = -1
            for i2 in cand:
                dd = _ham(d1[i1], d2[i2])
                if mdist[i2] <= dd:
                    continue
                if dd < b1:
                    b2, b1, bi = b1, dd, i2
                elif dd < b2:
                    b2 = dd
            if b1 <= 50 and np.float32(b1) < np.float32(b2) * nnr:
                if e21[bi] >= 0:
                    e12[e21[bi]] = -1
                    nm -= 1
                e12[i1], e21[bi], mdist[bi] = bi, i1, b1
                nm += 1
                hist[_bin(k1["angle"][i1], k2["angle"][bi])].append(i1)
        keep = _three_max_py([len(h) for h in hist])
        for b in range(30):
            if b not in keep:
                for i1 in hist[b]:
                    if e12[i1] >= 0:
                        e12[i1] = -1
                        nm -= 1
        exp_prev = prev.copy()
        for i1 in range(len(k1)):
            if e12[i1] >= 0:
                exp_prev[i1] = (k2["x"][e12[i1]], k2["y"][e12[i1]])
        assert np.array_equal(m12, e12) and n == nm and n > 60
        assert np.array_equal(prev_out, exp_prev)
        prev = prev_out  # Tracking::Initialize feeds the updated positions back in


def test_guided_search_rules_1_2_against_python_loop():
    """The ratio rules of the guided skeleton: rule 1 (ORBmatcher.cc:469, :586) and rule 2 (:113-121, with the
    best/second-best LEVELS of the Frame-vs-map-points search)."""
    k1, d1, k2, d2, rng = _pair(350, 31, jitter=5.0)
    f2 = O.OracleFrame(k2, d2, W, H)
    qr = np.full(len(k1), 14.0, np.float32)
    lo, hi = (k1["octave"] - 1).astype(np.int32), k1["octave"].astype(np.int32)
    occ = np.full(len(k2), -1, np.int32)
    occ[::23] = 9
    cells, gx, gy = _grid_py(k2)
    for rule, nnr in ((1, 0.8), (2, 0.8), (1, 0.6), (2, 0.95)):
        n, so = O.guided_search(f2, k1["x"], k1["y"], qr, lo, hi, d1, k1["angle"], rule, nnr, 0, 0, slot_owner=occ)
        exp = occ.astype(np.int64).copy()
        r = np.float32(nnr)
        for q in range(len(k1)):
            b1 = b2 = 2 ** 31 - 1
            bi, l1, l2 = -1, -1, -1
            for i2 in _area_py(k2, cells, gx, gy, k1["x"][q], k1["y"][q], 14.0, lo[q], hi[q]):
                if exp[i2] >= 0:
                    continue
                dd = _ham(d1[q], d2[i2])
                if dd < b1:
                    b2, b1, l2, l1, bi = b1, dd, l1, int(k2["octave"][i2]), i2
                elif dd < b2:
                    l2, b2 = int(k2["octave"][i2]), dd
            if rule == 1:
                ok = np.float32(b1) <= np.float32(b2) * r and b1 <= 100
            else:
                ok = b1 <= 100 and not (l1 == l2 and np.float32(b1) > r * np.float32(b2))
            if ok:
                exp[bi] = q
        assert np.array_equal(so, exp), (rule, nnr)
        assert n == int((exp >= 0).sum() - (occ >= 0).sum()) and n > 100


def test_search_by_projection_frames_against_python_loop():
    """ORBmatcher.cc:1507-1620 (Current vs Last with a pose): projection in cv::gemm's arithmetic (double accumulation,
    one rounding), octave window, free-slot best match <= TH_HIGH, rotation filter."""
    k1, d1, k2, d2, rng = _pair(400, 41, jitter=2.0)
    FX = FY = np.float32(500.0)
    CX, CY, Z = np.float32(W / 2), np.float32(H / 2), np.float32(4.0)
    world = np.empty((len(k1), 3), np.float32)
    world[:, 0] = (k1["x"] - CX) / FX * Z
    world[:, 1] = (k1["y"] - CY) / FY * Z
    world[:, 2] = Z
    T = np.zeros((3, 4), np.float32)
    T[0, 0] = T[1, 1] = T[2, 2] = 1
    T[0, 3], T[1, 3], T[2, 3] = 0.011, -0.007, 0.05
    has = (rng.random(len(k1)) < 0.9).astype(np.uint8)
    outl = (rng.random(len(k1)) < 0.05).astype(np.uint8)
    pre = np.full(len(k2), -1, np.int32)
    pre[::19] = 3
    fc, fl = O.OracleFrame(k2, d2, W, H), O.OracleFrame(k1, d1, W, H)
    for ori in (True, False):
        n, mp = O.search_by_projection_ff(fc, fl, has, outl, world, T, float(FX), float(FY), float(CX), float(CY), 7.0, ori, cur_mp=pre)
        cells, gx, gy = _grid_py(k2)
        exp = pre.astype(np.int64).copy()
        hist = [[] for _ in range(30)]
        nm = 0
        sf = fc.sf
        for i in range(len(k1)):
            if not has[i] or outl[i]:
                continue
            xc = [np.float32(np.float64(T[k, 0]) * np.float64(world[i, 0]) + np.float64(T[k, 1]) * np.float64(world[i, 1]) +
                             np.float64(T[k, 2]) * np.float64(world[i, 2]) + np.float64(T[k, 3])) for k in range(3)]
            invz = np.float32(1.0 / np.float64(xc[2]))
            u = FX * xc[0] * invz + CX
            v = FY * xc[1] * invz + CY
            if u < 0 or u > W or v < 0 or v > H:
                continue
            oc = int(k1["octave"][i])
            radius = np.float32(7.0) * sf[oc]
            b1, bi = 2 ** 31 - 1, -1
            for i2 in _area_py(k2, cells, gx, gy, u, v, radius, oc - 1, oc + 1):
                if exp[i2] >= 0:
                    continue
                dd = _ham(d1[i], d2[i2])
                if dd < b1:
                    b1, bi = dd, i2
            if b1 <= 100:
                exp[bi] = i
                nm += 1
                hist[_bin(k1["angle"][i], k2["angle"][bi])].append(bi)
        if ori:
            keep = _three_max_py([len(h) for h in hist])
            for b in range(30):
                if b not in keep:
                    for i2 in hist[b]:
                        exp[i2] = -1
                        nm -= 1
        assert np.array_equal(mp, exp), ori
        assert n == nm and n > 150


def _project_py(T, X, fx, fy, cx, cy):
    """x3Dc = Rcw*x3Dw + tcw (gemm: double accumulation, one rounding); invz = 1.0/z in double; u, v in float."""
    xc = [np.float32(np.float64(T[k, 0]) * np.float64(X[0]) + np.float64(T[k, 1]) * np.float64(X[1]) +
                     np.float64(T[k, 2]) * np.float64(X[2]) + np.float64(T[k, 3])) for k in range(3)]
    invz = np.float32(1.0 / np.float64(xc[2]))
    return np.float32(fx) * xc[0] * invz + np.float32(cx), np.float32(fy) * xc[1] * invz + np.float32(cy)


def test_search_local_points_against_python_loop():
    """ORBmatcher.cc:49-125 (Frame vs local-map points with cached projections)."""
    k1, d1, k2, d2, rng = _pair(400, 51, jitter=2.5)
    n1 = len(k1)
    f2 = O.OracleFrame(k2, d2, W, H)
    in_view = (rng.random(n1) < 0.9).astype(np.uint8)
    proj = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
    level = k1["octave"].astype(np.int32)
    vcos = np.where(rng.random(n1) < 0.5, 0.9995, 0.9).astype(np.float32)
    occ = np.full(len(k2), -1, np.int32)
    occ[::29] = 5
    cells, gx, gy = _grid_py(k2)
    for th, nnr in ((1.0, 0.8), (3.0, 0.7)):
        n, mp = O.search_local_points(f2, in_view, proj, level, vcos, d1, th, nnratio=nnr, f_mp=occ)
        exp = occ.astype(np.int64).copy()
        nm = 0
        for i in range(n1):
            if not in_view[i]:
                continue
            r = np.float32(2.5) if vcos[i] > 0.998 else np.float32(4.0)
            if th != 1.0:
                r = np.float32(r * np.float32(th))
            lv = int(level[i])
            b1 = b2 = 2 ** 31 - 1
            bi, l1, l2 = -1, -1, -1
            for i2 in _area_py(k2, cells, gx, gy, proj[i, 0], proj[i, 1], np.float32(r * f2.sf[lv]), lv - 1, lv):
                if exp[i2] >= 0:
                    continue
                dd = _ham(d1[i], d2[i2])
                if dd < b1:
                    b2, b1, l2, l1, bi = b1, dd, l1, int(k2["octave"][i2]), i2
                elif dd < b2:
                    l2, b2 = int(k2["octave"][i2]), dd
            if b1 <= 100:
                if l1 == l2 and np.float32(b1) > np.float32(nnr) * np.float32(b2):
                    continue
                exp[bi] = i
                nm += 1
        assert np.array_equal(mp, exp) and n == nm and n > 100, (th, nnr)


def test_search_by_projection_kf_against_python_loop():
    """ORBmatcher.cc:1622-1746 (relocalisation refinement): level predicted from dist3D / minDistance with lower_bound."""
    k1, d1, k2, d2, rng = _pair(400, 61, jitter=2.0)
    n1 = len(k1)
    FX, FY, CX, CY, Z = 500.0, 500.0, W / 2, H / 2, np.float32(4.0)
    world = np.empty((n1, 3), np.float32)
    world[:, 0] = (k1["x"] - np.float32(CX)) / np.float32(FX) * Z
    world[:, 1] = (k1["y"] - np.float32(CY)) / np.float32(FY) * Z
    world[:, 2] = Z
    T = np.zeros((3, 4), np.float32)
    T[0, 0] = T[1, 1] = T[2, 2] = 1
    T[0, 3], T[1, 3], T[2, 3] = -0.012, 0.004, 0.2
    min_dist = (Z / np.float32(1.2) ** k1["octave"].astype(np.float32) * rng.uniform(0.8, 1.1, n1)).astype(np.float32)
    valid = (rng.random(n1) < 0.85).astype(np.uint8)
    f2 = O.OracleFrame(k2, d2, W, H)
    cells, gx, gy = _grid_py(k2)
    # Ow = -Rcw.t()*tcw (gemm with alpha = -1: double accumulation, scaled, one rounding)
    Ow = [np.float32((np.float64(T[0, k]) * np.float64(T[0, 3]) + np.float64(T[1, k]) * np.float64(T[1, 3]) +
                      np.float64(T[2, k]) * np.float64(T[2, 3])) * -1.0) for k in range(3)]
    for th, od, ori in ((10.0, 100, True), (4.0, 64, False)):
        n, mp = O.search_by_projection_kf(f2, valid, world, min_dist, d1, k1["angle"], T, FX, FY, CX, CY, th, od, ori)
        exp = np.full(len(k2), -1, np.int64)
        hist = [[] for _ in range(30)]
        nm = 0
        for i in range(n1):
            if not valid[i]:
                continue
            u, v = _project_py(T, world[i], FX, FY, CX, CY)
            if u < 0 or u > W or v < 0 or v > H:
                continue
            PO = [np.float32(world[i, k] - Ow[k]) for k in range(3)]
            d3 = np.float32(np.sqrt(np.float64(PO[0]) * np.float64(PO[0]) + np.float64(PO[1]) * np.float64(PO[1]) + np.float64(PO[2]) * np.float64(PO[2])))
            ratio = np.float32(d3 / min_dist[i])
            lv = min(int(np.searchsorted(f2.sf, ratio, side="left")), len(f2.sf) - 1)
            radius = np.float32(np.float32(th) * f2.sf[lv])
            b1, bi = 2 ** 31 - 1, -1
            for i2 in _area_py(k2, cells, gx, gy, u, v, radius, lv - 1, lv + 1):
                if exp[i2] >= 0:
                    continue
                dd = _ham(d1[i], d2[i2])
                if dd < b1:
                    b1, bi = dd, i2
            if b1 <= od:
                exp[bi] = i
                nm += 1
                hist[_bin(k1["angle"][i], k2["angle"][bi])].append(bi)
        if ori:
            keep = _three_max_py([len(h) for h in hist])
            for b in range(30):
                if b not in keep:
                    for i2 in hist[b]:
                        exp[i2] = -1
                        nm -= 1
        assert np.array_equal(mp, exp) and n == nm and n > 40, (th, od, ori)


def test_search_by_projection_f1f2_against_python_loop():
    """ORBmatcher.cc:519-594 (F1 map points projected into F2, same-octave window, ratio test)."""
    k1, d1, k2, d2, rng = _pair(400, 71, jitter=3.0)
    n1 = len(k1)
    FX, FY, CX, CY, Z = 500.0, 500.0, W / 2, H / 2, np.float32(4.0)
    world = np.empty((n1, 3), np.float32)
    world[:, 0] = (k1["x"] - np.float32(CX)) / np.float32(FX) * Z
    world[:, 1] = (k1["y"] - np.float32(CY)) / np.float32(FY) * Z
    world[:, 2] = Z
    T = np.zeros((3, 4), np.float32)
    T[0, 0] = T[1, 1] = T[2, 2] = 1
    T[0, 3] = 0.01
    valid = (rng.random(n1) < 0.8).astype(np.uint8)
    occ = np.full(len(k2), -1, np.int32)
    occ[::31] = 2 ** 31 - 1
    f1, f2 = O.OracleFrame(k1, d1, W, H), O.OracleFrame(k2, d2, W, H)
    cells, gx, gy = _grid_py(k2)
    for win, nnr in ((15, 0.9), (40, 0.7)):
        n, mp = O.search_by_projection_f1f2(f1, f2, valid, world, T, FX, FY, CX, CY, win, nnratio=nnr, f2_mp=occ)
        exp = occ.astype(np.int64).copy()
        nm = 0
        for i in range(n1):
            if not valid[i]:
                continue
            u, v = _project_py(T, world[i], FX, FY, CX, CY)
            lv = int(k1["octave"][i])
            b1 = b2 = 2 ** 31 - 1
            bi = -1
            for i2 in _area_py(k2, cells, gx, gy, u, v, win, lv, lv):
                if exp[i2] >= 0:
                    continue
                dd = _ham(d1[i], d2[i2])
                if dd < b1:
                    b2, b1, bi = b1, dd, i2
                elif dd < b2:
                    b2 = dd
            if np.float32(b1) <= np.float32(b2) * np.float32(nnr) and b1 <= 100:
                exp[bi] = i
                nm += 1
        assert np.array_equal(mp, exp) and n == nm and n > 80, (win, nnr)


def _fv(nodes):
    """FeatureVector as CSR (ids ascending, feature indices ascending inside a node)."""
    ids = np.unique(nodes)
    ptr, items = [0], []
    for n in ids:
        items += list(np.nonzero(nodes == n)[0])
        ptr.append(len(items))
    return ids.astype(np.int32), np.array(ptr, np.int32), np.array(items, np.int32)


def test_search_for_triangulation_against_python_loop():
    """ORBmatcher.cc:852-1032 + CheckDistEpipolarLine :136-153: per shared vocabulary node, candidates <= TH_LOW sorted by
    (distance, index), the first one within round(2*best) that satisfies the epipolar constraint wins."""
    k1, d1, k2, d2, rng = _pair(500, 81, jitter=3.0, flip=0.04)
    k2["y"] = (k2["y"] - np.float32(1.0)).astype(np.float32)
    has1 = (rng.random(len(k1)) < 0.4).astype(np.uint8)
    has2 = (rng.random(len(k2)) < 0.4).astype(np.uint8)
    node1 = (d1[:, 3].astype(np.int32) >> 5) * 7 + 2
    node2 = (d2[:, 3].astype(np.int32) >> 5) * 7 + 2
    node2[rng.random(len(k2)) < 0.05] = 999
    fv1, fv2 = _fv(node1), _fv(node2)
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 1.0]], np.float32)   # epipolar lines y2 = y1 - 1
    sigma2 = (np.float32(1.2) ** np.arange(8, dtype=np.float32)) ** 2
    f32 = np.float32
    for ori in (True, False):
        n, m12 = O.search_for_triangulation(k1, d1, has1, fv1, k2, d2, has2, fv2, F12, sigma2, check_orientation=ori)
        exp = np.full(len(k1), -1, np.int64)
        matched2 = np.zeros(len(k2), bool)
        hist = [[] for _ in range(30)]
        nm = 0
        n2 = {int(i): fv2[2][fv2[1][q]:fv2[1][q + 1]] for q, i in enumerate(fv2[0])}
        for q, nid in enumerate(fv1[0]):
            if int(nid) not in n2:
                continue
            for idx1 in fv1[2][fv1[1][q]:fv1[1][q + 1]]:
                if has1[idx1]:
                    continue
                cand = []
                for idx2 in n2[int(nid)]:
                    if matched2[idx2] or has2[idx2]:
                        continue
                    dd = _ham(d1[idx1], d2[idx2])
                    if dd > 50:
                        continue
                    cand.append((dd, int(idx2)))
                if not cand:
                    continue
                cand.sort()
                dist_th = 2 * cand[0][0]
                for dd, idx2 in cand:
                    if dd > dist_th:
                        break
                    x1, y1, x2, y2 = f32(k1["x"][idx1]), f32(k1["y"][idx1]), f32(k2["x"][idx2]), f32(k2["y"][idx2])
                    a = f32(f32(x1 * F12[0, 0]) + f32(y1 * F12[1, 0])) + F12[2, 0]
                    b = f32(f32(x1 * F12[0, 1]) + f32(y1 * F12[1, 1])) + F12[2, 1]
                    c = f32(f32(x1 * F12[0, 2]) + f32(y1 * F12[1, 2])) + F12[2, 2]
                    num = f32(f32(f32(a * x2) + f32(b * y2)) + c)
                    den = f32(f32(a * a) + f32(b * b))
                    if den == 0:
                        continue
                    dsqr = f32(f32(num * num) / den)
                    if np.float64(dsqr) < 3.84 * np.float64(sigma2[int(k2["octave"][idx2])]):
                        matched2[idx2] = True
                        exp[idx1] = idx2
                        nm += 1
                        hist[_bin(k1["angle"][idx1], k2["angle"][idx2])].append(idx1)
                        break
        if ori:
            keep = _three_max_py([len(h) for h in hist])
            for bb in range(30):
                if bb not in keep:
                    for i1 in hist[bb]:
                        exp[i1] = -1
                        nm -= 1
        assert np.array_equal(m12, exp) and n == nm and n > 40, ori


def test_guided_best_against_python_loop():
    """The slot-free best-candidate search shared by Fuse / SearchBySim3 (level filter [lo, hi], best distance <= th)."""
    k1, d1, k2, d2, rng = _pair(350, 91, jitter=4.0)
    f2 = O.OracleFrame(k2, d2, W, H)
    qr = (np.float32(9.0) * np.float32(1.2) ** k1["octave"]).astype(np.float32)
    lo, hi = (k1["octave"] - 1).astype(np.int32), k1["octave"].astype(np.int32)
    cells, gx, gy = _grid_py(k2)
    for th in (50, 100):
        best = O.guided_best(f2, k1["x"], k1["y"], qr, lo, hi, d1, th)
        exp = np.full(len(k1), -1, np.int64)
        for q in range(len(k1)):
            b1, bi = 2 ** 31 - 1, -1
            for i2 in _area_py(k2, cells, gx, gy, k1["x"][q], k1["y"][q], qr[q], -1, -1):   # KeyFrame::GetFeaturesInArea: no filter
                if k2["octave"][i2] < lo[q] or k2["octave"][i2] > hi[q]:
                    continue
                dd = _ham(d1[q], d2[i2])
                if dd < b1:
                    b1, bi = dd, i2
            if b1 <= th:
                exp[q] = bi
        assert np.array_equal(best, exp) and (exp >= 0).sum() > 100, th
