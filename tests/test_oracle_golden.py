"""CPU tests: the oracle against the committed golden vectors (python cv2 outputs for the OpenCV
primitives the reference calls), against SURVEY.md's tabulated constants, and against independent
numpy re-implementations of the reference's own statics."""
import hashlib
import os

import numpy as np
import pytest

import oracle as O
from orb_slam_b200.synth import textured_frame

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "opencv_primitives.npz"))


def test_pattern_table_hash():
    inc = open(os.path.join(os.path.dirname(__file__), "..", "include", "orbfe_brief_pattern.inc")).read()
    body = inc[inc.index("*/") + 2:]
    nums = [int(x) for x in body.replace("\n", " ").split(",") if x.strip()]
    assert len(nums) == 1024 and min(nums) == -13 and max(nums) == 12
    h = hashlib.sha256(",".join(map(str, nums)).encode()).hexdigest()
    assert h == "88df8ca875cc8db56799edd57bb914edad8acb2d48c202b7a464a575b55dbdb8"


def test_ctor_tables():
    p = O.make_params(1000, 1.2, 8, 1, 20)
    assert list(p.quota)[:8] == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(p.umax) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert [O.level_size(p, l, 640, 480) for l in range(8)] == [(640, 480), (533, 400), (444, 333), (370, 278),
                                                                   (309, 231), (257, 193), (214, 161), (179, 134)]
    p = O.make_params(2000, 1.2, 8, 1, 20)
    assert list(p.quota)[:8] == [434, 362, 302, 251, 209, 175, 145, 122]
    assert [O.level_size(p, l, 1920, 1080) for l in range(8)] == [(1920, 1080), (1600, 900), (1333, 750), (1111, 625),
                                                                     (926, 521), (772, 434), (643, 362), (536, 301)]
    # cell grids at 1080p: cols x rows, cellW x cellH, per-cell quota
    exp = [(6, 10, 315, 105, 8), (6, 10, 262, 87, 7), (5, 8, 261, 90, 8), (5, 8, 216, 75, 7),
           (4, 7, 224, 70, 8), (4, 7, 185, 58, 7), (4, 7, 153, 48, 6), (3, 5, 168, 54, 9)]
    for l in range(8):
        w, h = O.level_size(p, l, 1920, 1080)
        rc, g = O.cell_grid(p, l, 1920, 1080, w, h)
        assert rc == 0 and (g.cols, g.rows, g.cell_w, g.cell_h, g.nf_cell) == exp[l]
    p = O.make_params(4000, 1.2, 12, 1, 20)
    assert list(p.quota)[:12] == [751, 626, 521, 435, 362, 302, 251, 210, 175, 146, 121, 100]
    assert O.level_size(p, 11, 3840, 2160) == (517, 291)


def test_resize_golden():
    prev = G["img"]
    for i in (1, 2, 3):
        ref = G["resize_%d" % i]
        prev = O.resize_linear(prev, ref.shape[1], ref.shape[0])
        assert np.array_equal(prev, ref)
    assert np.array_equal(O.resize_linear(G["noise"], 107, 80), G["resize_noise_107x80"])


def test_border_golden():
    assert np.array_equal(O.reflect101_pad(G["img"], 16), G["border16"])


@pytest.mark.parametrize("name,key,th", [("img", "fast_img_th20", 20), ("img", "fast_img_th7", 7),
                                          ("noise", "fast_noise_th20", 20), ("noise", "fast_noise_th7", 7)])
def test_fast_golden(name, key, th):
    xs, ys, sc = O.fast_detect(G[name], th)
    got = np.stack([xs, ys, sc], axis=1).reshape(-1, 3)
    assert np.array_equal(got, G[key])  # same keypoints, same scores, same (raster) order


def test_fast_roi_and_m_definition():
    img = G["img"]
    roi = np.ascontiguousarray(img[13:13 + 75, 29:29 + 111])
    xs, ys, sc = O.fast_detect(roi, 20)
    assert np.array_equal(np.stack([xs, ys, sc], axis=1).reshape(-1, 3), G["fast_roi_th20"])
    # threshold-free formulation used by the CUDA path: score = m-1, corner <=> m > t, NMS on m inside the window
    m = O.fast_m_map(roi)
    h, w = roi.shape
    for th, key in ((20, "fast_roi_th20"),):
        out = []
        for y in range(3, h - 3):
            for x in range(3, w - 3):
                v = m[y, x]
                if v <= th:
                    continue
                nb = m[y - 1:y + 2, x - 1:x + 2].copy()
                nb[1, 1] = -1
                # neighbours outside the detectable window [3,w-3)x[3,h-3) count as 0
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        if not (3 <= y + dy < h - 3 and 3 <= x + dx < w - 3):
                            nb[dy + 1, dx + 1] = 0
                if (v > nb).all():
                    out.append((x, y, v - 1))
        assert np.array_equal(np.array(out, np.int32).reshape(-1, 3), G[key])


def test_blur_golden():
    assert np.array_equal(O.blur7(G["img"]), G["blur_img"])
    assert np.array_equal(O.blur7(G["noise"]), G["blur_noise"])


def test_atan2_golden():
    got = np.array([O.fast_atan2(y, x) for y, x in zip(G["atan2_y"], G["atan2_x"])], np.float32)
    assert np.array_equal(got, G["atan2_deg"])


def test_ic_angle_and_brief_against_numpy():
    """Second, independent implementation (numpy, float32 semantics spelled out) of the two reference statics."""
    pat = np.array([int(x) for x in open(os.path.join(os.path.dirname(__file__), "..", "include", "orbfe_brief_pattern.inc"))
                    .read().split("*/")[1].replace("\n", " ").split(",") if x.strip()], np.int32).reshape(256, 2, 2)
    img = textured_frame(200, 160, seed=4)
    pad = O.reflect101_pad(img, 16).astype(np.int64)
    umax = O.UMAX
    rng = np.random.default_rng(0)
    f32 = np.float32
    for _ in range(40):
        x, y = int(rng.integers(16, 200 - 16)), int(rng.integers(16, 160 - 16))
        m01 = m10 = 0
        for v in range(-15, 16):
            d = umax[abs(v)]
            for u in range(-d, d + 1):
                val = pad[y + 16 + v, x + 16 + u]
                m10 += u * val
                m01 += v * val
        assert (m01, m10) == O.ic_moments(O.reflect101_pad(img, 16), x, y)
        ang = O.ic_angle(O.reflect101_pad(img, 16), None, x, y)
        assert abs(ang - (np.degrees(np.arctan2(m01, m10)) % 360.0)) < 0.02  # fastAtan2 accuracy ~0.01 deg
        th = f32(ang) * f32(np.pi / f32(180.0))
        a, b = f32(np.cos(np.float64(th))), f32(np.sin(np.float64(th)))
        bits = []
        for p in range(256):
            t = []
            for e in range(2):
                px, py = f32(pat[p, e, 0]), f32(pat[p, e, 1])
                ry = int(np.rint(np.float64(f32(f32(px * b) + f32(py * a)))))
                rx = int(np.rint(np.float64(f32(f32(px * a) - f32(py * b)))))
                t.append(pad[y + 16 + ry, x + 16 + rx])
            bits.append(1 if t[0] < t[1] else 0)
        ref = np.packbits(np.array(bits, np.uint8).reshape(32, 8)[:, ::-1], axis=1).ravel()
        got = O.brief(O.reflect101_pad(img, 16), x, y, ang)
        assert np.array_equal(got, ref)


def test_extract_invariants_and_modes():
    img = textured_frame(640, 480, seed=1)
    p = O.make_params(1000, 1.2, 8, 1, 20)
    rc, k, d, dump = O.extract(p, img, want_dump=True)
    assert rc == 0 and len(k) == 1000
    assert dump["n_level_kp"] == [217, 181, 151, 126, 105, 87, 73, 60]
    assert np.all(np.diff(k["octave"]) >= 0) and np.all(k["class_id"] == -1)
    lv0 = k[k["octave"] == 0]
    assert lv0["x"].min() >= 16 and lv0["x"].max() <= 640 - 17 and lv0["y"].min() >= 16 and lv0["y"].max() <= 480 - 17
    assert np.all((k["angle"] >= 0) & (k["angle"] < 360))
    assert set(np.unique(k["size"])) == {np.float32(int(np.float32(31) * s)) for s in list(p.scale)[:8]}
    assert 0.35 < np.unpackbits(d).mean() < 0.65
    # literal std::nth_element retention differs from the canonical rule only inside tie groups
    p2 = O.make_params(1000, 1.2, 8, 1, 20, ties_mode=O.TIES_NTH_ELEMENT)
    rc, k2, d2, _ = O.extract(p2, img)
    s1 = {(a["octave"], a["x"], a["y"]) for a in k}
    s2 = {(a["octave"], a["x"], a["y"]) for a in k2}
    assert len(k2) == 1000 and len(s1 & s2) >= 950
    assert sorted(k["response"].tolist()) == sorted(k2["response"].tolist())  # same multiset of scores
    # cosf/sinf (what the reference TU resolves to) vs correctly rounded: descriptors almost always identical
    p3 = O.make_params(1000, 1.2, 8, 1, 20, trig_mode=O.TRIG_LIBMF)
    rc, k3, d3, _ = O.extract(p3, img)
    assert np.array_equal(k3, k) and (d3 != d).any(axis=1).sum() <= 2


def test_extract_regression_hash():
    """Self-golden of the whole oracle pipeline (catches accidental changes; not a reference pin)."""
    img = textured_frame(320, 240, seed=5)
    p = O.make_params(500, 1.2, 6, 1, 20)
    rc, k, d, _ = O.extract(p, img)
    assert rc == 0
    h = hashlib.sha256(k.tobytes() + d.tobytes()).hexdigest()
    golden = open(os.path.join(os.path.dirname(__file__), "golden", "oracle_extract_320x240.sha256")).read().strip()
    assert h == golden


def test_fallback_cells_and_flat_image():
    img = textured_frame(640, 480, seed=1)
    img[100:260, 0:330] = (100 + 0.05 * (img[100:260, 0:330].astype(np.float32) - 128)).astype(np.uint8)
    p = O.make_params(1000, 1.2, 8, 1, 20)
    rc, k, d, dump = O.extract(p, img, want_dump=True)
    assert rc == 0 and dump["fallback"] > 0
    rc, k, d, _ = O.extract(p, np.full((480, 640), 9, np.uint8))
    assert rc == 0 and len(k) == 0
    rc, k, d, _ = O.extract(p, np.zeros((40, 40), np.uint8))
    assert rc == -2  # degenerate grid: outside the supported domain


def test_undistort_points_vs_cv2_golden():
    """Frame::UndistortKeyPoints' cv::undistortPoints(pts, K, D, I, K): the oracle against python-cv2 vectors, bit-exact."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "opencv_undistort.npz"))
    K = g["K"]
    for i, D in enumerate(g["coeffs"]):
        out = O.undistort_points(g["pts"], K[0, 0], K[1, 1], K[0, 2], K[1, 2], D)
        assert np.array_equal(out.view(np.uint32), g["out_%d" % i].view(np.uint32))
        # ComputeImageBounds from the four corners (the first four fixture points), Frame.cc:336-339
        c = g["out_%d" % i][:4]
        b = O.image_bounds(640, 480, K[0, 0], K[1, 1], K[0, 2], K[1, 2], D)
        exp = [min(np.floor(c[0, 0]), np.floor(c[2, 0])), min(np.floor(c[0, 1]), np.floor(c[1, 1])),
               max(np.ceil(c[1, 0]), np.ceil(c[3, 0])), max(np.ceil(c[2, 1]), np.ceil(c[3, 1]))]
        assert list(b) == exp
    # k1 == 0: keypoints are copied and the bounds are the image (Frame.cc:291-295, :343-348)
    kps = np.zeros(3, O.KP_DTYPE)
    kps["x"], kps["y"], kps["octave"] = [1.5, 2.5, 3.5], [4, 5, 6], [0, 1, 2]
    assert np.array_equal(O.undistort_keypoints(kps, 500, 500, 320, 240, [0, 0.3, 0.1, 0.1, 0]), kps)
    assert list(O.image_bounds(640, 480, 500, 500, 320, 240, [0, 0, 0, 0, 0])) == [0, 0, 640, 480]


def test_extract_composition_against_live_cv2():
    """ComputeKeyPoints (reference src/ORBextractor.cc:522-708) composed in Python from REAL OpenCV calls -- cv2.resize
    for the pyramid, cv2.FastFeatureDetector on every cell window (threshold fallback to 7), the quota redistribution
    loops, retainBest with the canonical tie rule -- against the oracle's extract(): same keypoints, same responses,
    same order.  Orientation and descriptors are cross-checked separately (test_ic_angle_and_brief_against_numpy)."""
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    import math
    f32 = np.float32

    def harris(lv, x, y):
        # HarrisResponses, ORBextractor.cc:79-120 (blockSize 7, k = 0.04): integer gradient sums, float32 finish
        a = b = c = 0
        L = lv.astype(np.int64)
        for i in range(-3, 4):
            for j in range(-3, 4):
                yy, xx = y + i, x + j
                Ix = (L[yy, xx + 1] - L[yy, xx - 1]) * 2 + (L[yy - 1, xx + 1] - L[yy - 1, xx - 1]) + (L[yy + 1, xx + 1] - L[yy + 1, xx - 1])
                Iy = (L[yy + 1, xx] - L[yy - 1, xx]) * 2 + (L[yy + 1, xx - 1] - L[yy - 1, xx - 1]) + (L[yy + 1, xx + 1] - L[yy - 1, xx + 1])
                a += Ix * Ix; b += Iy * Iy; c += Ix * Iy
        fa, fb, fc = f32(a), f32(b), f32(c)
        scale = f32(1.0) / f32(f32(28) * f32(255.0))
        s4 = f32(f32(f32(scale * scale) * scale) * scale)
        return float(f32(f32(f32(fa * fb) - f32(fc * fc)) - f32(f32(f32(0.04) * f32(fa + fb)) * f32(fa + fb))) * s4)

    def compose(img, nfeatures, nlevels, fast_th, score_type=1):
        p = O.make_params(nfeatures, 1.2, nlevels, score_type, fast_th)
        H0, W0 = img.shape
        ratio = f32(W0) / f32(H0)                                     # (float)cols/rows
        out, levels = [], []
        level_img = img
        for level in range(nlevels):
            if level > 0:
                w, h = int(np.rint(f32(W0) * f32(p.inv_scale[level]))), int(np.rint(f32(H0) * f32(p.inv_scale[level])))   # cvRound
                level_img = cv2.resize(level_img, (w, h), interpolation=cv2.INTER_LINEAR)
            h, w = level_img.shape
            nDesired = int(p.quota[level])
            cols = int(math.sqrt(f32(nDesired) / (f32(5) * ratio)))
            rows = int(ratio * f32(cols))
            minB, maxBX, maxBY = 16, w - 16, h - 16
            Wd, Hd = maxBX - minB, maxBY - minB
            cellW, cellH = int(math.ceil(f32(Wd) / f32(cols))), int(math.ceil(f32(Hd) / f32(rows)))
            nCells = rows * cols
            nfCell = int(math.ceil(f32(nDesired) / f32(nCells)))
            cells = [[[] for _ in range(cols)] for _ in range(rows)]
            nToRetain = [[0] * cols for _ in range(rows)]
            nTotal = [[0] * cols for _ in range(rows)]
            noMore = [[False] * cols for _ in range(rows)]
            iniX, iniY = [0] * cols, [0] * rows
            nNoMore = nToDistribute = 0
            hY = cellH + 6
            for i in range(rows):
                y0 = minB + i * cellH - 3
                iniY[i] = y0
                if i == rows - 1:
                    hY = maxBY + 3 - y0
                    if hY <= 0:
                        continue
                hX = cellW + 6
                for j in range(cols):
                    x0 = minB + j * cellW - 3
                    iniX[j] = x0
                    if j == cols - 1:
                        hX = maxBX + 3 - x0
                        if hX <= 0:
                            continue
                    roi = np.ascontiguousarray(level_img[y0:y0 + hY, x0:x0 + hX])
                    det = cv2.FastFeatureDetector_create(threshold=fast_th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
                    kps = det.detect(roi)
                    if len(kps) <= 3:
                        det = cv2.FastFeatureDetector_create(threshold=7, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
                        kps = det.detect(roi)
                    cells[i][j] = [(int(k.pt[0]), int(k.pt[1]),
                                    float(k.response) if score_type == 1 else harris(level_img, x0 + int(k.pt[0]), y0 + int(k.pt[1]))) for k in kps]
                    n = len(kps)
                    nTotal[i][j] = n
                    if n > nfCell:
                        nToRetain[i][j] = nfCell
                    else:
                        nToRetain[i][j] = n
                        nToDistribute += nfCell - n
                        noMore[i][j] = True
                        nNoMore += 1
            while nToDistribute > 0 and nNoMore < nCells:
                nNew = nfCell + int(math.ceil(f32(nToDistribute) / f32(nCells - nNoMore)))
                nToDistribute = 0
                for i in range(rows):
                    for j in range(cols):
                        if not noMore[i][j]:
                            if nTotal[i][j] > nNew:
                                nToRetain[i][j] = nNew
                            else:
                                nToRetain[i][j] = nTotal[i][j]
                                nToDistribute += nNew - nTotal[i][j]
                                noMore[i][j] = True
                                nNoMore += 1
            level_kps = []
            for i in range(rows):
                for j in range(cols):
                    # retainBest + resize(n): the n largest responses, ties at the cut by earlier raster position (canonical);
                    # survivors in raster order inside the cell (canonical output order)
                    c = sorted(cells[i][j], key=lambda t: (-t[2], t[1], t[0]))[:nToRetain[i][j]]
                    c.sort(key=lambda t: (t[1], t[0]))
                    level_kps += [(x + iniX[j], y + iniY[i], r, (i, j)) for x, y, r in c]
            if len(level_kps) > nDesired:
                # level-wide retainBest: ties at the cut by cell row-major order, then raster; order of the survivors kept
                order = sorted(range(len(level_kps)), key=lambda q: (-level_kps[q][2], q))[:nDesired]
                level_kps = [level_kps[q] for q in sorted(order)]
            out += [(x, y, level, r) for x, y, r, _ in level_kps]
            levels.append(level_img)
        return p, out, levels

    from orb_slam_b200.synth import textured_frame as tf
    for (W_, H_, nf, nl, th, seed, st) in ((640, 480, 1000, 8, 20, 1, 1), (752, 480, 600, 6, 12, 4, 1), (320, 240, 300, 4, 20, 9, 1),
                                           (400, 300, 400, 5, 20, 12, 0)):   # the last one: HARRIS_SCORE
        img = tf(W_, H_, seed=seed)
        p, exp, levels = compose(img, nf, nl, th, st)
        rc, ok, od, _ = O.extract(p, img)
        assert rc == 0 and len(ok) == len(exp), (len(ok), len(exp))
        sc = [f32(p.scale[l]) for l in range(nl)]
        ex = np.array([f32(x) * sc[l] if l else f32(x) for x, y, l, r in exp], np.float32)
        ey = np.array([f32(y) * sc[l] if l else f32(y) for x, y, l, r in exp], np.float32)
        assert np.array_equal(ok["x"], ex) and np.array_equal(ok["y"], ey)
        assert np.array_equal(ok["octave"], np.array([l for _, _, l, _ in exp]))
        assert np.array_equal(ok["response"], np.array([r for _, _, _, r in exp], np.float32))
        # orientation and descriptor of every keypoint, again from real OpenCV pieces: copyMakeBorder (the 16-px reflect-101
        # frame), sepFilter2D with the 2.4 integer taps (the smoothed interior), cv2.fastAtan2, and the rotated pattern
        # sampled in numpy with the reference's float32 expressions
        pat = np.array([int(x) for x in open(os.path.join(os.path.dirname(__file__), "..", "include", "orbfe_brief_pattern.inc"))
                        .read().split("*/")[1].replace("\n", " ").split(",") if x.strip()], np.int32).reshape(512, 2)
        taps = np.array([18, 34, 49, 55, 49, 34, 18], np.float64) / 256.0
        umax = O.UMAX
        frames = []
        for lv in levels:
            raw = cv2.copyMakeBorder(lv, 16, 16, 16, 16, cv2.BORDER_REFLECT_101)
            work = raw.copy()
            work[16:-16, 16:-16] = cv2.sepFilter2D(lv, cv2.CV_8U, taps, taps, borderType=cv2.BORDER_REFLECT_101)
            frames.append((raw.astype(np.int64), work.astype(np.int64)))
        px, py = pat[:, 0].astype(np.float32), pat[:, 1].astype(np.float32)
        us = [np.arange(-umax[abs(v)], umax[abs(v)] + 1) for v in range(-15, 16)]
        for q, (x, y, l, r) in enumerate(exp):
            raw, work = frames[l]
            m10 = m01 = 0
            for v, u in zip(range(-15, 16), us):
                row = raw[y + 16 + v, x + 16 + u]
                m10 += int((u * row).sum())
                m01 += v * int(row.sum())
            ang = f32(cv2.fastAtan2(float(m01), float(m10)))
            assert ang == ok["angle"][q], (q, ang, ok["angle"][q])
            th_ = f32(ang) * f32(np.pi / f32(180.0))
            a, b = f32(np.cos(np.float64(th_))), f32(np.sin(np.float64(th_)))
            ry = np.rint((px * b + py * a).astype(np.float64)).astype(np.int64)    # float32 products and sum, cvRound
            rx = np.rint((px * a - py * b).astype(np.float64)).astype(np.int64)
            vals = work[y + 16 + ry, x + 16 + rx].reshape(256, 2)
            bits = (vals[:, 0] < vals[:, 1]).astype(np.uint8)
            ref = np.packbits(bits.reshape(32, 8)[:, ::-1], axis=1).ravel()
            assert np.array_equal(od[q], ref), q


def test_levels_with_an_empty_cell_grid():
    """ORBextractor.cc:533-547: levelCols == 0 (quota below 5*imageRatio) leaves the level's cell vectors empty; the
    reference yields no keypoints there and keeps going.  The oracle must do the same (it used to reject the input)."""
    from orb_slam_b200.synth import textured_frame
    img = textured_frame(640, 480, seed=3)
    for nf, want in ((100, [22, 18, 15, 13, 10, 9, 7, 0]), (60, [13, 11, 9, 8, 0, 0, 0, 0]), (20, [0] * 8)):
        p = O.make_params(nf, 1.2, 8, 1, 20)
        rc, k, d, _ = O.extract(p, img)
        assert rc == 0 and list(np.bincount(k["octave"], minlength=8)) == want
        g = O.cell_grid(p, 7, 640, 480, *O.level_size(p, 7, 640, 480))
        assert g is not None
