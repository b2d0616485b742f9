"""Deterministic synthetic u8 frames for tests and bench (numpy only; no reference code involved).

Multi-octave smoothed noise plus random filled rectangles: gives many more FAST corners per cell than
the per-cell quota so that retention, the threshold-7 fallback (flat regions) and the level-wide trim are
all exercised.
"""
import numpy as np


def _box_blur(a, r):
    if r <= 0:
        return a
    k = 2 * r + 1
    c = np.cumsum(np.pad(a, ((0, 0), (r + 1, r)), mode="reflect"), axis=1, dtype=np.float64)
    a = (c[:, k:] - c[:, :-k]) / k
    c = np.cumsum(np.pad(a, ((r + 1, r), (0, 0)), mode="reflect"), axis=0, dtype=np.float64)
    return (c[k:, :] - c[:-k, :]) / k


def textured_frame(width, height, seed=0, n_rects=None, flat_band=True):
    rng = np.random.default_rng(seed)
    img = np.zeros((height, width), np.float64)
    for octave, (r, amp) in enumerate(((0, 10.0), (1, 30.0), (3, 45.0), (8, 50.0), (20, 40.0))):
        n = rng.standard_normal((height, width))
        n = _box_blur(n, r)
        n /= (n.std() + 1e-9)
        img += amp * n
    img = 128.0 + img * (60.0 / img.std())
    n_rects = n_rects if n_rects is not None else max(8, (width * height) // 6000)
    for _ in range(n_rects):
        w = int(rng.integers(4, max(6, width // 12)))
        h = int(rng.integers(4, max(6, height // 12)))
        x = int(rng.integers(0, max(1, width - w)))
        y = int(rng.integers(0, max(1, height - h)))
        img[y:y + h, x:x + w] = rng.integers(0, 256)
    if flat_band:
        # a low-contrast band: cells here yield <=3 corners at th=20 and fall back to th=7
        y0 = height // 3
        band = slice(y0, y0 + max(8, height // 6))
        img[band, : width // 2] = 100.0 + 0.12 * (img[band, : width // 2] - 128.0)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def shifted_frame(frame, dx, dy, seed=0, noise=2.0):
    """A translated copy with a little sensor noise: consecutive frames of a stream."""
    rng = np.random.default_rng(seed)
    out = np.roll(frame, (dy, dx), axis=(0, 1)).astype(np.float64)
    out += noise * rng.standard_normal(out.shape)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def random_descriptors(n, seed=0):
    return np.random.default_rng(seed).integers(0, 256, size=(n, 32), dtype=np.uint8)


def noisy_copies(desc, flip_prob, seed=0):
    """Descriptors with independent random bit flips (realistic best/second-best ratios)."""
    rng = np.random.default_rng(seed)
    bits = np.unpackbits(desc, axis=1)
    flips = rng.random(bits.shape) < flip_prob
    return np.packbits(bits ^ flips, axis=1)


def random_vocabulary(k=10, L=3, seed=0, flip=0.12, stop_fraction=0.02, ragged=False):
    """A synthetic DBoW2-style vocabulary tree as flat arrays (see include/orbfe_bow.h): complete k-ary tree of depth L
    whose node descriptors are noisy copies of their parent's (so that descents are decided by small margins and ties
    occur), word ids in leaf order, positive idf-like weights with a few stopped (zero-weight) words.  ragged=True
    prunes some subtrees so that leaves sit at different depths."""
    rng = np.random.default_rng(seed)
    desc = [np.zeros(32, np.uint8)]
    children_of = [[]]
    level_of = [0]
    frontier = [0]
    for lev in range(1, L + 1):
        nxt = []
        for parent in frontier:
            if ragged and lev > 1 and rng.random() < 0.15:
                continue  # this node stays a leaf
            base = rng.integers(0, 256, 32, dtype=np.uint8) if parent == 0 else desc[parent]
            for _ in range(k):
                bits = np.unpackbits(base)
                bits ^= (rng.random(256) < flip).astype(np.uint8)
                desc.append(np.packbits(bits))
                children_of.append([])
                level_of.append(lev)
                children_of[parent].append(len(desc) - 1)
                nxt.append(len(desc) - 1)
        frontier = nxt
    n = len(desc)
    child_ptr = np.zeros(n + 1, np.int32)
    children = []
    for i in range(n):
        children += children_of[i]
        child_ptr[i + 1] = len(children)
    word_id = np.full(n, -1, np.int32)
    weight = np.zeros(n, np.float64)
    w = 0
    for i in range(n):
        if not children_of[i] and i != 0:
            word_id[i] = w
            w += 1
            weight[i] = 0.0 if rng.random() < stop_fraction else float(rng.uniform(0.5, 9.0))
    return {"node_desc": np.stack(desc).astype(np.uint8), "child_ptr": child_ptr, "children": np.array(children, np.int32),
            "word_id": word_id, "weight": weight, "L": L, "k": k}


def random_keyframe_db(nkf=300, nwords=6000, words_per_kf=300, seed=0, loop_at=40):
    """A synthetic keyframe database for the KeyFrameDatabase rows: keyframes along a trajectory see a sliding window of a
    long random word sequence (neighbours share most words), a stretch near `loop_at` is revisited by the query.
    Returns dict(kf_ptr, db_ids, db_vals, covis_ptr, covis, connected, q_ids, q_vals)."""
    rng = np.random.default_rng(seed)
    track = rng.integers(0, nwords, nkf * 40 + words_per_kf * 2)

    def bow_at(pos, noise_seed):
        r = np.random.default_rng(noise_seed)
        w = track[pos:pos + words_per_kf].copy()
        flip = r.random(len(w)) < 0.25
        w[flip] = r.integers(0, nwords, int(flip.sum()))
        ids, cnt = np.unique(w, return_counts=True)
        vals = cnt.astype(np.float64) * r.uniform(0.5, 3.0, len(ids))
        return ids.astype(np.int32), vals / vals.sum()

    kf_ptr, ids_all, vals_all = [0], [], []
    for k in range(nkf):
        ids, vals = bow_at(40 * k, 1000 + k)
        ids_all.append(ids); vals_all.append(vals)
        kf_ptr.append(kf_ptr[-1] + len(ids))
    covis_ptr, covis = [0], []
    for k in range(nkf):
        nb = [k + d for d in (1, -1, 2, -2, 3, -3, 4, -4, 5, -5) if 0 <= k + d < nkf]
        rng.shuffle(nb)
        covis += nb[:int(rng.integers(0, 11))]
        covis_ptr.append(len(covis))
    connected = np.zeros(nkf, np.uint8)
    connected[max(0, nkf - 6):] = 1                      # the query keyframe follows the last ones
    q_ids, q_vals = bow_at(40 * loop_at + 7, 99)          # ... and revisits the place of keyframe loop_at
    return {"kf_ptr": np.array(kf_ptr, np.int32), "db_ids": np.concatenate(ids_all), "db_vals": np.concatenate(vals_all),
            "covis_ptr": np.array(covis_ptr, np.int32), "covis": np.array(covis, np.int32), "connected": connected,
            "q_ids": q_ids, "q_vals": q_vals}
