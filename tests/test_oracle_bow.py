"""CPU tests of the oracle's restatement of the DBoW2 transform and of MapPoint::ComputeDistinctiveDescriptors against
independent pure-Python statements of the same reference loops (dict-based, i.e. literally the std::map semantics)."""
import math

import numpy as np

import oracle as O
from orb_slam_b200.synth import random_descriptors, noisy_copies, random_vocabulary, random_keyframe_db


def _ham(a, b):
    return int(np.unpackbits(a ^ b).sum())


def _py_transform(voc, desc, levelsup, weighting, norm):
    """TemplatedVocabulary.h:1126-1260 with dicts standing in for the std::maps."""
    L = voc["L"]
    cp, ch = voc["child_ptr"], voc["children"]
    bow, fv = {}, {}
    leafs, nids = [], []
    for i, f in enumerate(desc):
        nid_level = L - levelsup
        nid, cur, level = 0, 0, 0
        while True:
            level += 1
            nodes = ch[cp[cur]:cp[cur + 1]]
            cur = int(nodes[0])
            best = _ham(f, voc["node_desc"][cur])
            for c in nodes[1:]:
                d = _ham(f, voc["node_desc"][c])
                if d < best:
                    best, cur = d, int(c)
            if level == nid_level:
                nid = cur
            if cp[cur + 1] == cp[cur]:
                break
        leafs.append(cur)
        nids.append(nid)
        w = float(voc["weight"][cur])
        wid = int(voc["word_id"][cur])
        if w > 0:
            if weighting in (0, 1):
                bow[wid] = bow.get(wid, 0.0) + w if wid in bow else w
            else:
                bow.setdefault(wid, w)
            fv.setdefault(nid, []).append(i)
    ids = sorted(bow)
    vals = [bow[k] for k in ids]
    if weighting in (0, 1) and vals and norm == 0:
        vals = [v / float(len(vals)) for v in vals]
    if norm:
        s = 0.0
        for v in vals:
            s += abs(v) if norm == 1 else v * v
        if norm == 2:
            s = math.sqrt(s)
        if s > 0:
            vals = [v / s for v in vals]
    return leafs, nids, ids, vals, fv


def test_bow_transform_oracle_vs_python():
    for seed, (k, L, ragged) in enumerate([(10, 3, False), (4, 5, True), (3, 2, False)]):
        voc = random_vocabulary(k, L, seed=seed, ragged=ragged)
        words = voc["node_desc"][voc["word_id"] >= 0]
        rng = np.random.default_rng(seed)
        desc = noisy_copies(words[rng.integers(0, len(words), 300)], 0.08, seed + 10)
        for levelsup, weighting, norm in [(1, 0, 1), (2, 1, 0), (L, 2, 2), (L + 2, 3, 1), (0, 0, 2)]:
            leaf, node = O.bow_descend(voc, desc, levelsup)
            (ids, vals), (fids, fptr, ffeat) = O.bow_transform(voc, desc, levelsup, weighting, norm)
            pl, pn, pids, pvals, pfv = _py_transform(voc, desc, levelsup, weighting, norm)
            assert list(leaf) == pl and list(node) == pn
            assert list(ids) == pids
            assert np.array_equal(vals, np.array(pvals, np.float64))  # same additions in the same order: bit-equal
            assert list(fids) == sorted(pfv)
            for kk, nid in enumerate(fids):
                assert list(ffeat[fptr[kk]:fptr[kk + 1]]) == pfv[int(nid)]
            if norm == 1 and len(vals):
                assert abs(vals.sum() - 1.0) < 1e-12


def test_distinctive_oracle_vs_python():
    rng = np.random.default_rng(3)
    groups, ptr = [], [0]
    for g in range(60):
        n = int(rng.integers(0, 40)) if g % 7 else int(rng.integers(0, 3))
        base = random_descriptors(1, 100 + g)
        groups.append(noisy_copies(np.repeat(base, n, axis=0), rng.uniform(0.02, 0.3), 200 + g) if n else np.zeros((0, 32), np.uint8))
        ptr.append(ptr[-1] + n)
    desc = np.concatenate(groups)
    best = O.distinctive_descriptors(desc, np.array(ptr, np.int32))
    for g in range(60):
        d = desc[ptr[g]:ptr[g + 1]]
        N = len(d)
        if N == 0:
            assert best[g] == -1
            continue
        bm, bi = 2 ** 31 - 1, 0
        for i in range(N):
            row = sorted(_ham(d[i], d[j]) for j in range(N))
            med = row[int(0.5 * (N - 1))]
            if med < bm:
                bm, bi = med, i
        assert best[g] == bi


def _py_db_detect(mode, db, min_score):
    """KeyFrameDatabase.cc:73-308 with Python lists/dicts standing in for the inverted file and the keyframe fields."""
    f32 = np.float32
    kf_ptr, ids, vals = db["kf_ptr"], db["db_ids"], db["db_vals"]
    nkf = len(kf_ptr) - 1
    inv = {}
    for k in range(nkf):
        for w in ids[kf_ptr[k]:kf_ptr[k + 1]]:
            inv.setdefault(int(w), []).append(k)
    query, words, score, sharing = set(), {}, {}, []
    for w in db["q_ids"]:
        for k in inv.get(int(w), []):
            if k not in query:
                words[k] = 0
                if mode == 1 or not db["connected"][k]:
                    query.add(k)
                    sharing.append(k)
            words[k] += 1
    if not sharing:
        return []
    max_common = max(words[k] for k in sharing)
    min_common = int(f32(max_common) * f32(0.8))
    qd = dict(zip(db["q_ids"].tolist(), db["q_vals"].tolist()))
    lsm = []
    for k in sharing:
        if words[k] > min_common:
            s = 0.0
            for w, v in zip(ids[kf_ptr[k]:kf_ptr[k + 1]].tolist(), vals[kf_ptr[k]:kf_ptr[k + 1]].tolist()):
                if w in qd:
                    s += abs(qd[w] - v) - abs(qd[w]) - abs(v)
            si = f32(-s / 2.0)
            score[k] = si
            if mode == 1 or si >= f32(min_score):
                lsm.append((si, k))
    if not lsm:
        return []
    acc = []
    best_acc = f32(min_score) if mode == 0 else f32(0)
    for si, k in lsm:
        best_s, acc_s, best_k = si, si, k
        for k2 in db["covis"][db["covis_ptr"][k]:db["covis_ptr"][k + 1]]:
            k2 = int(k2)
            if k2 not in query:
                continue
            if mode == 0 and not words[k2] > min_common:
                continue
            s2 = score.get(k2, f32(0))
            acc_s = f32(acc_s + s2)
            if s2 > best_s:
                best_k, best_s = k2, s2
        acc.append((acc_s, best_k))
        if acc_s > best_acc:
            best_acc = acc_s
    keep = f32(0.75) * best_acc
    out = []
    for a, k in acc:
        if a > keep and k not in out:
            out.append(k)
    return out


def test_db_detect_oracle_vs_python():
    for seed in range(4):
        db = random_keyframe_db(nkf=120 + 30 * seed, nwords=3000, words_per_kf=200, seed=seed, loop_at=20 + seed)
        for mode, ms in ((0, 0.0), (0, 0.05), (1, 0.0)):
            cand, common, score = O.bow_db_detect(mode, db["q_ids"], db["q_vals"], db["kf_ptr"], db["db_ids"], db["db_vals"],
                                                  db["connected"], db["covis_ptr"], db["covis"], ms)
            assert list(cand) == _py_db_detect(mode, db, ms), (seed, mode, ms)
            assert len(cand) > 0
