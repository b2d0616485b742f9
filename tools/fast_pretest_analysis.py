#!/usr/bin/env python3
"""Why fast_nms_tma_kernel evaluates the exact 9-arc network for every pixel ("commit the measured pass
fraction per tile at t=20 vs t=7").  CPU analysis on the bench's own frames (bench.make_stream, seed 11) with the oracle's
threshold-free score map m (a pixel is a FAST-9/16 corner at t  <=>  m > t):

  corner      fraction of pixels with m > t (what an exact scorer must be run on, at the very least)
  compass     necessary test on the 4 compass ring pixels (two adjacent ones all brighter / all darker by more than t)
  opp8        necessary test on the 8 opposite ring pairs (every 9-arc contains one pixel of each pair)
  opp4        the same on 4 of the pairs

Cost model (thread instructions on the ALU pipe per pixel, u16x2-packed where possible): the exact network is 40 /px;
opp8 costs 14 /px, compaction ~1 /px, the exact score of a compacted pixel ~60 (scalar gather + network; it no longer
shares the packed ring registers of its neighbour).  A two-pass schedule (everything at fastTh, threshold 7 only for the
cells that fell back, ORBextractor.cc:609-614) therefore costs 15 + pass*60 per pixel: at the measured pass fractions
(19-24 % at t=20, 29-38 % at t=7) that is 26-29 /px at t=20 and 32-38 /px at t=7 against 40 /px -- a 10 % gain on the
whole kernel (87 /px all told) in the best case, for a second pixel pass over the fallback cells, a scalar scorer that cannot use
VIMNMX3.U16x2 at full rate and divergent queue code.  The synthetic stream is corner-dense by construction (SURVEY 8d asks
for many more corners per cell than the quota): 14-18 % of ALL pixels are corners at t=20.  On such input FAST is bound by
the integer ALU pipe, not by HBM, and a pre-test does not change that; the kernel stays exact-everywhere."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402  (analysis tool: not part of the product)
from bench import make_stream, level_sizes  # noqa: E402

RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def shifted(img, dx, dy):
    h, w = img.shape
    return img[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx].astype(np.int16)


def stats(img):
    m = O.fast_m_map(img)[3:-3, 3:-3].astype(np.int16)
    v = shifted(img, 0, 0)
    r = [shifted(img, dx, dy) for dx, dy in RING]
    out = {}
    for t in (7, 20):
        corner = m > t
        c = [r[0], r[4], r[8], r[12]]
        br, dk = [x > v + t for x in c], [x < v - t for x in c]
        comp = np.zeros_like(corner)
        for j in range(4):
            comp |= (br[j] & br[(j + 1) % 4]) | (dk[j] & dk[(j + 1) % 4])
        ob = np.ones_like(corner); od = np.ones_like(corner)
        ob4 = np.ones_like(corner); od4 = np.ones_like(corner)
        for k in range(8):
            b, d = (r[k] > v + t) | (r[k + 8] > v + t), (r[k] < v - t) | (r[k + 8] < v - t)
            ob &= b; od &= d
            if k % 2 == 0:
                ob4 &= b; od4 &= d
        assert not (corner & ~comp).any() and not (corner & ~(ob | od)).any()   # both tests are necessary conditions
        out[t] = (corner.mean(), comp.mean(), (ob | od).mean(), (ob4 | od4).mean())
    return out


if __name__ == "__main__":
    frames, _ = make_stream(40, seed=11)
    print("frame level  w x h      | t=7: corner compass opp8 opp4 | t=20: corner compass opp8 opp4")
    for fi in (0, 5, 37):
        img = frames[fi]
        for l, (w, h) in enumerate(level_sizes()):
            if l:
                img = O.resize_linear(img, w, h)
            s = stats(img)
            print("%5d %5d %5dx%-5d| %s | %s" % (fi, l, w, h, " ".join("%.3f" % x for x in s[7]), " ".join("%.3f" % x for x in s[20])))
