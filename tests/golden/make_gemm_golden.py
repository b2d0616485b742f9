"""Generates tests/golden/opencv_gemm.npz: inputs and python-cv2 outputs of cv2.gemm / cv2.norm for the matrix shapes the
reference's matchers use (ORBmatcher.cc:302,323,542,1286,1330-1331,1530,1628,1648; Frame.cc:134; KeyFrame.cc:73-105).
Pins the float semantics of `Rcw*x3Dw+tcw` and friends: OpenCV sums 3x3*3x1, 3x3*3x3, 4x4*4x4 products in FLOAT (the
unrolled small-matrix branch of matmul.cpp) and everything else (e.g. -Rcw.t()*tcw) in double.
Run here (python-cv2 4.13 is in the image): python tests/golden/make_gemm_golden.py"""
import os

import cv2
import numpy as np

rng = np.random.default_rng(20260923)
cases = []   # (name, A shape, B shape, with C, alpha, beta, flags)
for name, sa, sb, wc, al, be, fl in [
        ("R*x+t", (3, 3), (3, 1), True, 1.0, 1.0, 0), ("-sR*t", (3, 3), (3, 1), False, -1.0, 0.0, 0),
        ("R*R", (3, 3), (3, 3), False, 1.0, 0.0, 0), ("K*T34", (3, 3), (3, 4), False, 1.0, 0.0, 0),
        ("T*T", (4, 4), (4, 4), False, 1.0, 0.0, 0), ("-Rt*t", (3, 3), (3, 1), False, -1.0, 0.0, cv2.GEMM_1_T),
        ("row*col", (1, 3), (3, 1), False, 1.0, 0.0, 0), ("3x4*4x1", (3, 4), (4, 1), False, 1.0, 0.0, 0)]:
    n = 400
    A = rng.standard_normal((n,) + sa).astype(np.float32) * rng.choice([0.01, 1.0, 50.0], (n, 1, 1)).astype(np.float32)
    B = rng.standard_normal((n,) + sb).astype(np.float32) * rng.choice([0.1, 1.0, 20.0], (n, 1, 1)).astype(np.float32)
    rows = sa[1] if fl & cv2.GEMM_1_T else sa[0]
    Cm = rng.standard_normal((n, rows, sb[1])).astype(np.float32) if wc else None
    out = np.stack([cv2.gemm(A[i], B[i], al, None if Cm is None else Cm[i], be, flags=fl) for i in range(n)])
    cases.append((name, A, B, Cm, al, be, fl, out))
v = rng.standard_normal((400, 3)).astype(np.float32) * 10
norms = np.array([cv2.norm(x.reshape(3, 1)) for x in v])
d = {"names": np.array([c[0] for c in cases]), "norm_in": v, "norm_out": norms, "cv2_version": np.array(cv2.__version__)}
for k, (name, A, B, Cm, al, be, fl, out) in enumerate(cases):
    d["A%d" % k], d["B%d" % k], d["out%d" % k] = A, B, out
    d["C%d" % k] = Cm if Cm is not None else np.zeros(0, np.float32)
    d["p%d" % k] = np.array([al, be, fl], np.float64)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "opencv_gemm.npz"), **d)
print("wrote opencv_gemm.npz with", len(cases), "cases, cv2", cv2.__version__)
