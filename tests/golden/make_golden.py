#!/usr/bin/env python3
"""Generates tests/golden/*.npz: golden vectors for the OpenCV primitives the reference calls
(cv::resize, cv::FAST, the integer GaussianBlur engine, cv::fastAtan2, cv::copyMakeBorder), produced by
python cv2 (opencv_python_headless 4.13.0) in the build container.  The reference itself ships no golden
vectors and cannot be compiled here (needs OpenCV 2.4 C++ / ROS / Boost), so these pin the oracle's
restatement of the un-vendored dependency; see oracle/orb_oracle.h "PARITY PIN STATUS".

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from orb_slam_b200.synth import textured_frame  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
cv2.setNumThreads(1)


def fast_list(img, th):
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(img)
    a = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps], np.int32).reshape(-1, 3)
    return a


def main():
    img = textured_frame(320, 240, seed=5)
    rng = np.random.default_rng(9)
    noise = rng.integers(0, 256, (96, 128), dtype=np.uint8)
    g = {"img": img, "noise": noise}
    # resize chain with the reference's level sizes for 320x240, scale 1.2 (cvRound of float32 products)
    sizes = [(320, 240), (267, 200), (222, 167), (185, 139)]
    prev = img
    for i, (w, h) in enumerate(sizes[1:], 1):
        prev = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
        g["resize_%d" % i] = prev
    g["resize_noise_107x80"] = cv2.resize(noise, (107, 80), interpolation=cv2.INTER_LINEAR)
    g["border16"] = cv2.copyMakeBorder(img, 16, 16, 16, 16, cv2.BORDER_REFLECT_101)
    for th in (20, 7):
        g["fast_img_th%d" % th] = fast_list(img, th)
        g["fast_noise_th%d" % th] = fast_list(noise, th)
    # a cell-like sub-image with odd origin (cv::FAST on a Mat ROI)
    g["fast_roi_th20"] = fast_list(np.ascontiguousarray(img[13:13 + 75, 29:29 + 111]), 20)
    k = np.array([18, 34, 49, 55, 49, 34, 18], np.float64) / 256.0
    g["blur_img"] = cv2.sepFilter2D(img, cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101)
    g["blur_noise"] = cv2.sepFilter2D(noise, cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101)
    ys = rng.integers(-120000, 120000, 4000).astype(np.float32)
    xs = rng.integers(-120000, 120000, 4000).astype(np.float32)
    ys[:8] = [0, 0, 5, 0, -5, 3, -3, 1]
    xs[:8] = [0, 5, 0, -5, 0, 3, 3, -1]
    g["atan2_y"], g["atan2_x"] = ys, xs
    g["atan2_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    np.savez_compressed(os.path.join(OUT, "opencv_primitives.npz"), **g)
    print("wrote", os.path.join(OUT, "opencv_primitives.npz"), "cv2", cv2.__version__)

    # cv::undistortPoints(pts, K, D, R = I, P = K) as Frame::UndistortKeyPoints calls it (reference src/Frame.cc:289-319):
    # a separate small file so that the older fixture stays byte-identical
    u = {}
    K = np.array([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], np.float32)
    u["K"] = K
    pts = np.stack([rng.uniform(0, 640, 3000), rng.uniform(0, 480, 3000)], axis=1).astype(np.float32)
    pts[:4] = [[0, 0], [640, 0], [0, 480], [640, 480]]  # the corners ComputeImageBounds undistorts
    u["pts"] = pts
    coeffs = np.array([[0.2624, -0.9531, -0.0054, 0.0026, 1.1633],     # TUM fr1 (strong)
                       [-0.28, 0.07, 0.0002, 0.00002, 0.0],            # 4-coefficient model (ORB-SLAM settings files)
                       [0.1, 0.0, 0.0, 0.0, 0.0]], np.float32)
    u["coeffs"] = coeffs
    for i, D in enumerate(coeffs):
        u["out_%d" % i] = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    np.savez_compressed(os.path.join(OUT, "opencv_undistort.npz"), **u)
    print("wrote", os.path.join(OUT, "opencv_undistort.npz"))


if __name__ == "__main__":
    main()
