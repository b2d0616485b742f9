"""One small extraction + one matcher call: the workload `compute-sanitizer --tool racecheck` is run on (tools/gpu_check.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import orb_slam_b200 as fe
from orb_slam_b200.synth import textured_frame
img = textured_frame(752, 480, seed=2)
ex = fe.ORBextractor(1500, 1.2, 8)
k, d = ex(img)
k2, d2 = ex(img)
assert np.array_equal(d, d2) and np.array_equal(k["x"], k2["x"])
print("racecheck workload ok:", len(k), "keypoints")
ex.close()
