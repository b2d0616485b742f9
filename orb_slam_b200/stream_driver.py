"""ctypes binding of tools/e2e_driver.cpp (orb_slam_b200/libe2e_driver.so): bench.py's end-to-end stream pipeline on C++
threads over the PUBLIC C-ABI of liborbfe.so (orbfe_extract_batch with pinned host frames, orbfe_search_by_projection_frames
on host views).  Bench / test harness, not part of the product library."""
import ctypes as C

import numpy as np

from . import lib
from .build import build_e2e_driver

_FN_NAMES = ["orbfe_extractor_create", "orbfe_extractor_destroy", "orbfe_extract_batch", "orbfe_extractor_last_launches",
             "orbfe_matcher_create", "orbfe_matcher_destroy", "orbfe_matcher_counters", "orbfe_search_by_projection_frames",
             "orbfe_frame_scale_factors", "orbfe_last_error"]


class E2eConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("W", "H", "nfeat", "nlevels", "fast_th", "B", "NB", "nex", "nmatch", "nbuf", "device")] + \
               [(n, C.c_float) for n in ("scale", "fx", "fy", "cx", "cy", "depth", "th")]


class StreamDriver:
    """frames_ptr: NB*B pinned host frames (W x H u8, contiguous); Tcws: (NB*B, 12) float32; out_*_ptrs: nbuf pinned host output
    sets (B x nfeat keypoints of 28 bytes, B x nfeat x 32 descriptor bytes, B int32 counts)."""

    def __init__(self, W, H, nfeat, nlevels, scale, fast_th, B, NB, nex, nmatch, device, fx, fy, cx, cy, depth, th,
                 frames_ptr, Tcws, out_kps_ptrs, out_desc_ptrs, out_cnt_ptrs):
        dl = C.CDLL(build_e2e_driver())
        dl.e2e_create.restype = C.c_void_p
        dl.e2e_create.argtypes = [C.POINTER(E2eConfig)] + [C.c_void_p] * 6
        dl.e2e_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
        dl.e2e_last_matches.argtypes = [C.c_void_p, C.c_void_p]
        dl.e2e_destroy.argtypes = [C.c_void_p]
        dl.e2e_error.restype = C.c_char_p
        dl.e2e_error.argtypes = [C.c_void_p]
        L = lib()
        nbuf = len(out_kps_ptrs)
        assert nbuf >= nex + 1 and len(out_desc_ptrs) == nbuf and len(out_cnt_ptrs) == nbuf
        self._fns = (C.c_void_p * len(_FN_NAMES))(*[C.cast(getattr(L, n), C.c_void_p).value for n in _FN_NAMES])
        self._cfg = E2eConfig(W, H, nfeat, nlevels, fast_th, B, NB, nex, nmatch, nbuf, device, scale, fx, fy, cx, cy, depth, th)
        self._Tcws = np.ascontiguousarray(Tcws, np.float32)
        assert self._Tcws.shape == (NB * B, 12)
        arr = lambda ps: (C.c_void_p * nbuf)(*[int(p) for p in ps])
        self._dl, self.nfeat, self.nbuf = dl, nfeat, nbuf
        self._h = dl.e2e_create(C.byref(self._cfg), self._fns, int(frames_ptr), self._Tcws.ctypes.data, arr(out_kps_ptrs), arr(out_desc_ptrs),
                                arr(out_cnt_ptrs))
        if not self._h:
            raise RuntimeError("e2e_create failed: " + L.orbfe_last_error().decode("utf-8", "replace"))

    def run(self, nbatches):
        """Processes the next `nbatches` batches of the stream.  Returns a dict: keypoints, matches, extractor kernel launches,
        matcher H2D / D2H bytes and launches, index (in batches since the driver was created) of the last matched batch, and the
        summed host seconds inside the extract calls, the view / map-point glue and the matcher calls."""
        out = (C.c_longlong * 7)()
        hs = (C.c_double * 3)()
        if self._dl.e2e_run(self._h, int(nbatches), out, hs) != 0:
            raise RuntimeError("e2e driver failed: " + self._dl.e2e_error(self._h).decode("utf-8", "replace"))
        return {"keypoints": out[0], "matches": out[1], "extract_launches": out[2], "match_h2d": out[3], "match_d2h": out[4],
                "match_launches": out[5], "last_batch": out[6], "extract_s": hs[0], "views_s": hs[1], "match_s": hs[2]}

    def last_matches(self):
        """Match vectors (4 x nfeat int32) of the first four pairs of the last matched batch."""
        mp4 = np.empty((4, self.nfeat), np.int32)
        if self._dl.e2e_last_matches(self._h, mp4.ctypes.data) != 0:
            raise RuntimeError("no batch matched yet")
        return mp4

    def close(self):
        if self._h:
            self._dl.e2e_destroy(self._h)
            self._h = None
