"""ctypes mirror of include/orbfe_comm.h (tests and bench.py only): NCCL communicator, descriptor-block all-gather, sharded
keyframe-database sweep and the rig exchange fused into the extractor.  The communicator id travels between the ranks
through torch.distributed (any out-of-band channel would do); the data path itself never touches torch or the host."""
import ctypes as C

import numpy as np

from . import OrbfeError, lib

ID_BYTES = 128
_bound = False


def _bind():
    global _bound
    L = lib()
    if _bound:
        return L
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    L.orbfe_comm_unique_id.argtypes = [vp]
    L.orbfe_comm_create.argtypes = [vp, i, i, i, C.POINTER(vp)]
    L.orbfe_comm_destroy.argtypes = [vp]
    L.orbfe_comm_world.argtypes = [vp]
    L.orbfe_comm_rank.argtypes = [vp]
    L.orbfe_comm_sync.argtypes = [vp]
    L.orbfe_comm_barrier.argtypes = [vp, vp]
    L.orbfe_allgather_desc.argtypes = [vp, vp, vp, vp, i, i, vp, vp, vp, vp]
    L.orbfe_comm_broadcast.argtypes = [vp, vp, sz, i, vp]
    L.orbfe_comm_allgather.argtypes = [vp, vp, vp, sz, vp]
    L.orbfe_shard_range.argtypes = [i, i, i, C.POINTER(i), C.POINTER(i)]
    L.orbfe_knn2_sweep_sharded.argtypes = [vp, vp, vp, i, i, vp, i, i, vp, vp, vp, vp, vp]
    L.orbfe_rig_exchange_create.argtypes = [vp, i, i, C.POINTER(vp)]
    L.orbfe_rig_exchange_destroy.argtypes = [vp]
    L.orbfe_extract_batch_device_exchange.argtypes = [vp, vp, i, i, sz, sz, i, vp, vp]
    L.orbfe_search_for_initialization_exchange.argtypes = [vp, vp, i, vp, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, i, C.c_float, i,
                                                           vp, vp, vp]
    L.orbfe_rig_exchange_wait.argtypes = [vp, vp]
    L.orbfe_rig_exchange_release.argtypes = [vp, vp]
    L.orbfe_rig_exchange_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.orbfe_rig_exchange_check.argtypes = [vp, vp]
    L.orbfe_rig_exchange_bytes.argtypes = [vp]
    L.orbfe_rig_exchange_bytes.restype = sz
    _bound = True
    return L


def _check(rc):
    if rc != 0:
        raise OrbfeError(rc, lib().orbfe_last_error().decode("utf-8", "replace"))


def shard_range(n_items, world, rank):
    lo, hi = C.c_int(0), C.c_int(0)
    _check(_bind().orbfe_shard_range(n_items, world, rank, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def nccl_version():
    return _bind().orbfe_comm_nccl_version()


class Comm:
    """OrbfeComm.  create(dist, device): rank 0 makes the NCCL id, torch.distributed ships it (plumbing only)."""

    def __init__(self, handle, world, rank):
        self.h, self.world, self.rank = handle, world, rank

    @classmethod
    def create(cls, torch, dist, device_index):
        L = _bind()
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        ident = np.zeros(ID_BYTES, np.uint8)
        if rank == 0:
            _check(L.orbfe_comm_unique_id(ident.ctypes.data))
        if world > 1:
            t = torch.from_numpy(ident).to("cuda:%d" % device_index) if dist.get_backend() == "nccl" else torch.from_numpy(ident)
            dist.broadcast(t, 0)
            ident = t.cpu().numpy().copy()
        h = C.c_void_p()
        _check(L.orbfe_comm_create(ident.ctypes.data, world, rank, device_index, C.byref(h)))
        return cls(h, world, rank)

    def close(self):
        if self.h:
            _bind().orbfe_comm_destroy(self.h)
            self.h = None

    def sync(self):
        _check(_bind().orbfe_comm_sync(self.h))

    def barrier(self, stream=0):
        _check(_bind().orbfe_comm_barrier(self.h, C.c_void_p(stream)))

    def allgather_desc(self, d_kps, d_desc, d_counts, cap, nslots, d_all_kps, d_all_desc, d_all_counts, stream=0):
        vp = C.c_void_p
        _check(_bind().orbfe_allgather_desc(self.h, vp(d_kps), vp(d_desc), vp(d_counts), cap, nslots, vp(d_all_kps), vp(d_all_desc),
                                            vp(d_all_counts), vp(stream)))

    def broadcast(self, d_buf, nbytes, root=0, stream=0):
        _check(_bind().orbfe_comm_broadcast(self.h, C.c_void_p(d_buf), nbytes, root, C.c_void_p(stream)))

    def knn2_sweep_sharded(self, matcher, d_query, nq, root, d_db_shard, ngroups_total, group_size, d_best, d_idx, d_second, d_scratch, stream=0):
        vp = C.c_void_p
        _check(_bind().orbfe_knn2_sweep_sharded(self.h, matcher.handle, vp(d_query), nq, root, vp(d_db_shard), ngroups_total, group_size,
                                                vp(d_best), vp(d_idx), vp(d_second), vp(d_scratch), vp(stream)))


class RigExchange:
    """OrbfeRigExchange: the extractor's descriptor kernel stores straight into every rank's gather buffer (NVLink peer
    stores) and publishes an epoch flag; wait() / release() are one tiny kernel each."""

    def __init__(self, comm, cap, nslots):
        self.comm, self.cap, self.nslots = comm, cap, nslots
        self.h = C.c_void_p()
        _check(_bind().orbfe_rig_exchange_create(comm.h, cap, nslots, C.byref(self.h)))

    def close(self):
        if self.h:
            _bind().orbfe_rig_exchange_destroy(self.h)
            self.h = None

    def extract(self, extractor, d_imgs, W, H, stride, frame_stride, stream=0):
        _check(_bind().orbfe_extract_batch_device_exchange(extractor._h, C.c_void_p(d_imgs), W, H, stride, frame_stride, self.nslots,
                                                           self.h, C.c_void_p(stream)))

    def search_for_initialization(self, matcher, npairs, d_f1_idx, d_f2_idx, d_prev_matched, width, height, window, d_match12, d_nmatches,
                                  stream=0):
        """orbfe_search_for_initialization_exchange: the matcher waits for the epoch's data and releases it itself."""
        vp = C.c_void_p
        _check(_bind().orbfe_search_for_initialization_exchange(matcher.handle, self.h, npairs, vp(d_f1_idx), vp(d_f2_idx), vp(d_prev_matched),
                                                                0.0, 0.0, float(width), float(height), int(window), float(matcher.mfNNratio),
                                                                int(matcher.mbCheckOrientation), vp(d_match12), vp(d_nmatches), vp(stream)))

    def wait(self, stream=0):
        _check(_bind().orbfe_rig_exchange_wait(self.h, C.c_void_p(stream)))

    def release(self, stream=0):
        _check(_bind().orbfe_rig_exchange_release(self.h, C.c_void_p(stream)))

    def buffers(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(_bind().orbfe_rig_exchange_buffers(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def buffers_of_current_epoch(self):
        """Gathered views of the epoch just PRODUCED (before any wait): only this rank's own slot is valid without waiting."""
        L = _bind()
        L.orbfe_rig_exchange_buffers_produced.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(L.orbfe_rig_exchange_buffers_produced(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def check(self, stream=0):
        _check(_bind().orbfe_rig_exchange_check(self.h, C.c_void_p(stream)))

    def bytes_pushed(self):
        return int(_bind().orbfe_rig_exchange_bytes(self.h))
