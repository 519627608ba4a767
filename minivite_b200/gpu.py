"""ctypes binding of libmvgpu.so (include/mvgpu.h) + the host-side mirror of the reference seam.

`dist_louvain_method(me, nprocs, shard, lower, thresh)` has the argument meaning of the reference's
`distLouvainMethod` (dspl.hpp:1280-1283): it returns (modularity, iters).  There is NO CPU fallback:
if the CUDA library is missing or no GPU is usable every call raises.
"""
import ctypes
import os

import numpy as np

_LIB_PATH = os.environ.get("MVGPU_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmvgpu.so")
UNIQUE_ID_BYTES = 128

TRACE_DTYPE = np.dtype([("modularity", "<f8"), ("moved", "<i8"), ("chash", "<u8")])


class Timings(ctypes.Structure):
    _fields_ = [("total_s", ctypes.c_double), ("setup_s", ctypes.c_double), ("scan_s", ctypes.c_double),
                ("fold_s", ctypes.c_double), ("exchange_s", ctypes.c_double), ("h2d_s", ctypes.c_double),
                ("scan_launches", ctypes.c_int64), ("kernel_launches", ctypes.c_int64),
                ("iters", ctypes.c_int32), ("unit_weight", ctypes.c_int32), ("reorder_s", ctypes.c_double),
                ("reordered", ctypes.c_int32), ("scan_kernel_chosen", ctypes.c_int32), ("h2d_bytes", ctypes.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ["mvgpu_last_error", "mvgpu_device_count", "mvgpu_create", "mvgpu_destroy", "mvgpu_get_unique_id",
           "mvgpu_comm_init", "mvgpu_upload_shard", "mvgpu_attach_shard_device", "mvgpu_generate_rgg_shard", "mvgpu_generate_rgg_shard_ex",
           "mvgpu_download_shard", "mvgpu_louvain",
           "mvgpu_get_communities", "mvgpu_get_communities_device", "mvgpu_upload_shard32", "mvgpu_louvain32", "mvgpu_get_communities32", "mvgpu_set_option", "mvgpu_get_trace",
           "mvgpu_get_timings", "mvgpu_get_scan_times", "mvgpu_get_constant", "mvgpu_get_shard_info", "mvgpu_dist_louvain_method"]

_lib = None


def lib():
    """Load libmvgpu.so (fails loudly if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} missing: the CUDA extension is not built "
                               "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        L = ctypes.CDLL(_LIB_PATH)
        vp, i64, dbl, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int
        L.mvgpu_last_error.restype = ctypes.c_char_p
        L.mvgpu_create.argtypes = [ctypes.POINTER(vp), ci, ci, ci]
        L.mvgpu_destroy.argtypes = [vp]
        L.mvgpu_get_unique_id.argtypes = [vp]
        L.mvgpu_comm_init.argtypes = [vp, vp]
        L.mvgpu_upload_shard.argtypes = [vp, i64, vp, i64, i64, vp, vp]
        L.mvgpu_attach_shard_device.argtypes = [vp, i64, vp, i64, i64, vp, vp]
        L.mvgpu_generate_rgg_shard.argtypes = [vp, i64, ci, ctypes.POINTER(i64)]
        L.mvgpu_generate_rgg_shard_ex.argtypes = [vp, i64, ci, ci, ctypes.c_double, ctypes.POINTER(i64)]
        L.mvgpu_download_shard.argtypes = [vp, vp, vp]
        L.mvgpu_louvain.argtypes = [vp, dbl, dbl, ctypes.POINTER(ci), ctypes.POINTER(dbl)]
        L.mvgpu_get_communities.argtypes = [vp, vp]
        L.mvgpu_upload_shard32.argtypes = [vp, ctypes.c_int32, vp, ctypes.c_int32, ctypes.c_int32, vp, vp]
        L.mvgpu_louvain32.argtypes = [vp, ctypes.c_float, ctypes.c_float, ctypes.POINTER(ci), ctypes.POINTER(ctypes.c_float)]
        L.mvgpu_get_communities32.argtypes = [vp, vp]
        L.mvgpu_get_communities_device.argtypes = [vp, ctypes.POINTER(vp)]
        L.mvgpu_set_option.argtypes = [vp, ctypes.c_char_p, i64]
        L.mvgpu_get_trace.argtypes = [vp, ci, vp, ctypes.POINTER(ci)]
        L.mvgpu_get_timings.argtypes = [vp, ctypes.POINTER(Timings)]
        L.mvgpu_get_scan_times.argtypes = [vp, ci, vp, ctypes.POINTER(ci)]
        L.mvgpu_get_constant.argtypes = [vp, ctypes.POINTER(dbl)]
        L.mvgpu_get_shard_info.argtypes = [vp, vp]
        L.mvgpu_dist_louvain_method.argtypes = [ci, i64, i64, vp, vp, dbl, dbl, ctypes.POINTER(ci),
                                                ctypes.POINTER(dbl), vp]
        _lib = L
    return _lib


class MvgpuError(RuntimeError):
    pass


def _ck(rc):
    if rc != 0:
        raise MvgpuError(lib().mvgpu_last_error().decode(errors="replace"))


def device_count():
    n = lib().mvgpu_device_count()
    if n < 0:
        raise MvgpuError(lib().mvgpu_last_error().decode(errors="replace"))
    return n


def get_unique_id():
    buf = ctypes.create_string_buffer(UNIQUE_ID_BYTES)
    _ck(lib().mvgpu_get_unique_id(buf))
    return buf.raw


class LouvainGPU:
    """One rank == one GPU (the per-process state of the reference's distLouvainMethod)."""

    def __init__(self, device=0, rank=0, nranks=1):
        self._h = ctypes.c_void_p()
        self.rank, self.nranks, self.device = rank, nranks, device
        _ck(lib().mvgpu_create(ctypes.byref(self._h), device, rank, nranks))
        self.lnv = 0
        self._keep = None

    def comm_init(self, unique_id: bytes):
        assert len(unique_id) == UNIQUE_ID_BYTES
        _ck(lib().mvgpu_comm_init(self._h, ctypes.c_char_p(unique_id)))

    def upload(self, nv_global, parts, rowptr, edges):
        """Host arrays of a reference Graph shard (graph.hpp:289-293) -> HBM."""
        parts = np.ascontiguousarray(parts, dtype=np.int64)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        edges = np.ascontiguousarray(edges)
        assert edges.dtype.itemsize == 16, "edge records must be {int64 tail; double weight}"
        lnv, lne = len(rowptr) - 1, len(edges)
        assert int(rowptr[-1]) == lne
        _ck(lib().mvgpu_upload_shard(self._h, int(nv_global), parts.ctypes.data, lnv, lne, rowptr.ctypes.data,
                                     edges.ctypes.data if lne else None))
        self.lnv = lnv

    # ---- the reference's USE_32_BIT_GRAPH build: int32 ids, float weights, float results (include/mvgpu.h)
    def upload32(self, nv_global, parts, rowptr, edges):
        parts = np.ascontiguousarray(parts, dtype=np.int32)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        edges = np.ascontiguousarray(edges)
        assert edges.dtype.itemsize == 8, "edge records must be {int32 tail; float weight}"
        lnv, lne = len(rowptr) - 1, len(edges)
        assert int(rowptr[-1]) == lne
        _ck(lib().mvgpu_upload_shard32(self._h, int(nv_global), parts.ctypes.data, lnv, lne, rowptr.ctypes.data,
                                       edges.ctypes.data if lne else None))
        self.lnv = lnv

    def louvain32(self, lower=-1.0, thresh=1.0e-6):
        iters = ctypes.c_int(0)
        mod = ctypes.c_float(0.0)
        _ck(lib().mvgpu_louvain32(self._h, lower, thresh, ctypes.byref(iters), ctypes.byref(mod)))
        return mod.value, iters.value

    def communities32(self):
        out = np.empty(self.lnv, dtype=np.int32)
        _ck(lib().mvgpu_get_communities32(self._h, out.ctypes.data))
        return out

    def attach_device(self, nv_global, parts, lnv, lne, d_rowptr_ptr, d_edges_ptr, keepalive=None):
        """Arrays already resident in this GPU's memory (raw device pointers)."""
        parts = np.ascontiguousarray(parts, dtype=np.int64)
        _ck(lib().mvgpu_attach_shard_device(self._h, int(nv_global), parts.ctypes.data, int(lnv), int(lne),
                                            ctypes.c_void_p(d_rowptr_ptr), ctypes.c_void_p(d_edges_ptr)))
        self.lnv = int(lnv)
        self._keep = keepalive

    def generate_rgg(self, nv_global, unit_weight=True, lcg=False, random_edge_percent=0.0):
        """Build this rank's strip of `miniVite -n nv_global [-w] [-l] [-p pct]` on the device (reference GenerateRGG);
        returns lne.  With several ranks and random_edge_percent > 0 the call is collective (after comm_init)."""
        lne = ctypes.c_int64(0)
        _ck(lib().mvgpu_generate_rgg_shard_ex(self._h, int(nv_global), int(bool(unit_weight)), int(bool(lcg)),
                                              float(random_edge_percent), ctypes.byref(lne)))
        self.lnv = int(nv_global) // self.nranks
        self._lne = lne.value
        return lne.value

    def download_shard(self):
        """The shard's reference-format arrays (int64 rowptr, {tail, weight} records) as numpy arrays."""
        info = self.shard_info()
        rowptr = np.zeros(info["lnv"] + 1, dtype=np.int64)
        edges = np.zeros(info["lne"], dtype=np.dtype([("tail", "<i8"), ("weight", "<f8")]))
        _ck(lib().mvgpu_download_shard(self._h, rowptr.ctypes.data, edges.ctypes.data if info["lne"] else None))
        return rowptr, edges

    def set_option(self, name, value):
        _ck(lib().mvgpu_set_option(self._h, name.encode(), int(value)))

    def louvain(self, lower=-1.0, thresh=1.0e-6):
        iters = ctypes.c_int(0)
        mod = ctypes.c_double(0.0)
        _ck(lib().mvgpu_louvain(self._h, lower, thresh, ctypes.byref(iters), ctypes.byref(mod)))
        return mod.value, iters.value

    def communities(self, out=None):
        """currComm of this rank's vertices (global community ids).  `out`: optional preallocated int64 array
        (a pinned one is filled at full PCIe rate)."""
        if out is None:
            out = np.empty(self.lnv, dtype=np.int64)
        assert out.dtype == np.int64 and out.size >= self.lnv and out.flags["C_CONTIGUOUS"]
        _ck(lib().mvgpu_get_communities(self._h, out.ctypes.data))
        return out[:self.lnv]

    def trace(self):
        n = ctypes.c_int(0)
        _ck(lib().mvgpu_get_trace(self._h, 0, None, ctypes.byref(n)))
        out = np.zeros(n.value, dtype=TRACE_DTYPE)
        if n.value:
            _ck(lib().mvgpu_get_trace(self._h, n.value, out.ctypes.data, ctypes.byref(n)))
        return out

    def timings(self):
        t = Timings()
        _ck(lib().mvgpu_get_timings(self._h, ctypes.byref(t)))
        return t.as_dict()

    def scan_times(self):
        n = ctypes.c_int(0)
        _ck(lib().mvgpu_get_scan_times(self._h, 0, None, ctypes.byref(n)))
        out = np.zeros(n.value, dtype=np.float64)
        if n.value:
            _ck(lib().mvgpu_get_scan_times(self._h, n.value, out.ctypes.data, ctypes.byref(n)))
        return out

    def constant(self):
        v = ctypes.c_double(0)
        _ck(lib().mvgpu_get_constant(self._h, ctypes.byref(v)))
        return v.value

    def shard_info(self):
        info = (ctypes.c_int64 * 6)()
        _ck(lib().mvgpu_get_shard_info(self._h, info))
        return dict(zip(["lnv", "lne", "nghost", "nsend", "nheavy", "maxdeg"], [int(x) for x in info]))

    def close(self):
        if self._h:
            lib().mvgpu_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def dist_louvain_method(me, nprocs, shard, lower=-1.0, thresh=1.0e-6, device=None, unique_id=None, want_comm=False):
    """Mirror of the reference call `distLouvainMethod(me, nprocs, g, ..., lower, thresh, iters)`
    (dspl.hpp:1280-1283, called from main.cpp:168).  `shard` is a hostgraph.Shard (the arrays of the
    reference's Graph).  Returns (modularity, iters[, communities])."""
    g = LouvainGPU(device if device is not None else me, me, nprocs)
    try:
        if nprocs > 1:
            g.comm_init(unique_id)
        g.upload(shard.nv, shard.parts, shard.rowptr, shard.edges)
        mod, iters = g.louvain(lower, thresh)
        if want_comm:
            return mod, iters, g.communities()
        return mod, iters
    finally:
        g.close()
