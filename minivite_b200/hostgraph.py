"""ctypes view of libmvhost.so: RGG generator + binary graph file I/O (host side, no CUDA).

The shards come back as numpy views over the C++ `Graph` objects (reference graph.hpp:85-296
layout: int64 rowptr[lnv+1], {int64 tail; double weight}[lne]).
"""
import ctypes
import os

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmvhost.so")
EDGE_DTYPE = np.dtype([("tail", "<i8"), ("weight", "<f8")])
assert EDGE_DTYPE.itemsize == 16

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = ctypes.CDLL(_LIB_PATH)
        L.mvh_last_error.restype = ctypes.c_char_p
        L.mvh_rgg_generate.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_double, ctypes.POINTER(ctypes.c_void_p)]
        L.mvh_rgg_radius.restype = ctypes.c_double
        L.mvh_rgg_radius.argtypes = [ctypes.c_int64, ctypes.c_int]
        L.mvh_graph_read.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_void_p)]
        L.mvh_graph_count.argtypes = [ctypes.c_void_p]
        L.mvh_graph_shard.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int)]
        L.mvh_graph_write.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.mvh_graph_free.argtypes = [ctypes.c_void_p]
        L.mvh_set_num_threads.argtypes = [ctypes.c_int]
        L.mvh_reseeder.restype = ctypes.c_int64
        L.mvh_reseeder.argtypes = [ctypes.c_uint]
        L.mvh_rgg_points.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                     ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


def set_num_threads(n):
    """OpenMP width of the host-side generator (torchrun exports OMP_NUM_THREADS=1)."""
    lib().mvh_set_num_threads(int(n))


def _check(rc):
    if rc != 0:
        raise RuntimeError("libmvhost: " + lib().mvh_last_error().decode())


class Shard:
    """One vertex-range shard: the arrays a reference rank would hold."""

    def __init__(self, owner, base, bound, lnv, lne, nv, ne, rowptr, edges, parts):
        self._owner = owner          # keeps the C++ objects alive
        self.base, self.bound, self.lnv, self.lne, self.nv, self.ne = base, bound, lnv, lne, nv, ne
        self.rowptr, self.edges, self.parts = rowptr, edges, parts

    @property
    def tails(self):
        return self.edges["tail"]

    @property
    def weights(self):
        return self.edges["weight"]


class ShardSet:
    def __init__(self, handle):
        self._h = ctypes.c_void_p(handle)
        self.shards = []
        L = lib()
        for i in range(L.mvh_graph_count(self._h)):
            info = (ctypes.c_int64 * 6)()
            rp, ed, pp = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
            npart = ctypes.c_int()
            _check(L.mvh_graph_shard(self._h, i, info, ctypes.byref(rp), ctypes.byref(ed), ctypes.byref(pp),
                                     ctypes.byref(npart)))
            base, bound, lnv, lne, nv, ne = [int(x) for x in info]
            rowptr = np.ctypeslib.as_array(ctypes.cast(rp, ctypes.POINTER(ctypes.c_int64)), shape=(lnv + 1,))
            if lne:
                raw = np.ctypeslib.as_array(ctypes.cast(ed, ctypes.POINTER(ctypes.c_uint8)), shape=(lne * 16,))
                edges = raw.view(EDGE_DTYPE)
            else:
                edges = np.zeros(0, EDGE_DTYPE)
            parts = np.ctypeslib.as_array(ctypes.cast(pp, ctypes.POINTER(ctypes.c_int64)),
                                          shape=(npart.value,)).copy()
            self.shards.append(Shard(self, base, bound, lnv, lne, nv, ne, rowptr, edges, parts))

    def write(self, path):
        _check(lib().mvh_graph_write(self._h, os.fsencode(path)))

    def close(self):
        if self._h:
            self.shards = []
            lib().mvh_graph_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def generate_rgg(nv, nprocs=1, r_begin=0, r_end=-1, lcg=False, unit_weight=True, random_edge_percent=0.0):
    """Shards [r_begin, r_end) of the graph `miniVite -n nv [-l] [-w] [-p pct]` builds on nprocs ranks."""
    h = ctypes.c_void_p()
    _check(lib().mvh_rgg_generate(nv, nprocs, r_begin, r_end, int(lcg), int(unit_weight),
                                  float(random_edge_percent), ctypes.byref(h)))
    return ShardSet(h.value)


def reseeder(initseed=1):
    """The reference's seed helper (utils.hpp:91-98)."""
    return int(lib().mvh_reseeder(initseed))


def rgg_points(nv, nprocs=1, rank=0, count=None, lcg=False):
    """First `count` coordinates (X, Y) rank `rank` of `miniVite -n nv` on nprocs ranks draws (graph.hpp:680-700)."""
    n = nv // nprocs if count is None else min(count, nv // nprocs)
    x, y = np.zeros(n, np.float64), np.zeros(n, np.float64)
    _check(lib().mvh_rgg_points(nv, nprocs, rank, int(lcg), n, x.ctypes.data, y.ctypes.data))
    return x, y


def rgg_radius(nv, nprocs=1):
    return lib().mvh_rgg_radius(nv, nprocs)


def read_graph(path, me=0, nprocs=1, balanced=False):
    """Shard `me` of `nprocs` of a binary graph file (reference graph.hpp:315 / 466)."""
    h = ctypes.c_void_p()
    _check(lib().mvh_graph_read(os.fsencode(path), me, nprocs, int(balanced), ctypes.byref(h)))
    return ShardSet(h.value)


def write_graph_arrays(path, nv, rowptr, tails, weights=None):
    """Write global CSR arrays in the reference's binary format (graph.hpp:342-403)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    ne = int(rowptr[-1])
    e = np.zeros(ne, EDGE_DTYPE)
    e["tail"] = tails
    e["weight"] = 1.0 if weights is None else weights
    with open(path, "wb") as f:
        f.write(np.array([nv, ne], dtype=np.int64).tobytes())
        f.write(rowptr.tobytes())
        f.write(e.tobytes())


EDGE32_DTYPE = np.dtype([("tail", "<i4"), ("weight", "<f4")])


def write_graph_arrays32(path, nv, rowptr, tails, weights=None):
    """Same file for a miniVite compiled with -DUSE_32_BIT_GRAPH (utils.hpp:72-82): every GraphElem is int32, every
    GraphWeight float -- header {int32 nv; int32 ne}, int32 offsets, {int32 tail; float weight} records."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    ne = int(rowptr[-1])
    e = np.zeros(ne, EDGE32_DTYPE)
    e["tail"] = tails
    e["weight"] = 1.0 if weights is None else weights
    with open(path, "wb") as f:
        f.write(np.array([nv, ne], dtype=np.int32).tobytes())
        f.write(rowptr.tobytes())
        f.write(e.tobytes())
