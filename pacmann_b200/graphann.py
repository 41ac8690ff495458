"""Python face of the host-side mirror of the reference's `graphann` package and of private-search.go's
PIRGraphInfo (csrc/host/graphann.cpp).  Names follow graphann/search.go."""
import ctypes as C

import numpy as np

from . import _host
from .pianopir import SimpleBatchPianoPIR


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def L2Dist(v1, v2, device=0):
    """graphann.L2Dist (build_graph.go:119-127): one distance, evaluated on the GPU in the reference's fp32 order."""
    v1 = np.ascontiguousarray(v1, np.float32)
    v2 = np.ascontiguousarray(v2, np.float32)
    assert v1.size == v2.size
    return np.float32(_host.lib().pmh_l2dist(_p(v1), _p(v2), v1.size, device))


class GraphANNFrontend:
    """graphann.GraphANNFrontend over BasicGraphInfo (non-private) or PIRGraphInfo (private, private-search.go)."""

    def __init__(self, vectors, graph, private=False, skipPrep=False, nonPrivateMode=False, seed=1, device=0, resident=True,
                 share_db_with=None):
        """share_db_with: another private frontend (already preprocessed) whose GPU-resident rawDB this client reuses --
        one DB replica per GPU, one client (keys, hint tables, search state) per user."""
        self.vectors = np.ascontiguousarray(vectors, np.float32)
        self.graph = np.ascontiguousarray(graph, np.int32)
        self.n, self.dim = self.vectors.shape
        self.m = self.graph.shape[1]
        L = _host.lib()
        self._shared = share_db_with      # keep the owner of the DB alive
        if share_db_with is not None:
            h = L.pmh_frontend_pir_shared(share_db_with.h, seed, int(resident))
        elif private:
            h = L.pmh_frontend_pir(self.n, self.dim, self.m, _p(self.graph), _p(self.vectors), int(skipPrep), int(nonPrivateMode), seed, device, int(resident))
        else:
            h = L.pmh_frontend_basic(self.n, self.dim, self.m, _p(self.graph), _p(self.vectors))
        if not h:
            raise _host.HostError(L.pmh_last_error().decode())
        self.h = C.c_void_p(h)
        self.private = private or share_db_with is not None

    def __del__(self):
        if getattr(self, "h", None):
            _host.lib().pmh_frontend_free(self.h)
            self.h = None

    def Preprocess(self):
        _host.check(_host.lib().pmh_frontend_preprocess(self.h))

    def GetMetadata(self):
        return self.n, self.dim, self.m

    def StartVertexIds(self):
        out = np.zeros(self.n, np.int64)
        k = _host.lib().pmh_frontend_start_ids(self.h, _p(out), out.size)
        return out[:k].copy()

    def SetStartVertexIds(self, ids):
        ids = np.ascontiguousarray(ids, np.int64)
        _host.lib().pmh_frontend_set_start_ids(self.h, _p(ids), ids.size)

    def SetRandSeed(self, seed):
        _host.lib().pmh_frontend_set_rand_seed(self.h, seed)

    def SearchKNNBatch(self, queryVectors, k, maxStep, parallel, benchmarking=False):
        q = np.ascontiguousarray(queryVectors, np.float32).reshape(-1, self.dim)
        ret = np.zeros((q.shape[0], k), np.int64)
        step = np.zeros((q.shape[0], k), np.int64)
        rc = _host.check(_host.lib().pmh_frontend_search_knn(self.h, _p(q), q.shape[0], k, maxStep, parallel, int(benchmarking), _p(ret), _p(step)))
        if rc != 0:
            raise RuntimeError("GetVertexInfo failed")
        return ret, step

    def SearchKNN(self, queryVector, k, maxStep, parallel, benchmarking=False):
        r, s = self.SearchKNNBatch(np.asarray(queryVector, np.float32).reshape(1, -1), k, maxStep, parallel, benchmarking)
        return r[0], s[0]

    @property
    def PIR(self):
        h = _host.lib().pmh_frontend_pir_handle(self.h)
        return SimpleBatchPianoPIR(self.n, (self.dim + self.m) * 4, self.m, None, 8, _borrow=h) if h else None

    @property
    def totalQueryNum(self):
        return _host.lib().pmh_frontend_stat(self.h, 0)

    @property
    def succQueryNum(self):
        return _host.lib().pmh_frontend_stat(self.h, 1)
