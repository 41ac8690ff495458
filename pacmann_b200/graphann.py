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

    def __init__(self, vectors, graph, private=False, skipPrep=False, nonPrivateMode=False, seed=None, device=0, resident=True,
                 share_db_with=None, group_lanes=1, lane_of=None, lane=0):
        """seed: None = the client draws its keys, replacement and dummy offsets and start vertices from the OS CSPRNG (what
        a deployment wants); an integer > 0 makes everything deterministic -- the hook the parity tests use.
        share_db_with: another private frontend (already preprocessed) whose GPU-resident rawDB this client reuses --
        one DB replica per GPU, one client (keys, hint tables, search state) per user.
        group_lanes / lane_of / lane: client groups for SearchKNNLockstep -- the first client is created with
        group_lanes=L and preprocessed, clients 1..L-1 with lane_of=first, lane=i (see make_client_group)."""
        self.vectors = np.ascontiguousarray(vectors, np.float32)
        self.graph = np.ascontiguousarray(graph, np.int32)
        self.n, self.dim = self.vectors.shape
        self.m = self.graph.shape[1]
        seed = 0 if seed is None else int(seed)
        L = _host.lib()
        self._shared = share_db_with if share_db_with is not None else lane_of      # keep the owner of the DB / client group alive
        if lane_of is not None:
            h = L.pmh_frontend_pir_lane(lane_of.h, seed, lane)
        elif share_db_with is not None:
            h = L.pmh_frontend_pir_shared(share_db_with.h, seed, int(resident))
        elif private:
            h = L.pmh_frontend_pir(self.n, self.dim, self.m, _p(self.graph), _p(self.vectors), int(skipPrep), int(nonPrivateMode), seed, device, int(resident))
        else:
            h = L.pmh_frontend_basic(self.n, self.dim, self.m, _p(self.graph), _p(self.vectors))
        if not h:
            raise _host.HostError(L.pmh_last_error().decode())
        self.h = C.c_void_p(h)
        self.private = private or share_db_with is not None or lane_of is not None
        if group_lanes > 1:
            _host.check(L.pmh_frontend_set_group_lanes(self.h, group_lanes))

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


def make_client_group(vectors, graph, lanes, seeds=None, skipPrep=False, device=0):
    """L independent private clients (own keys, hint tables, caches, search state) over ONE GPU-resident rawDB whose hint
    tables live in ONE pm_client, preprocessed and ready for SearchKNNLockstep.  seeds[i] = client i's seed (tests);
    None = every client draws its own secrets from the OS CSPRNG."""
    seeds = list(seeds) if seeds is not None else [None] * lanes
    first = GraphANNFrontend(vectors, graph, private=True, skipPrep=skipPrep, seed=seeds[0], device=device, group_lanes=lanes)
    first.Preprocess()
    group = [first]
    for i in range(1, lanes):
        f = GraphANNFrontend(first.vectors, first.graph, seed=seeds[i], lane_of=first, lane=i)
        f.Preprocess()
        group.append(f)
    return group


def SearchKNNLockstep(lanes, queryVectors, k, maxStep, parallel, benchmarking=False):
    """SearchKNNBatch over the clients of a group in lock step: query i is searched by lanes[i % L]; the results equal
    each client's own SearchKNNBatch over its queries, every step's fetches of all lanes share one device call."""
    dim = lanes[0].dim
    q = np.ascontiguousarray(queryVectors, np.float32).reshape(-1, dim)
    ret = np.zeros((q.shape[0], k), np.int64)
    step = np.zeros((q.shape[0], k), np.int64)
    hs = (C.c_void_p * len(lanes))(*[f.h for f in lanes])
    rc = _host.check(_host.lib().pmh_search_knn_lockstep(hs, len(lanes), _p(q), q.shape[0], k, maxStep, parallel, int(benchmarking), _p(ret), _p(step)))
    if rc != 0:
        raise RuntimeError("GetVertexInfo failed")
    return ret, step


def DeviceSearchStats():
    """(rounds, queries searched with the frontier on the GPU, queries of lanes in host mode) since the process started"""
    out = np.zeros(3, np.uint64)
    _host.lib().pmh_device_search_stats(_p(out))
    return tuple(int(x) for x in out)


def RobustPruneBatch(vectors, us, candidates, m, alpha, device=0):
    """robustPrune (build_graph.go:169-236) for many vertices at once: candidates[i] are the candidates of vertex us[i]
    (an [n][k] array; k <= m returns them unchanged as the reference does).  Every distance the reference evaluates --
    u to candidate, candidate to candidate -- comes from one GPU launch; returns a list of neighbour-id arrays."""
    v = np.ascontiguousarray(vectors, np.float32)
    us = np.ascontiguousarray(us, np.int64).reshape(-1)
    cand = np.ascontiguousarray(candidates, np.int64).reshape(us.size, -1)
    k = cand.shape[1]
    if k <= m:                     # the reference returns the candidates unchanged (build_graph.go:170-172)
        return [cand[i].copy() for i in range(us.size)]
    out = np.zeros((us.size, max(m, 1)), np.int64)
    lens = np.zeros(us.size, np.int64)
    _host.check(_host.lib().pmh_robust_prune_batch(_p(v), v.shape[0], v.shape[1], _p(us), us.size, _p(cand), k, m, float(alpha), device, _p(out), _p(lens)))
    return [out[i, :lens[i]].copy() for i in range(us.size)]


def EvaluateGraphQuality(vectors, graph, numQueries=100, seed=0):
    """EvaluateGraphQuality (build_graph.go:776-817): search numQueries random dataset vertices (k 20, step 20, parallel 2,
    non-private) and report (hit rate, average step at which a hit target was reached)."""
    f = GraphANNFrontend(vectors, graph)
    f.Preprocess()
    n = f.n
    targets = np.random.default_rng(seed).integers(0, n, numQueries)
    ret, steps = f.SearchKNNBatch(f.vectors[targets], 20, 20, 2)
    hit = ret[:, 0] == targets
    avg = float(steps[hit, 0].mean()) if hit.any() else float("nan")
    return float(hit.mean()), avg
