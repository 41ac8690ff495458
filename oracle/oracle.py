"""ctypes shim over liboracle.so -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (pacmann_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "pacmann_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _SO


_lib = None
u8p, u32p, u64p, i32p, i64p, f32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int32, C.c_int64, C.c_float))


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        sig = {
            "orc_force_portable": (None, [C.c_int]),
            "orc_cpu_features": (C.c_int, []),
            "orc_mix64": (C.c_uint64, [C.c_uint64, C.c_uint64]),
            "orc_sbox": (None, [u8p]),
            "orc_expand_key": (None, [u8p, u32p]),
            "orc_encrypt_aes128": (None, [u32p, u8p, u8p]),
            "orc_aes128_mmo": (None, [u32p, u8p, u8p]),
            "orc_prf": (C.c_uint64, [u32p, C.c_uint64, C.c_uint64]),
            "orc_prf_batch": (None, [u32p, u64p, u64p, C.c_uint64, u64p]),
            "orc_prf_eval4": (C.c_uint64, [u8p, C.c_uint64]),
            "orc_xor_slices": (None, [u64p, u64p, C.c_int64]),
            "orc_l2_distance_simd": (C.c_float, [f32p, f32p, C.c_int64]),
            "orc_l2dist": (C.c_float, [f32p, f32p, C.c_int64]),
            "orc_l2dist_batch": (None, [f32p, C.c_int64, C.c_int64, f32p, i64p, C.c_int64, C.c_int64, f32p]),
            "orc_robust_prune": (C.c_int64, [f32p, C.c_int64, C.c_int64, i64p, C.c_int64, C.c_int64, C.c_float, i64p]),
            "orc_inner_product": (C.c_uint32, [u32p, u32p, C.c_int64]),
            "orc_ip_scan": (None, [u32p, C.c_int64, C.c_int64, u32p, C.c_int64, u32p, C.c_int]),
            "orc_gen_params": (None, [C.c_uint64, u64p, u64p]),
            "orc_client_params": (None, [C.c_uint64] * 4 + [u64p] * 3),
            "orc_pir_new": (C.c_void_p, [C.c_uint64, C.c_uint64, u64p, C.c_uint64]),
            "orc_pir_free": (None, [C.c_void_p]),
            "orc_pir_initialization": (None, [C.c_void_p, u8p]),
            "orc_pir_preprocessing": (None, [C.c_void_p, u8p, C.c_uint64, C.c_int]),
            "orc_pir_preprocessing_range": (None, [C.c_void_p, u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]),
            "orc_pir_dummy_preprocessing": (None, [C.c_void_p, u8p]),
            "orc_pir_private_query": (None, [C.c_void_p, u32p, u64p]),
            "orc_pir_nonprivate_query": (C.c_int, [C.c_void_p, C.c_uint64, u64p]),
            "orc_pir_client_query": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, u64p, u32p]),
            "orc_pir_query": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, u64p, u8p, C.c_uint64]),
            "orc_pir_local_storage": (C.c_double, [C.c_void_p]),
            "orc_pir_comm_cost": (C.c_double, [C.c_void_p]),
            "orc_pir_get": (C.c_uint64, [C.c_void_p, C.c_int]),
            "orc_pir_table": (u64p, [C.c_void_p, C.c_int]),
            "orc_pir_long_key": (u32p, [C.c_void_p]),
            "orc_pir_set_dummy_seed": (None, [C.c_void_p, C.c_uint64]),
            "orc_batch_new": (C.c_void_p, [C.c_uint64, C.c_uint64, C.c_uint64, u64p, C.c_uint64]),
            "orc_batch_free": (None, [C.c_void_p]),
            "orc_batch_sub": (C.c_void_p, [C.c_void_p, C.c_uint64]),
            "orc_batch_get": (C.c_uint64, [C.c_void_p, C.c_int]),
            "orc_derive_key": (None, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, u8p]),
            "orc_batch_preprocessing": (None, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int]),
            "orc_batch_dummy_preprocessing": (None, [C.c_void_p, C.c_uint64]),
            "orc_batch_query": (C.c_int, [C.c_void_p, u64p, C.c_uint64, u64p, C.POINTER(C.c_int)]),
            "orc_batch_local_storage": (C.c_double, [C.c_void_p]),
            "orc_batch_comm_online": (C.c_uint64, [C.c_void_p]),
            "orc_pack_db": (None, [f32p, i32p, C.c_uint64, C.c_uint64, C.c_uint64, u64p]),
            "orc_unpack_entry": (None, [u64p, C.c_uint64, C.c_uint64, f32p, i64p]),
            "orc_search_knn_basic": (C.c_int, [f32p, i32p] + [C.c_int64] * 3 + [i64p, C.c_int64, f32p] + [C.c_int64] * 4 + [i64p, i64p]),
            "orc_search_knn_private": (C.c_int, [C.c_void_p, f32p, i32p] + [C.c_int64] * 3 + [i64p, C.c_int64, f32p] + [C.c_int64] * 4 + [C.c_int, C.c_uint64, i64p, i64p, i64p]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def mix64(seed, ctr):
    return lib().orc_mix64(seed, ctr)


def key_bytes(key):
    k = np.frombuffer(bytes(key), dtype=np.uint8).copy()
    assert k.size == 16
    return k


def expand_key(key):
    """GetLongKey (util.go:167-171): 16-byte key -> 44 LE uint32 round-key words."""
    rk = np.zeros(44, np.uint32)
    lib().orc_expand_key(_p(key_bytes(key), u8p), _p(rk, u32p))
    return rk


def encrypt_aes128(rk, block):
    src, dst = key_bytes(block), np.zeros(16, np.uint8)
    lib().orc_encrypt_aes128(_p(rk, u32p), _p(dst, u8p), _p(src, u8p))
    return bytes(dst)


def aes128_mmo(rk, block):
    src, dst = key_bytes(block), np.zeros(16, np.uint8)
    lib().orc_aes128_mmo(_p(rk, u32p), _p(dst, u8p), _p(src, u8p))
    return bytes(dst)


def prf(rk, tag, x):
    return lib().orc_prf(_p(rk, u32p), tag, x)


def prf_batch(rk, tags, xs):
    tags = np.ascontiguousarray(tags, np.uint64)
    xs = np.ascontiguousarray(xs, np.uint64)
    out = np.zeros(tags.size, np.uint64)
    lib().orc_prf_batch(_p(rk, u32p), _p(tags, u64p), _p(xs, u64p), tags.size, _p(out, u64p))
    return out


def xor_slices(dst, src):
    """xorSlices(dst, src, n): in place; count comes from len(src) (aes_amd64.s:136)."""
    lib().orc_xor_slices(_p(dst, u64p), _p(src, u64p), src.size)
    return dst


def l2dist(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return np.float32(lib().orc_l2dist(_p(a, f32p), _p(b, f32p), a.size))


def l2dist_batch(vecs, dim, queries, ids):
    """vecs: [N][stride] f32 (first dim floats of each row are the vector); ids [Q][K]."""
    vecs = np.ascontiguousarray(vecs, np.float32)
    queries = np.ascontiguousarray(queries, np.float32)
    ids = np.ascontiguousarray(ids, np.int64)
    out = np.zeros(ids.shape, np.float32)
    lib().orc_l2dist_batch(_p(vecs, f32p), vecs.shape[1], dim, _p(queries, f32p), _p(ids, i64p), ids.shape[0], ids.shape[1], _p(out, f32p))
    return out


def robust_prune(vectors, u, candidates, m, alpha):
    """robustPrune (build_graph.go:169-236); ties in the distance to u keep candidate order."""
    vectors = np.ascontiguousarray(vectors, np.float32)
    cand = np.ascontiguousarray(candidates, np.int64)
    out = np.zeros(max(cand.size, 1), np.int64)
    n = lib().orc_robust_prune(_p(vectors, f32p), vectors.shape[1], int(u), _p(cand, i64p), cand.size, int(m), float(alpha), _p(out, i64p))
    return out[:n].copy()


def inner_product(a, b):
    a = np.ascontiguousarray(a, np.uint32)
    b = np.ascontiguousarray(b, np.uint32)
    return lib().orc_inner_product(_p(a, u32p), _p(b, u32p), a.size)


def ip_scan(rows, queries, threads=1):
    rows = np.ascontiguousarray(rows, np.uint32)
    queries = np.ascontiguousarray(queries, np.uint32).reshape(-1, rows.shape[1])
    out = np.zeros(queries.shape[0], np.uint32)
    lib().orc_ip_scan(_p(rows, u32p), rows.shape[0], rows.shape[1], _p(queries, u32p), queries.shape[0], _p(out, u32p), threads)
    return out


def gen_params(db_size):
    c, s = C.c_uint64(), C.c_uint64()
    lib().orc_gen_params(db_size, C.byref(c), C.byref(s))
    return c.value, s.value


def client_params(db_size, fail_log2):
    c, s = gen_params(db_size)
    a, b, d = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib().orc_client_params(db_size, c, s, fail_log2, C.byref(a), C.byref(b), C.byref(d))
    return dict(chunk_size=c, set_size=s, max_query_num=a.value, primary_hint_num=b.value, max_query_per_chunk=d.value)


_GET = dict(entry_u64=0, db_size=1, chunk_size=2, set_size=3, max_query_num=4, primary_hint_num=5,
            max_query_per_chunk=6, finished_query_num=7, n_private_queries=8)


class PianoPIR:
    """Oracle PianoPIR (pir.go:473-548): client + server over an aliased flat rawDB."""

    def __init__(self, db_size, entry_bytes, raw_db, fail_log2, _handle=None, _keep=None):
        self._owned = _handle is None
        if _handle is None:
            self.raw_db = np.ascontiguousarray(raw_db, np.uint64)
            assert self.raw_db.size == db_size * (entry_bytes // 8)
            _handle = lib().orc_pir_new(db_size, entry_bytes, _p(self.raw_db, u64p), fail_log2)
        else:
            self.raw_db = _keep
        self.h = C.c_void_p(_handle)
        for k, v in _GET.items():
            if k not in ("finished_query_num", "n_private_queries"):
                setattr(self, k, lib().orc_pir_get(self.h, v))

    def __del__(self):
        if getattr(self, "_owned", False) and self.h:
            lib().orc_pir_free(self.h)
            self.h = None

    def get(self, name):
        return lib().orc_pir_get(self.h, _GET[name])

    def preprocessing(self, key, repl_seed=0, threads=1):
        lib().orc_pir_preprocessing(self.h, _p(key_bytes(key), u8p), repl_seed, threads)

    def preprocessing_range(self, key, h0, h1, repl_seed=0, do_repl=False):
        lib().orc_pir_preprocessing_range(self.h, _p(key_bytes(key), u8p), repl_seed, h0, h1, int(do_repl))

    def dummy_preprocessing(self, key):
        lib().orc_pir_dummy_preprocessing(self.h, _p(key_bytes(key), u8p))

    def long_key(self):
        return np.ctypeslib.as_array(lib().orc_pir_long_key(self.h), shape=(44,)).copy()

    def table(self, name):
        E, P, S, M = self.entry_u64, self.primary_hint_num, self.set_size, self.max_query_per_chunk
        which = dict(primary_short_tag=(0, (P,)), primary_parity=(1, (P, E)), primary_program_point=(2, (P,)),
                     replacement_idx=(3, (S, M)), replacement_val=(4, (S, M, E)), backup_short_tag=(5, (S, M)),
                     backup_parity=(6, (S, M, E)), query_histogram=(7, (S,)))[name]
        ptr = lib().orc_pir_table(self.h, which[0])
        return np.ctypeslib.as_array(ptr, shape=which[1])

    def private_query(self, offsets):
        offsets = np.ascontiguousarray(offsets, np.uint32)
        assert offsets.size == self.set_size
        ret = np.zeros(self.entry_u64, np.uint64)
        lib().orc_pir_private_query(self.h, _p(offsets, u32p), _p(ret, u64p))
        return ret

    def client_query(self, idx, real=True, want_offsets=False):
        ret = np.zeros(self.entry_u64, np.uint64)
        offs = np.zeros(self.set_size, np.uint32)
        rc = lib().orc_pir_client_query(self.h, idx, int(real), _p(ret, u64p), _p(offs, u32p))
        return (ret, rc, offs) if want_offsets else (ret, rc)

    def query(self, idx, real=True, rekey=bytes(16), repl_seed=0):
        ret = np.zeros(self.entry_u64, np.uint64)
        rc = lib().orc_pir_query(self.h, idx, int(real), _p(ret, u64p), _p(key_bytes(rekey), u8p), repl_seed)
        return ret, rc

    def local_storage_size(self):
        return lib().orc_pir_local_storage(self.h)

    def comm_cost_per_query(self):
        return lib().orc_pir_comm_cost(self.h)


class SimpleBatchPianoPIR:
    """Oracle SimpleBatchPianoPIR (batch-pir.go)."""

    def __init__(self, db_size, entry_bytes, batch_size, raw_db, fail_log2):
        self.raw_db = np.ascontiguousarray(raw_db, np.uint64)
        assert self.raw_db.size == db_size * (entry_bytes // 8)
        self.entry_u64 = entry_bytes // 8
        self.h = C.c_void_p(lib().orc_batch_new(db_size, entry_bytes, batch_size, _p(self.raw_db, u64p), fail_log2))
        self.partition_num = lib().orc_batch_get(self.h, 0)
        self.partition_size = lib().orc_batch_get(self.h, 1)
        self.db_size, self.entry_bytes, self.batch_size = db_size, entry_bytes, batch_size

    def __del__(self):
        if self.h:
            lib().orc_batch_free(self.h)
            self.h = None

    def sub(self, i):
        return PianoPIR(0, 0, None, 0, _handle=lib().orc_batch_sub(self.h, i), _keep=self.raw_db)

    def preprocessing(self, key_seed, repl_seed=0, threads=1):
        lib().orc_batch_preprocessing(self.h, key_seed, repl_seed, threads)

    def dummy_preprocessing(self, key_seed):
        lib().orc_batch_dummy_preprocessing(self.h, key_seed)

    def query(self, idx, want_status=False):
        idx = np.ascontiguousarray(idx, np.uint64)
        out = np.zeros((idx.size, self.entry_u64), np.uint64)
        st = np.zeros(idx.size, np.int32)
        rc = lib().orc_batch_query(self.h, _p(idx, u64p), idx.size, _p(out, u64p), st.ctypes.data_as(C.POINTER(C.c_int)))
        if rc < 0:
            raise IndexError("index out of range")
        return (out, st) if want_status else out

    @property
    def finished_batch_num(self):
        return lib().orc_batch_get(self.h, 2)

    @property
    def queries_made_in_partition(self):
        return lib().orc_batch_get(self.h, 3)

    @property
    def support_batch_num(self):
        return lib().orc_batch_get(self.h, 4)

    def local_storage_size(self):
        return lib().orc_batch_local_storage(self.h)

    def comm_cost_per_batch_online(self):
        return lib().orc_batch_comm_online(self.h)


def derive_key(key_seed, epoch, parts, i):
    k = np.zeros(16, np.uint8)
    lib().orc_derive_key(key_seed, epoch, parts, i, _p(k, u8p))
    return bytes(k)


def pack_db(vectors, graph):
    vectors = np.ascontiguousarray(vectors, np.float32)
    graph = np.ascontiguousarray(graph, np.int32)
    n, dim = vectors.shape
    m = graph.shape[1]
    raw = np.zeros(n * (dim + m) // 2, np.uint64)
    lib().orc_pack_db(_p(vectors, f32p), _p(graph, i32p), n, dim, m, _p(raw, u64p))
    return raw


def search_knn_basic(vectors, graph, start_ids, queries, k, max_step, parallel):
    vectors = np.ascontiguousarray(vectors, np.float32)
    graph = np.ascontiguousarray(graph, np.int32)
    start_ids = np.ascontiguousarray(start_ids, np.int64)
    queries = np.ascontiguousarray(queries, np.float32).reshape(-1, vectors.shape[1])
    nq = queries.shape[0]
    ret, step = np.zeros((nq, k), np.int64), np.zeros((nq, k), np.int64)
    rc = lib().orc_search_knn_basic(_p(vectors, f32p), _p(graph, i32p), vectors.shape[0], vectors.shape[1], graph.shape[1],
                                    _p(start_ids, i64p), start_ids.size, _p(queries, f32p), nq, k, max_step, parallel,
                                    _p(ret, i64p), _p(step, i64p))
    assert rc == 0
    return ret, step


def search_knn_private(pir, vectors, graph, start_ids, queries, k, max_step, parallel, benchmarking=False, rand_seed=0):
    vectors = np.ascontiguousarray(vectors, np.float32)
    graph = np.ascontiguousarray(graph, np.int32)
    start_ids = np.ascontiguousarray(start_ids, np.int64)
    queries = np.ascontiguousarray(queries, np.float32).reshape(-1, vectors.shape[1])
    nq = queries.shape[0]
    ret, step = np.zeros((nq, k), np.int64), np.zeros((nq, k), np.int64)
    stats = np.zeros(2, np.int64)
    rc = lib().orc_search_knn_private(pir.h, _p(vectors, f32p), _p(graph, i32p), vectors.shape[0], vectors.shape[1],
                                      graph.shape[1], _p(start_ids, i64p), start_ids.size, _p(queries, f32p), nq, k,
                                      max_step, parallel, int(benchmarking), rand_seed, _p(ret, i64p), _p(step, i64p),
                                      _p(stats, i64p))
    assert rc == 0
    return ret, step, stats
