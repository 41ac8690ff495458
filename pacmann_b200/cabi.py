"""ctypes binding of libpacmann_cuda.so, one Python function per C-ABI entry point.

This is the same boundary a cgo bridge binds (include/pacmann_cuda.h).  There is no
fallback of any kind: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpacmann_cuda.so")

PM_OK, PM_ERR_ARG, PM_ERR_CUDA, PM_ERR_UNSUPPORTED, PM_ERR_NOMEM = 0, -1, -2, -3, -4
PM_NO_SKIP = -1

u8p, u32p, u64p, i32p, i64p, f32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int32, C.c_int64, C.c_float))


class PacmannError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpacmann_cuda error {code}: {msg}")
        self.code = code


class HintJob(C.Structure):
    """struct pm_hint_job"""
    _fields_ = [
        ("row0", C.c_uint64), ("n_rows", C.c_uint64),
        ("chunk_size", C.c_uint64), ("set_size", C.c_uint64),
        ("rk", C.c_uint32 * 44),
        ("hint_begin", C.c_uint64), ("n_hints", C.c_uint64),
        ("n_primary", C.c_uint64), ("backup_group", C.c_uint64),
        ("tags", C.c_void_p), ("skip_chunk", C.c_void_p), ("parity_out", C.c_void_p), ("offsets_out", C.c_void_p),
    ]


# every symbol include/pacmann_cuda.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "pm_version": (C.c_char_p, []),
    "pm_last_error": (C.c_char_p, []),
    "pm_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pm_launch_count": (C.c_uint64, []),
    "pm_tuning_set": (C.c_int, [C.c_char_p, C.c_int]),
    "pm_tuning_get": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "pm_db_create": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]),
    "pm_db_create_empty": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]),
    "pm_db_wrap": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]),
    "pm_db_upload": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "pm_db_info": (C.c_int, [C.c_void_p, u64p, u64p, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "pm_db_destroy": (C.c_int, [C.c_void_p]),
    "pm_db_sync": (C.c_int, [C.c_void_p]),
    "pm_buf_alloc": (C.c_int, [C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]),
    "pm_buf_free": (C.c_int, [C.c_void_p, C.c_int]),
    "pm_buf_ipc_export": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "pm_buf_ipc_open": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "pm_buf_ipc_close": (C.c_int, [C.c_void_p, C.c_int]),
    "pm_buf_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]),
    "pm_buf_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]),
    "pm_buf_zero": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int]),
    "pm_buf_copy_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "pm_flag_signal_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "pm_flag_wait_dev": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]),
    "pm_host_register": (C.c_int, [C.c_void_p, C.c_uint64]),
    "pm_host_unregister": (C.c_int, [C.c_void_p]),
    "pm_expand_key": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pm_expand_key_batch": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "pm_prf_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "pm_xor_slices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "pm_hintgen": (C.c_int, [C.c_void_p, C.POINTER(HintJob), C.c_uint64]),
    "pm_hintgen_dev": (C.c_int, [C.c_void_p, C.POINTER(HintJob), C.c_uint64, C.c_void_p]),
    "pm_gather_rows": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]),
    "pm_answer_batch": (C.c_int, [C.c_void_p] * 6 + [C.c_uint64, C.c_uint64, C.c_void_p]),
    "pm_answer_batch_dev": (C.c_int, [C.c_void_p] * 6 + [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "pm_client_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "pm_client_destroy": (C.c_int, [C.c_void_p]),
    "pm_client_preprocess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]),
    "pm_client_query_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "pm_client_query_batch_l2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "pm_client_query_batch_l2m": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                            C.c_uint64, C.c_void_p]),
    "pm_search_create": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "pm_search_destroy": (C.c_int, [C.c_void_p]),
    "pm_search_set_start": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pm_search_set_dummy_seed": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "pm_search_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]),
    "pm_search_fetch": (C.c_int, [C.c_void_p, C.c_int]),
    "pm_search_apply": (C.c_int, [C.c_void_p]),
    "pm_search_finish": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pm_search_cache_download": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "pm_host_alloc": (C.c_int, [C.c_void_p, C.c_uint64]),
    "pm_host_free": (C.c_int, [C.c_void_p]),
    "pm_client_download": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64]),
    "pm_l2_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]),
    "pm_l2_query": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]),
    "pm_l2_batch": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]),
    "pm_l2_batch_dev": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "pm_l2_idpairs": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "pm_l2_idpairs_dev": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "pm_ip_u32_scan": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "pm_ip_u32_scan_dev": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """Load libpacmann_cuda.so (raises if it has not been built: there is no CPU path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C pacmann_b200/csrc).  pacmann_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != PM_OK:
        raise PacmannError(rc, lib().pm_last_error().decode(errors="replace"))


def _ptr(a):
    """host numpy array / raw int address -> void*"""
    if a is None:
        return None
    if isinstance(a, (int, C.c_void_p)):
        return a
    return a.ctypes.data_as(C.c_void_p)


def _addr(a):
    """host numpy array / raw int address -> integer address (None stays None)"""
    if a is None or isinstance(a, int):
        return a
    if isinstance(a, C.c_void_p):
        return a.value
    return a.ctypes.data


def _arr(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def version():
    return lib().pm_version().decode()


def device_count():
    n = C.c_int()
    check(lib().pm_device_count(C.byref(n)))
    return n.value


def launch_count():
    return lib().pm_launch_count()


def tuning_set(name, value):
    check(lib().pm_tuning_set(name.encode(), int(value)))


def tuning_get(name):
    v = C.c_int()
    check(lib().pm_tuning_get(name.encode(), C.byref(v)))
    return v.value


class DB:
    """pm_db handle: device-resident rows[n_rows][entry_u64] (rawDB, pianopir/pir.go:28-39)."""

    def __init__(self, rows=None, n_rows=None, entry_u64=None, device=0, device_ptr=None):
        self.h = C.c_void_p()
        if device_ptr is not None:
            check(lib().pm_db_wrap(device_ptr, n_rows, entry_u64, device, C.byref(self.h)))
        elif rows is not None:
            rows = _arr(rows, np.uint64)
            if rows.ndim == 2:
                n_rows, entry_u64 = rows.shape
            assert rows.size == n_rows * entry_u64
            check(lib().pm_db_create(_ptr(rows), n_rows, entry_u64, device, C.byref(self.h)))
        else:
            check(lib().pm_db_create_empty(n_rows, entry_u64, device, C.byref(self.h)))
        self.n_rows, self.entry_u64, self.device = n_rows, entry_u64, device

    def upload(self, row0, rows):
        rows = _arr(rows, np.uint64)
        check(lib().pm_db_upload(self.h, row0, rows.size // self.entry_u64, _ptr(rows)))

    def device_ptr(self):
        p = C.c_void_p()
        check(lib().pm_db_info(self.h, None, None, None, C.byref(p)))
        return p.value

    def sync(self):
        check(lib().pm_db_sync(self.h))

    def close(self):
        if self.h:
            lib().pm_db_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def buf_alloc(nbytes, device=0):
    p = C.c_void_p()
    check(lib().pm_buf_alloc(nbytes, device, C.byref(p)))
    return p.value


def buf_free(ptr, device=0):
    check(lib().pm_buf_free(ptr, device))


def buf_ipc_export(ptr, device=0):
    h = (C.c_uint8 * 64)()
    check(lib().pm_buf_ipc_export(ptr, device, h))
    return bytes(h)


def buf_ipc_open(handle, device=0):
    h = (C.c_uint8 * 64).from_buffer_copy(handle)
    p = C.c_void_p()
    check(lib().pm_buf_ipc_open(h, device, C.byref(p)))
    return p.value


def buf_ipc_close(ptr, device=0):
    check(lib().pm_buf_ipc_close(ptr, device))


def buf_upload(ptr, host, device=0):
    host = np.ascontiguousarray(host)
    check(lib().pm_buf_upload(ptr, _ptr(host), host.nbytes, device))


def buf_download(ptr, host, device=0):
    assert host.flags.c_contiguous
    check(lib().pm_buf_download(_ptr(host), ptr, host.nbytes, device))
    return host


def buf_copy_dev(dst, src, nbytes, device=0, stream=None):
    check(lib().pm_buf_copy_dev(dst, src, nbytes, device, stream))


def buf_zero(ptr, nbytes, device=0):
    check(lib().pm_buf_zero(ptr, nbytes, device))


def flag_signal_dev(flag_ptr, device=0, stream=None):
    check(lib().pm_flag_signal_dev(flag_ptr, device, stream))


def flag_wait_dev(flags_ptr, n_flags, target, timeout_ms=10000, device=0, stream=None):
    check(lib().pm_flag_wait_dev(flags_ptr, n_flags, target & 0xFFFFFFFF, timeout_ms, device, stream))


def host_register(addr, nbytes):
    check(lib().pm_host_register(addr, nbytes))


def host_unregister(addr):
    check(lib().pm_host_unregister(addr))


def expand_key(key):
    key = np.frombuffer(bytes(key), np.uint8).copy()
    assert key.size == 16
    rk = np.zeros(44, np.uint32)
    check(lib().pm_expand_key(_ptr(key), _ptr(rk)))
    return rk


def prf_batch(rk, tags, xs):
    rk, tags, xs = _arr(rk, np.uint32), _arr(tags, np.uint64), _arr(xs, np.uint64)
    assert rk.size == 44 and tags.size == xs.size
    out = np.zeros(tags.size, np.uint64)
    check(lib().pm_prf_batch(_ptr(rk), _ptr(tags), _ptr(xs), tags.size, _ptr(out)))
    return out


def xor_slices(dst, src):
    assert dst.dtype == np.uint64 and src.dtype == np.uint64 and dst.flags.c_contiguous and src.flags.c_contiguous
    assert dst.size >= (src.size & ~3)
    check(lib().pm_xor_slices(_ptr(dst), _ptr(src), src.size))
    return dst


def make_job(row0, n_rows, chunk_size, set_size, rk, hint_begin, n_hints, n_primary, backup_group,
             tags=None, skip_chunk=None, parity_out=None, offsets_out=None):
    j = HintJob()
    j.row0, j.n_rows, j.chunk_size, j.set_size = row0, n_rows, chunk_size, set_size
    rk = _arr(rk, np.uint32)
    assert rk.size == 44
    C.memmove(j.rk, rk.ctypes.data, 176)
    j.hint_begin, j.n_hints, j.n_primary, j.backup_group = hint_begin, n_hints, n_primary, backup_group
    j.tags, j.skip_chunk, j.parity_out, j.offsets_out = _addr(tags), _addr(skip_chunk), _addr(parity_out), _addr(offsets_out)
    return j


def hintgen(db, jobs):
    """pm_hintgen over host buffers; jobs is a list of HintJob whose pointers are host addresses."""
    arr = (HintJob * len(jobs))(*jobs)
    check(lib().pm_hintgen(db.h, arr, len(jobs)))


def hintgen_dev(db, jobs, stream=None):
    arr = (HintJob * len(jobs))(*jobs)
    check(lib().pm_hintgen_dev(db.h, arr, len(jobs), stream))


def gather_rows(db, row0, n_rows, idx):
    idx = _arr(idx, np.uint64)
    out = np.zeros((idx.size, db.entry_u64), np.uint64)
    check(lib().pm_gather_rows(db.h, row0, n_rows, _ptr(idx), idx.size, _ptr(out)))
    return out


def answer_batch(db, row0, n_rows, chunk_size, set_size, offsets):
    """offsets: [q][stride] uint32 -> [q][entry_u64] uint64"""
    offsets = _arr(offsets, np.uint32)
    q, stride = offsets.shape
    row0, n_rows = _arr(np.broadcast_to(row0, (q,)), np.uint64), _arr(np.broadcast_to(n_rows, (q,)), np.uint64)
    chunk_size, set_size = _arr(np.broadcast_to(chunk_size, (q,)), np.uint32), _arr(np.broadcast_to(set_size, (q,)), np.uint32)
    out = np.zeros((q, db.entry_u64), np.uint64)
    check(lib().pm_answer_batch(db.h, _ptr(row0), _ptr(n_rows), _ptr(chunk_size), _ptr(set_size), _ptr(offsets), stride, q, _ptr(out)))
    return out


def l2_pairs(a, b, device=0):
    a, b = _arr(a, np.float32), _arr(b, np.float32)
    assert a.shape == b.shape and a.ndim == 2
    out = np.zeros(a.shape[0], np.float32)
    check(lib().pm_l2_pairs(_ptr(a), _ptr(b), a.shape[0], a.shape[1], _ptr(out), device))
    return out


def l2_query(vecs, query, device=0):
    vecs, query = _arr(vecs, np.float32), _arr(query, np.float32)
    assert vecs.ndim == 2 and query.size == vecs.shape[1]
    out = np.zeros(vecs.shape[0], np.float32)
    check(lib().pm_l2_query(_ptr(vecs), vecs.shape[0], vecs.shape[1], _ptr(query), _ptr(out), device))
    return out


def l2_batch(db, dim, queries, ids):
    queries, ids = _arr(queries, np.float32), _arr(ids, np.int64)
    nq, k = ids.shape
    assert queries.shape == (nq, dim)
    out = np.zeros((nq, k), np.float32)
    check(lib().pm_l2_batch(db.h, dim, _ptr(queries), nq, _ptr(ids), k, _ptr(out)))
    return out


def l2_idpairs(db, dim, ids_a, ids_b):
    """out[p] = L2Dist(row ids_a[p], row ids_b[p]) over the resident table"""
    ids_a, ids_b = _arr(ids_a, np.int64).reshape(-1), _arr(ids_b, np.int64).reshape(-1)
    assert ids_a.size == ids_b.size
    out = np.zeros(ids_a.size, np.float32)
    check(lib().pm_l2_idpairs(db.h, dim, _ptr(ids_a), _ptr(ids_b), ids_a.size, _ptr(out)))
    return out


def ip_u32_scan(db, dim, queries, want_products=False):
    queries = _arr(queries, np.uint32).reshape(-1, dim)
    nq = queries.shape[0]
    cs = np.zeros(nq, np.uint32)
    ip = np.zeros((nq, db.n_rows), np.uint32) if want_products else None
    check(lib().pm_ip_u32_scan(db.h, dim, _ptr(queries), nq, _ptr(cs), _ptr(ip)))
    return (cs, ip) if want_products else cs
