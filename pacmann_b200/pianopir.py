"""Python face of the host-side mirror of the reference's `pianopir` Go package.

Same names and argument meaning as pianopir/pir.go and pianopir/batch-pir.go so that tests read like the
reference's own (pianopir/pir_test.go).  The logic lives in C++ (csrc/host/pianopir.cpp) and calls the
CUDA kernels through the C-ABI; nothing here computes on the CPU and nothing falls back to it.
"""
import ctypes as C

import numpy as np

from . import _host

DefaultProgramPoint = 0x7FFFFFFF
QueryPerPartition = 2
RealQueryPerPartition = 2

# error codes of PianoPIRClient.Query (the Go code returns error strings, pir.go:377,390,399,418)
ERR_NONE, ERR_OUT_OF_RANGE, ERR_BUDGET, ERR_TOO_MANY_IN_CHUNK, ERR_NO_HIT_HINT = 0, 1, 2, 3, 4

_GET = dict(DBEntrySize=0, DBSize=1, ChunkSize=2, SetSize=3, MaxQueryNum=4, primaryHintNum=5, maxQueryPerChunk=6,
            FinishedQueryNum=7, ThreadNum=8, FailureProbLog2=9)
_TABLE = dict(primaryShortTag=0, primaryParity=1, primaryProgramPoint=2, replacementIdx=3, replacementVal=4,
              backupShortTag=5, backupParity=6, QueryHistogram=7)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class PianoPIRConfig:
    def __init__(self, pir):
        for k in ("DBEntrySize", "DBSize", "ChunkSize", "SetSize", "ThreadNum", "FailureProbLog2"):
            setattr(self, k, pir._get(k))
        self.DBEntryByteNum = self.DBEntrySize * 8

    def __repr__(self):
        return f"PianoPIRConfig({self.__dict__})"


class PianoPIR:
    """pianopir.PianoPIR (pir.go:473-548)."""

    def __init__(self, DBSize, DBEntryByteNum, rawDB, FailureProbLog2, device=0, _borrow=None, _keep=None):
        L = _host.lib()
        if _borrow is not None:
            self.h, self._boxed, self._keep = C.c_void_p(_borrow), 0, _keep
        else:
            rawDB = np.ascontiguousarray(rawDB, np.uint64).reshape(-1)
            if rawDB.size != DBSize * (DBEntryByteNum // 8):
                raise ValueError(f"Piano PIR len(rawDB) = {rawDB.size}; want {DBSize * (DBEntryByteNum // 8)}")  # pir.go:483-485
            h = L.pmh_pir_new(DBSize, DBEntryByteNum, _p(rawDB), FailureProbLog2, device)
            if not h:
                raise _host.HostError(L.pmh_last_error().decode())
            self.h, self._boxed, self._keep = C.c_void_p(h), 1, None

    def __del__(self):
        if getattr(self, "_boxed", 0) and self.h:
            _host.lib().pmh_pir_free(self.h)
            self.h = None

    def _get(self, name):
        return _host.lib().pmh_pir_get(self.h, self._boxed, _GET[name])

    # -- reference API --
    def Config(self):
        return PianoPIRConfig(self)

    def Preprocessing(self):
        _host.check(_host.lib().pmh_pir_preprocessing(self.h, self._boxed))

    def DummyPreprocessing(self):
        _host.check(_host.lib().pmh_pir_dummy_preprocessing(self.h, self._boxed))

    def Query(self, idx, realQuery=True):
        """returns (entry, err) like the Go method; err is an ERR_* code (0 = nil)"""
        out = np.zeros(self._get("DBEntrySize"), np.uint64)
        rc = _host.check(_host.lib().pmh_pir_query(self.h, self._boxed, int(idx), int(realQuery), _p(out)))
        return out, rc

    def LocalStorageSize(self):
        return _host.lib().pmh_pir_local_storage(self.h, self._boxed)

    def CommCostPerQuery(self):
        return _host.lib().pmh_pir_comm_cost(self.h, self._boxed)

    # -- server side (pir.go:41-88) --
    def PrivateQuery(self, offsets):
        offsets = np.ascontiguousarray(offsets, np.uint32)
        out = np.zeros(self._get("DBEntrySize"), np.uint64)
        _host.check(_host.lib().pmh_pir_private_query(self.h, self._boxed, _p(offsets), _p(out)))
        return out

    def NonePrivateQuery(self, idx):
        out = np.zeros(self._get("DBEntrySize"), np.uint64)
        rc = _host.check(_host.lib().pmh_pir_nonprivate_query(self.h, self._boxed, int(idx), _p(out)))
        return out, rc

    # -- test hooks: deterministic seeds and table access (the Go fields are package-private) --
    def SetSeeds(self, key_seed, epoch=0, repl_seed=0):
        _host.lib().pmh_pir_set_seeds(self.h, self._boxed, key_seed, epoch, repl_seed)

    def client(self, name):
        return self._get(name)

    def table(self, name):
        E, P, S, M = (self._get(k) for k in ("DBEntrySize", "primaryHintNum", "SetSize", "maxQueryPerChunk"))
        shape = dict(primaryShortTag=(P,), primaryParity=(P, E), primaryProgramPoint=(P,), replacementIdx=(S, M),
                     replacementVal=(S, M, E), backupShortTag=(S, M), backupParity=(S, M, E), QueryHistogram=(S,))[name]
        ptr = _host.lib().pmh_pir_table(self.h, self._boxed, _TABLE[name])
        return np.ctypeslib.as_array(ptr, shape=shape)

    def long_key(self):
        return np.ctypeslib.as_array(_host.lib().pmh_pir_long_key(self.h, self._boxed), shape=(44,)).copy()


def NewPianoPIR(DBSize, DBEntryByteNum, rawDB, FailureProbLog2, device=0):
    return PianoPIR(DBSize, DBEntryByteNum, rawDB, FailureProbLog2, device)


class SimpleBatchPianoPIRConfig:
    pass


class SimpleBatchPianoPIR:
    """pianopir.SimpleBatchPianoPIR (batch-pir.go:40-276)."""

    def __init__(self, DBSize, DBEntryByteNum, BatchSize, rawDB, FailureProbLog2, device=0, _borrow=None):
        L = _host.lib()
        self._owned = _borrow is None
        if _borrow is not None:
            self.h = C.c_void_p(_borrow)
        else:
            rawDB = np.ascontiguousarray(rawDB, np.uint64).reshape(-1)
            h = L.pmh_batch_new(DBSize, DBEntryByteNum, BatchSize, _p(rawDB), rawDB.size, FailureProbLog2, device)
            if not h:
                raise ValueError(L.pmh_last_error().decode())   # batch-pir.go:57-59 log.Fatalf
            self.h = C.c_void_p(h)
        self.DBSize, self.DBEntryByteNum, self.BatchSize = DBSize, DBEntryByteNum, BatchSize
        self.DBEntrySize = DBEntryByteNum // 8

    def __del__(self):
        if getattr(self, "_owned", False) and self.h:
            _host.lib().pmh_batch_free(self.h)
            self.h = None

    def _get(self, what):
        return _host.lib().pmh_batch_get(self.h, what)

    def Config(self):
        c = SimpleBatchPianoPIRConfig()
        c.DBEntryByteNum, c.DBEntrySize, c.DBSize, c.BatchSize = self.DBEntryByteNum, self.DBEntrySize, self.DBSize, self.BatchSize
        c.PartitionNum, c.PartitionSize, c.ThreadNum = self._get(0), self._get(1), 1
        return c

    def SetSeeds(self, key_seed, repl_seed=0):
        _host.lib().pmh_batch_set_seeds(self.h, key_seed, repl_seed)

    def EnableResidentClient(self):
        """keep the hint tables in HBM and run the online client on the GPU (pm_client_*)"""
        _host.check(_host.lib().pmh_batch_enable_resident(self.h))

    def Preprocessing(self):
        _host.check(_host.lib().pmh_batch_preprocessing(self.h))

    def DummyPreprocessing(self):
        _host.check(_host.lib().pmh_batch_dummy_preprocessing(self.h))

    def Query(self, idx):
        """returns (responses [len(idx)][DBEntrySize], err)"""
        idx = np.ascontiguousarray(idx, np.uint64)
        out = np.zeros((idx.size, self.DBEntrySize), np.uint64)
        rc = _host.check(_host.lib().pmh_batch_query(self.h, _p(idx), idx.size, _p(out)))
        if rc != 0:
            raise IndexError("index out of range")   # the Go code panics on partitionQueries[partitionIdx]
        return out, None

    def subPIR(self, i):
        h = _host.lib().pmh_batch_sub(self.h, i)
        if not h:
            raise _host.HostError(_host.lib().pmh_last_error().decode())
        return PianoPIR(0, 0, None, 0, _borrow=h, _keep=self)

    FinishedBatchNum = property(lambda s: s._get(2))
    QueriesMadeInPartition = property(lambda s: s._get(3))
    SupportBatchNum = property(lambda s: s._get(4))
    serverQueries = property(lambda s: s._get(5))
    serverLaunches = property(lambda s: s._get(6))

    def LocalStorageSize(self):
        return _host.lib().pmh_batch_local_storage(self.h)

    def CommCostPerBatchOnline(self):
        return self._get(7)

    def CommCostPerBatchOffline(self):
        return self._get(8)

    def PreprocessingTime(self):
        return _host.lib().pmh_batch_prep_time(self.h)

    def PreprocessingTotal(self):
        """(seconds spent in all Preprocessing() calls so far, their number): the maintenance the reference reports separately"""
        return _host.lib().pmh_batch_prep_total(self.h), _host.lib().pmh_batch_prep_count(self.h)


def NewSimpleBatchPianoPIR(DBSize, DBEntryByteNum, BatchSize, rawDB, FailureProbLog2, device=0):
    return SimpleBatchPianoPIR(DBSize, DBEntryByteNum, BatchSize, rawDB, FailureProbLog2, device)


def GetLongKey(key):
    """util.go:167-171 (runs the key schedule on the GPU through pm_expand_key)"""
    from . import cabi
    return cabi.expand_key(key)


def PRFEvalWithLongKeyAndTag(longKey, tag, x):
    """util.go:157-165, single evaluation (host AES, as the online client uses it)"""
    lk = np.ascontiguousarray(longKey, np.uint32)
    return _host.lib().pmh_prf(_p(lk), tag, x)
