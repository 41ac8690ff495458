"""The data formats on either side of the hot path: the on-disk formats graphann/loader.go reads
(vectors: .bvecs / .fvecs / .txt / .npy float64 2-D; graphs: .npy int32 2-D / .txt / .ivecs) and the entry wire format
private-search.go:355-439 packs them into (what is uploaded as rawDB).  numpy only: this is I/O, not arithmetic.

Semantics follow the reference function by function: the first `n` records are used, a file with fewer records or a
wrong record width is an error (the Go code prints and returns what it has; here it raises), .bvecs bytes become the
float32 of their value (no scaling, loader.go:47-51), .npy vectors must be float64 and are narrowed to float32
(loader.go:163-195), .npy graphs must be int32 (loader.go:217-248)."""
import os

import numpy as np


class LoaderError(ValueError):
    pass


def _load_xvecs(filename, n, dim, dtype):
    """[int32 dim][dim x dtype] records (kshard/fvecs decoder as loader.go uses it)."""
    rec = 4 + dim * np.dtype(dtype).itemsize
    raw = np.fromfile(filename, dtype=np.uint8, count=n * rec)
    if raw.size < n * rec:
        raise LoaderError(f"{filename}: {raw.size // rec} records of dim {dim}, {n} wanted")
    raw = raw.reshape(n, rec)
    dims = raw[:, :4].copy().view("<i4")[:, 0]
    if (dims != dim).any():
        raise LoaderError(f"{filename}: record dimension {int(dims[dims != dim][0])} != {dim}")
    return raw[:, 4:].copy().view(np.dtype(dtype).newbyteorder("<")).reshape(n, dim)


def LoadFloat32MatrixFromBvecs(filename, n, dim):   # loader.go:16-63
    return _load_xvecs(filename, n, dim, np.uint8).astype(np.float32)


def LoadFloat32MatrixFromFvecs(filename, n, dim):   # loader.go:65-89
    return np.ascontiguousarray(_load_xvecs(filename, n, dim, np.float32), dtype=np.float32)


def LoadIntMatrixFromIvecs(filename, n, dim):       # loader.go:91-121
    return _load_xvecs(filename, n, dim, np.uint32).astype(np.int64)


def _load_txt(filename, n, dim, conv, dtype):
    out = np.zeros((n, dim), dtype)
    with open(filename) as fh:
        i = 0
        for line in fh:
            if i >= n:
                break
            fields = line.split()
            if len(fields) < dim:
                raise LoaderError(f"{filename}: line {i + 1} has {len(fields)} fields, {dim} wanted")
            out[i] = [conv(x) for x in fields[:dim]]
            i += 1
    if i < n:
        raise LoaderError(f"{filename}: {i} lines, {n} wanted")
    return out


def LoadFloat32MatrixFromTxt(filename, n, dim):     # loader.go:122-160 (ParseFloat(.., 32))
    return _load_txt(filename, n, dim, np.float32, np.float32)


def LoadFloat32MatrixFromNpy(filename, n, dim):     # loader.go:163-195
    a = np.load(filename, mmap_mode="r")
    if a.ndim != 2 or a.shape[0] < n or a.shape[1] != dim:
        raise LoaderError(f"{filename}: invalid shape {a.shape}, expected ({n}, {dim})")
    if a.dtype != np.float64:
        raise LoaderError(f"{filename}: dtype {a.dtype}, the reference reads float64 (gonpy GetFloat64)")
    return np.asarray(a[:n], dtype=np.float32)


def LoadFloat32Matrix(filename, n, dim):            # loader.go:197-215
    ext = os.path.splitext(filename)[1]
    fn = {".bvecs": LoadFloat32MatrixFromBvecs, ".fvecs": LoadFloat32MatrixFromFvecs, ".txt": LoadFloat32MatrixFromTxt,
          ".npy": LoadFloat32MatrixFromNpy}.get(ext)
    if fn is None:
        raise LoaderError(f"unsupported file extension: {ext}")
    return fn(filename, n, dim)


def LoadGraphFromNpyFile(filename, n, m):           # loader.go:217-248
    a = np.load(filename, mmap_mode="r")
    if a.ndim != 2 or a.shape[0] < n or a.shape[1] != m:
        raise LoaderError(f"{filename}: invalid shape {a.shape}")
    if a.dtype != np.int32:
        raise LoaderError(f"{filename}: dtype {a.dtype}, the reference reads int32 (gonpy GetInt32)")
    return np.asarray(a[:n], dtype=np.int32)


def LoadGraphFromTxtFile(filename, n, m):           # loader.go:250-285
    return _load_txt(filename, n, m, int, np.int64).astype(np.int32)


def LoadGraphFromFile(filename, n, m):              # loader.go:287-300
    ext = os.path.splitext(filename)[1]
    if ext == ".npy":
        return LoadGraphFromNpyFile(filename, n, m)
    if ext == ".txt":
        return LoadGraphFromTxtFile(filename, n, m)
    if ext == ".ivecs":
        return LoadIntMatrixFromIvecs(filename, n, m).astype(np.int32)
    raise LoaderError(f"unsupported file extension: {ext}")


LoadIntMatrixFromFile = LoadGraphFromFile           # loader.go:302-304


def SaveGraphToNpyFile(filename, graph):            # loader.go:306-326: int32, 2-D
    with open(filename, "wb") as fh:
        np.save(fh, np.ascontiguousarray(graph, dtype=np.int32))


def SaveGraphToTxtFile(filename, graph):            # loader.go:328-347: "%d " per value, one row per line
    with open(filename, "w") as fh:
        for row in np.asarray(graph):
            fh.write("".join(f"{int(v)} " for v in row) + "\n")


def SaveGraphToFile(filename, graph):               # loader.go:349-360
    ext = os.path.splitext(filename)[1]
    if ext == ".npy":
        return SaveGraphToNpyFile(filename, graph)
    if ext == ".txt":
        return SaveGraphToTxtFile(filename, graph)
    raise LoaderError(f"unsupported file extension: {ext}")


SaveIntMatrixToFile = SaveGraphToFile


# ---- entry wire format (private-search.go:355-439) -------------------------------------------------------------
def pack_db(vectors, graph):
    """rawDB as PIRGraphInfo.Preprocess builds it (private-search.go:371-397): entry i = the `dim` little-endian float32
    bit patterns of vectors[i] followed by the `m` neighbour ids of graph[i] as little-endian uint32, viewed as
    (dim + m) / 2 little-endian uint64.  Returns a flat uint64 array of n * (dim + m) / 2 words."""
    v = np.ascontiguousarray(vectors, dtype="<f4")
    g = np.ascontiguousarray(graph)
    n, dim = v.shape
    m = g.shape[1]
    if g.shape[0] != n:
        raise LoaderError("vectors and graph have different row counts")
    if (dim + m) % 2:
        raise LoaderError("dim + m must be even: an entry is a whole number of uint64 words (private-search.go:362-366)")
    raw = np.empty((n, dim + m), dtype="<u4")
    raw[:, :dim] = v.view("<u4")
    raw[:, dim:] = g.astype(np.int64).astype("<u4")      # uint32(neighbour id), as the Go conversion
    return raw.reshape(-1).view("<u8")


def unpack_entry(entry, dim, m):
    """Entry2VectorAndNeighbors (private-search.go:418-439): one entry -> (float32[dim], int64[m])."""
    w = np.ascontiguousarray(entry, dtype="<u8").view("<u4")
    if w.size < dim + m:
        raise LoaderError("entry too short")
    return w[:dim].view("<f4").astype(np.float32), w[dim:dim + m].astype(np.int64)
