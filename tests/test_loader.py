"""On-disk formats of graphann/loader.go and the entry wire format of private-search.go:355-439.
CPU only.  The wire format is checked against the oracle's packing (the one the parity tests upload)."""
import numpy as np
import pytest

from pacmann_b200 import loader


def write_xvecs(path, a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    with open(path, "wb") as fh:
        for row in a:
            fh.write(np.int32(a.shape[1]).tobytes())
            fh.write(row.tobytes())


def test_bvecs_fvecs_ivecs(tmp_path):
    rng = np.random.default_rng(1)
    b = rng.integers(0, 256, (7, 16), dtype=np.uint8)
    write_xvecs(tmp_path / "v.bvecs", b, np.uint8)
    got = loader.LoadFloat32Matrix(str(tmp_path / "v.bvecs"), 5, 16)
    assert got.dtype == np.float32 and (got == b[:5].astype(np.float32)).all()        # byte value, not /255 (loader.go:47-51)
    f = rng.standard_normal((6, 8)).astype(np.float32)
    write_xvecs(tmp_path / "v.fvecs", f, np.float32)
    assert (loader.LoadFloat32Matrix(str(tmp_path / "v.fvecs"), 6, 8).view(np.uint32) == f.view(np.uint32)).all()
    g = rng.integers(0, 2**31, (4, 32)).astype(np.uint32)
    write_xvecs(tmp_path / "g.ivecs", g, np.uint32)
    assert (loader.LoadGraphFromFile(str(tmp_path / "g.ivecs"), 4, 32) == g.astype(np.int32)).all()
    with pytest.raises(loader.LoaderError):
        loader.LoadFloat32Matrix(str(tmp_path / "v.fvecs"), 7, 8)      # fewer records than asked for
    with pytest.raises(loader.LoaderError):
        loader.LoadFloat32Matrix(str(tmp_path / "v.fvecs"), 3, 9)      # wrong record width
    with pytest.raises(loader.LoaderError):
        loader.LoadFloat32Matrix(str(tmp_path / "v.xyz"), 1, 1)


def test_npy_and_txt(tmp_path):
    rng = np.random.default_rng(2)
    v = rng.standard_normal((9, 12))
    np.save(tmp_path / "v.npy", v)                                       # float64, as the reference expects
    got = loader.LoadFloat32Matrix(str(tmp_path / "v.npy"), 8, 12)
    assert got.shape == (8, 12) and (got == v[:8].astype(np.float32)).all()
    np.save(tmp_path / "v32.npy", v.astype(np.float32))
    with pytest.raises(loader.LoaderError):
        loader.LoadFloat32Matrix(str(tmp_path / "v32.npy"), 8, 12)       # GetFloat64 fails on float32 files
    with pytest.raises(loader.LoaderError):
        loader.LoadFloat32Matrix(str(tmp_path / "v.npy"), 10, 12)        # shape[0] < n
    g = rng.integers(0, 1000, (9, 5)).astype(np.int32)
    for name in ("g.npy", "g.txt"):
        loader.SaveGraphToFile(str(tmp_path / name), g)
        assert (loader.LoadGraphFromFile(str(tmp_path / name), 9, 5) == g).all()
    assert open(tmp_path / "g.txt").readline() == "".join(f"{x} " for x in g[0]) + "\n"   # "%d " per value (loader.go:340)
    np.save(tmp_path / "g64.npy", g.astype(np.int64))
    with pytest.raises(loader.LoaderError):
        loader.LoadGraphFromFile(str(tmp_path / "g64.npy"), 9, 5)
    with open(tmp_path / "v.txt", "w") as fh:
        for row in v[:4]:
            fh.write(" ".join(repr(float(x)) for x in row) + "\n")
    assert (loader.LoadFloat32Matrix(str(tmp_path / "v.txt"), 4, 12) == v[:4].astype(np.float32)).all()


def test_entry_wire_format_matches_oracle(oracle):
    rng = np.random.default_rng(3)
    n, dim, m = 50, 24, 8
    v = rng.standard_normal((n, dim)).astype(np.float32)
    g = rng.integers(0, n, (n, m)).astype(np.int32)
    raw = loader.pack_db(v, g)
    assert raw.dtype == np.uint64 and raw.size == n * (dim + m) // 2
    assert (raw == np.asarray(oracle.pack_db(v, g)).reshape(-1)).all()
    vec, nbr = loader.unpack_entry(raw.reshape(n, -1)[17], dim, m)
    assert (vec.view(np.uint32) == v[17].view(np.uint32)).all() and (nbr == g[17]).all()
    with pytest.raises(loader.LoaderError):
        loader.pack_db(v[:, :23], g)                                     # odd dim + m
