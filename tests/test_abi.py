"""CPU tests of the drop-in boundary: the shared libraries load, export every symbol include/pacmann_cuda.h
declares with the binding's signature table in sync, and fail loudly (never fall back) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "pacmann_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pm_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound(cabi):
    syms = header_symbols()
    assert len(syms) >= 20
    L = ctypes.CDLL(cabi.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in pacmann_cuda.h but not exported by libpacmann_cuda.so"
        assert s in cabi.SIGNATURES, f"{s} has no ctypes signature in pacmann_b200/cabi.py"
    assert sorted(cabi.SIGNATURES) == syms


def test_hint_job_struct_layout_matches_header(cabi):
    # struct pm_hint_job: 4 u64, 44 u32, 4 u64, 4 pointers (offsets_out, the optional offset-index output, is the last)
    assert ctypes.sizeof(cabi.HintJob) == 4 * 8 + 44 * 4 + 4 * 8 + 4 * 8
    assert cabi.HintJob.rk.offset == 32 and cabi.HintJob.hint_begin.offset == 32 + 176
    assert cabi.HintJob.parity_out.offset == 32 + 176 + 32 + 16
    assert cabi.HintJob.offsets_out.offset == 32 + 176 + 32 + 24


def test_version_and_host_library_load(cabi):
    assert "sm_100a" in cabi.version()
    from pacmann_b200 import _host
    L = _host.lib()
    from pacmann_b200.keys import derive_key, mix64
    assert L.pmh_mix64(7, 9) == mix64(7, 9)
    out = (ctypes.c_uint8 * 16)()
    L.pmh_derive_key(1, 2, 16, 3, out)
    assert bytes(out) == derive_key(1, 2, 16, 3)


def test_key_derivation_matches_oracle(oracle):
    from pacmann_b200.keys import derive_key, mix64
    for seed, ctr in [(0, 0), (1, 2), (2**64 - 1, 2**63)]:
        assert mix64(seed, ctr) == oracle.mix64(seed, ctr)
    assert derive_key(5, 1, 16, 7) == oracle.derive_key(5, 1, 16, 7)


def test_host_prf_matches_golden(oracle):
    # the host mirror's online-path PRF (AES-NI or portable) against the oracle on random inputs
    from pacmann_b200 import pianopir
    rng = np.random.default_rng(1)
    rk = oracle.expand_key(bytes(range(16)))
    for _ in range(200):
        tag, x = int(rng.integers(0, 2**29)), int(rng.integers(0, 2**20))
        assert pianopir.PRFEvalWithLongKeyAndTag(rk, tag, x) == oracle.prf(rk, tag, x)


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu(cabi):
    """Without a CUDA device every compute entry point must raise PM_ERR_CUDA: there is no CPU path."""
    with pytest.raises(cabi.PacmannError) as e:
        cabi.expand_key(bytes(16))
    assert e.value.code == cabi.PM_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(cabi.PacmannError):
        cabi.DB(np.zeros((4, 4), np.uint64))
    with pytest.raises(cabi.PacmannError):
        cabi.l2_pairs(np.zeros((1, 8), np.float32), np.zeros((1, 8), np.float32))
    from pacmann_b200 import pianopir
    with pytest.raises(Exception):
        pianopir.NewPianoPIR(4, 32, np.zeros(16, np.uint64), 8)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under pacmann_b200/ or the C-ABI sources may reference it."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "pacmann_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"\boracle\b", txt) and not f == "__init__.py":
                    for line in txt.splitlines():
                        if re.search(r"(import|include|from|CDLL|dlopen|-l).*oracle", line):
                            bad.append((f, line.strip()))
    assert not bad, bad


@pytest.mark.parametrize("threads", [1, 2, 5])
def test_host_building_blocks_selftest(threads):
    """The host mirror's open-addressing map (against std::unordered_map under random operations) and its worker pool
    (every index exactly once over thousands of loops, exception propagation): no GPU involved."""
    from pacmann_b200 import _host
    assert _host.lib().pmh_selftest(threads) == 0


def test_tuning_knobs_need_no_gpu(cabi):
    """launch knobs are plain process state: readable and settable without a device, unknown names are errors"""
    for name in ("hg_sync", "hg_warps", "hg_ntab", "hg_tail_split", "hg_serpentine", "hg_xbytes", "hg_d2h_groups", "ans_split", "search_ans_stream"):
        old = cabi.tuning_get(name)
        cabi.tuning_set(name, 3)
        assert cabi.tuning_get(name) == 3
        cabi.tuning_set(name, old)
        assert cabi.tuning_get(name) == old
    with pytest.raises(cabi.PacmannError) as e:
        cabi.tuning_set("no_such_knob", 1)
    assert e.value.code == cabi.PM_ERR_ARG
