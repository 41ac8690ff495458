"""Device-resident timing (CUDA events, dedicated stream) of the non-headline kernels against their rooflines:
server Answer (pm_answer_batch_dev), L2 batch (pm_l2_batch_dev), uint32 inner-product scan (pm_ip_u32_scan_dev)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pacmann_b200 import cabi

PEAK = 6548.8
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, stream, iters=10, warm=3):
    torch.cuda.synchronize()   # inputs were produced on torch's default stream
    with torch.cuda.stream(stream):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


def main():
    torch.cuda.init()
    stream = torch.cuda.Stream()
    st = stream.cuda_stream
    res = {}
    for name, n, E, dim in (("msmarco", 3201821, 112, 192), ("sift", 1000000, 80, 128)):
        g = torch.Generator(device="cuda").manual_seed(1)
        buf = torch.randint(-2**62, 2**62, (n * E,), dtype=torch.int64, device="cuda", generator=g)
        # make the leading `dim` floats of every row finite (distance kernel input)
        f = buf.view(torch.float32).view(n, 2 * E)
        f[:, :dim] = torch.randn(n, dim, device="cuda")
        db = cabi.DB(n_rows=n, entry_u64=E, device=0, device_ptr=buf.data_ptr())
        parts = 16
        ps = (n + parts - 1) // parts
        import math
        c = 1
        while c < int(2 * math.sqrt(ps)):
            c *= 2
        s = (math.ceil(ps / c) + 3) // 4 * 4
        for q in (96, 96 * 1000):
            part = torch.randint(0, parts, (q,), device="cuda")
            row0 = (part * ps).to(torch.int64)
            nrows = torch.minimum(torch.full_like(row0, ps), n - row0)
            chunk = torch.full((q,), c, dtype=torch.int32, device="cuda")
            sets = torch.full((q,), s, dtype=torch.int32, device="cuda")
            offs = torch.randint(0, c, (q, s), dtype=torch.int32, device="cuda")
            out = torch.empty(q * E, dtype=torch.int64, device="cuda")
            fn = lambda: cabi.check(cabi.lib().pm_answer_batch_dev(db.h, row0.data_ptr(), nrows.data_ptr(), chunk.data_ptr(), sets.data_ptr(),
                                                                   offs.data_ptr(), s, q, out.data_ptr(), st))
            med, best = timeit(fn, stream)
            byt = q * (s * E * 8 + s * 4 + E * 8)
            res[f"answer_{name}_q{q}"] = dict(ms=med, best_ms=best, gbs=byt / med / 1e6, frac_hbm=byt / med / 1e6 / PEAK, bytes=byt)
        for nq, k in ((1, 1789 if name == "msmarco" else 1000), (1000, 96)):
            queries = torch.randn(nq, dim, device="cuda")
            ids = torch.randint(0, n, (nq, k), dtype=torch.int64, device="cuda")
            out = torch.empty(nq * k, dtype=torch.float32, device="cuda")
            fn = lambda: cabi.check(cabi.lib().pm_l2_batch_dev(db.h, dim, queries.data_ptr(), nq, ids.data_ptr(), k, out.data_ptr(), st))
            med, best = timeit(fn, stream)
            byt = nq * k * dim * 4
            res[f"l2_{name}_q{nq}_k{k}"] = dict(ms=med, best_ms=best, gbs=byt / med / 1e6, frac_hbm=byt / med / 1e6 / PEAK, bytes=byt)
        db.close()
        del buf, f
    # cfg5: uint32 inner-product scan, 3 201 821 x 192
    n, d = 3201821, 192
    rows = torch.randint(0, 2**31, (n, d), dtype=torch.int32, device="cuda")
    db = cabi.DB(n_rows=n, entry_u64=d // 2, device=0, device_ptr=rows.data_ptr())
    for nq in (1, 16, 1000):
        qs = torch.randint(0, 2**31, (nq, d), dtype=torch.int32, device="cuda")
        cs = torch.empty(nq, dtype=torch.int32, device="cuda")
        fn = lambda: cabi.check(cabi.lib().pm_ip_u32_scan_dev(db.h, d, qs.data_ptr(), nq, cs.data_ptr(), None, st))
        med, best = timeit(fn, stream, iters=5 if nq > 100 else 10, warm=2)
        passes = (nq + 15) // 16 if nq > 4 else 1
        byt = passes * n * d * 4
        res[f"ip_scan_q{nq}"] = dict(ms=med, best_ms=best, gbs=byt / med / 1e6, frac_hbm=byt / med / 1e6 / PEAK, bytes=byt,
                                      gmacs=n * d * nq / med / 1e6)
    for k, v in res.items():
        print(k, json.dumps({a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()}))


if __name__ == "__main__":
    main()
