"""Scratch timing of pm_hintgen_dev at a BASELINE.json shape (device-resident, CUDA events)."""
import argparse
import ctypes as C
import math
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pacmann_b200 import cabi


def params(n, fail_log2):
    target = int(2 * math.sqrt(n))
    c = 1
    while c < target:
        c *= 2
    s = (math.ceil(n / c) + 3) // 4 * 4
    maxq = int(math.sqrt(n) * math.log(n))
    p = math.ceil(math.log(2) * (fail_log2 + 1)) * c
    p = (p + 7) // 8 * 8
    mq = 3 * int(maxq / s)
    mq = (mq + 7) // 8 * 8
    return c, s, p, mq


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=3201821)
    ap.add_argument("--entry-u64", type=int, default=112)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--fail", type=int, default=8)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--shard", default="0/1", help="r/N: time rank r's share of the hints of an N-GPU run")
    ap.add_argument("--sweep", default="", help="knob=v1,v2,...: time every value in this process (same DB) and check that the parities are bit-identical")
    a = ap.parse_args()
    torch.cuda.init()
    E = a.entry_u64
    parts = a.batch // 2 if a.batch else 1
    ps = (a.n + parts - 1) // parts
    g = torch.Generator(device="cuda").manual_seed(1)
    nbytes = a.n * E * 8
    buf = torch.randint(-2**62, 2**62, (a.n * E,), dtype=torch.int64, device="cuda", generator=g)
    db = cabi.DB(n_rows=a.n, entry_u64=E, device=0, device_ptr=buf.data_ptr())
    rk = (np.arange(44, dtype=np.uint64) * 2654435761 % (2**32)).astype(np.uint32)
    jobs, outs, total_h, prf = [], [], 0, 0
    for i in range(parts):
        n_i = min(ps, a.n - i * ps)
        c, s, p, mq = params(n_i, a.fail)
        H = p + s * mq
        out = torch.empty(H * E, dtype=torch.int64, device="cuda")
        outs.append(out)
        r_, n_ = (int(x) for x in a.shard.split("/"))
        hb, he = H * r_ // n_, H * (r_ + 1) // n_
        jobs.append(cabi.make_job(i * ps, n_i, c, s, rk.astype(np.uint32), hb, he - hb, p, mq, parity_out=out.data_ptr()))
        total_h += he - hb
        prf += s * (he - hb)
    print(f"parts={parts} chunk={c} set={s} primary={p} mqpc={mq} hints/part={H} prf={prf} xor_bytes={prf*E*8/1e9:.3f} GB db={nbytes/1e9:.3f} GB")
    stream = torch.cuda.Stream()
    st = stream.cuda_stream
    knob, values = (a.sweep.split("=")[0], [int(v) for v in a.sweep.split("=")[1].split(",")]) if a.sweep else (None, [None])
    ref = None
    for v in values:
        if knob:
            cabi.tuning_set(knob, v)
        with torch.cuda.stream(stream):
            for _ in range(2):
                cabi.hintgen_dev(db, jobs, st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for _ in range(a.iters):
                e0.record(stream)
                cabi.hintgen_dev(db, jobs, st)
                e1.record(stream)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        same = ""
        if knob:
            if ref is None:
                ref = [o.clone() for o in outs]
            else:
                same = "  parities identical: %s" % all(torch.equal(x, y) for x, y in zip(ref, outs))
        print(f"{knob}={v} " if knob else "", f"hintgen: {ms:.3f} ms (median {sorted(ts)[len(ts)//2]:.3f}, all {['%.3f' % x for x in ts]})  db-scan {nbytes/ms/1e6:.1f} GB/s  prf {prf/ms/1e6:.2f} G/s  xor-gather {prf*E*8/ms/1e6:.1f} GB/s{same}", flush=True)


if __name__ == "__main__":
    main()
