"""Row-sharded uint32 linear scan over N GPUs: rank r keeps rows [N*r/W, N*(r+1)/W) resident, scans them
(pm_ip_u32_scan_dev: tensor-core GEMM for many queries, integer pipe for few) and the per-query checksums are all-reduced
over NCCL.  cfg5 shape: 3 201 821 x 192 uint32, v[i][j] = i + j, q_t[j] = j + t (graphann_test.go:258-266 for t = 0);
rank 0 checks every checksum against the closed form.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P scripts/scan_multi_gpu.py [--q 1000]
(W = 1 works without torchrun.)  Prints one JSON line: ms per scan (CUDA events, max over ranks), u32 MAC/s."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pacmann_b200 import cabi, sharding


def closed_form(n, d, nq):
    """sum_i sum_j (i + j) * (j + t) mod 2^32, exact in Python integers"""
    out = []
    s_i = n * (n - 1) // 2
    for t in range(nq):
        sq = sum(j + t for j in range(d))
        sjq = sum(j * (j + t) for j in range(d))
        out.append((s_i * sq + n * sjq) % (1 << 32))
    return np.array(out, dtype=np.uint64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=3201821)
    ap.add_argument("--dim", type=int, default=192)
    ap.add_argument("--q", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = sharding.row_shard(a.n, rank, world)
    i = torch.arange(lo, hi, device=dev, dtype=torch.int64)[:, None]
    j = torch.arange(a.dim, device=dev, dtype=torch.int64)[None, :]
    rows = ((i + j) & 0xFFFFFFFF).to(torch.int64)
    rows = (rows - ((rows >> 31) << 32)).to(torch.int32).contiguous()           # uint32 bit patterns in an int32 tensor
    t = torch.arange(a.q, device=dev, dtype=torch.int64)[:, None]
    qs = (j + t).to(torch.int32).contiguous()
    db = cabi.DB(n_rows=hi - lo, entry_u64=a.dim // 2, device=local, device_ptr=rows.data_ptr())
    part32 = torch.empty(a.q, dtype=torch.int32, device=dev)
    total = torch.empty(a.q, dtype=torch.int64, device=dev)
    stream = torch.cuda.Stream()   # a stream of our own: handle entry points read stream 0 as "the handle's own stream"
    torch.cuda.synchronize()

    def scan():
        cabi.check(cabi.lib().pm_ip_u32_scan_dev(db.h, a.dim, qs.data_ptr(), a.q, part32.data_ptr(), None, stream.cuda_stream))
        total.copy_(part32.to(torch.int64) & 0xFFFFFFFF)
        if world > 1:
            sharding.allreduce_checksums(dist, total)

    ts = []
    with torch.cuda.stream(stream):
        scan()
        torch.cuda.synchronize()
        for _ in range(a.iters):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            scan()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ts.append(float(ms[0]))
        torch.cuda.synchronize()
    if rank == 0:
        ok = bool((total.cpu().numpy().astype(np.uint64) == closed_form(a.n, a.dim, a.q)).all())
        best = min(ts)
        print(json.dumps({"workload": "row-sharded uint32 linear scan (cfg5)", "n": a.n, "dim": a.dim, "queries": a.q, "n_gpus": world,
                          "ms_per_scan": best, "all_ms": [round(x, 3) for x in ts], "u32_mac_per_s": a.n * a.dim * a.q / best * 1e3,
                          "checksums_equal_closed_form": ok, "exchange": "NCCL all-reduce of int64 partial checksums, reduced mod 2^32"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
