#!/usr/bin/env python
"""bench.py -- headline benchmark: PianoPIR hint-generation DB-scan GB/s on the MS-MARCO-shaped batch-PIR
workload (BASELINE.json configs[3]: 3 201 821 entries x 896 B, batch 32 -> 16 sub-PIRs, FailureProbLog2 8).

A "step" is one full SimpleBatchPianoPIR.Preprocessing() (pianopir/batch-pir.go:119-155): every primary and
backup hint parity of all 16 sub-PIRs over the whole DB.  The DB is replicated per GPU and hints are sharded
by hint set across ranks; with N > 1 the parities are gathered on rank 0 over NCCL inside the
timed region, so the job is the same at every N ("strong" scaling: total work fixed).

  value  = N*EB / t            DB-scan GB/s, DB resident in HBM, parities left in HBM (rank 0 after the gather)
  e2e    = same metric through the host-buffer C-ABI call pm_hintgen(): job descriptors in, parities copied
           back into pinned host memory inside the timed region (what the cgo bridge would do per Preprocessing)
  roofline: algorithmic HBM bytes B_hbm = N*EB + sum_parts (P+B)*EB (DB read once + parities written once,
           SURVEY.md 8d) / kernel time, against MEASURED_PEAKS.json hbm_gbs.  The kernel is NOT HBM-bound
           (see DESIGN.md): xor_gather GB/s and PRF/s are reported next to it.
  cpu_baseline: the C oracle (a port: the reference is Go and cannot be built here) on 1 thread, as the
           reference runs (ThreadNum = 1), on a bounded sample (4 of the 16 sub-PIRs).
  --impl reference: the same oracle with all host threads on the full workload (rank 0 only).

Only the cpu_baseline / --impl reference legs import oracle/; the timed GPU path is the C-ABI library.
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

N_ROWS = 3201821
ENTRY_U64 = 112
BATCH = 32
FAIL_LOG2 = 8
SEED = 20241600
NCU_DRAM_BYTES_PER_LAUNCH = 9896459000 + 295330304   # profiles/r01_hintgen_msmarco_v3_ncu_full.csv


def pir_params(n, fail_log2):
    """NewPianoPIR / NewPianoPIRClient parameter derivation (pianopir/pir.go:487-494, 138-142)."""
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


def partitions(n, batch):
    parts = batch // 2
    ps = (n + parts - 1) // parts
    out = []
    for i in range(parts):
        n_i = min(ps, n - i * ps)
        c, s, p, mq = pir_params(n_i, FAIL_LOG2)
        out.append(dict(row0=i * ps, n_rows=n_i, chunk=c, set=s, primary=p, mqpc=mq, hints=p + s * mq))
    return out


def gen_db(n_rows, entry_u64, row0=0, rows=None):
    """Synthetic rawDB rows [row0, row0+rows): rng.Uint64() per word as TestBatchPIRPerf (pir_test.go:218-223),
    from a seeded, block-indexed generator so any slice can be produced independently."""
    rows = n_rows - row0 if rows is None else rows
    out = np.empty((rows, entry_u64), np.uint64)
    blk = 1 << 16
    r = row0
    while r < row0 + rows:
        b = r // blk
        lo, hi = b * blk, min((b + 1) * blk, n_rows)
        data = np.random.Generator(np.random.PCG64(SEED + b)).integers(0, 2**64, size=(hi - lo, entry_u64), dtype=np.uint64)
        a, z = max(r, lo), min(row0 + rows, hi)
        out[a - row0:z - row0] = data[a - lo:z - lo]
        r = z
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_reference(args, rank):
    """--impl reference: the CPU oracle port with all host threads on the full workload; rank 0 only."""
    if rank != 0:
        return
    from oracle import oracle as o
    o.lib()
    threads = os.cpu_count() or 1
    parts = partitions(N_ROWS, BATCH)
    db = gen_db(N_ROWS, ENTRY_U64)
    pir = o.SimpleBatchPianoPIR(N_ROWS, ENTRY_U64 * 8, BATCH, db.reshape(-1), FAIL_LOG2)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        pir.preprocessing(key_seed=SEED + it, repl_seed=it, threads=threads)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    val = N_ROWS * ENTRY_U64 * 8 / t / 1e9
    line = {
        "impl": "reference", "metric": "pir_hintgen_db_scan_gbs", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(parts, 1),
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": "full workload: all 16 sub-PIRs, OpenMP over sub-PIRs then hint ranges"},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Go + Plan-9 asm and cannot be built here (no Go toolchain): this is the C oracle port "
                "(same AES-NI / AVX2 instructions), an upper bound on the Go code's speed",
    }
    print(json.dumps(line), flush=True)


def workload_config(parts, n_gpus, exchange="none"):
    return {
        "workload": "MS-MARCO-shaped batch-PIR hint preprocessing (BASELINE.json configs[3])",
        "n_entries": N_ROWS, "entry_bytes": ENTRY_U64 * 8, "batch_size": BATCH, "sub_pirs": len(parts),
        "fail_prob_log2": FAIL_LOG2, "chunk_size": parts[0]["chunk"], "set_size": parts[0]["set"],
        "primary_hints": parts[0]["primary"], "backup_hints": parts[0]["set"] * parts[0]["mqpc"],
        "sharding": f"hint-set x{n_gpus}, DB replicated per GPU", "exchange": exchange,
        "l2_policy": "inputs_exceed_l2 (2.87 GB table vs 126 MB L2)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N > 1: how parities reach rank 0")
    ap.add_argument("--no-search", action="store_true", help="skip the private-ANN queries/s part")
    ap.add_argument("--search-queries", type=int, default=96, help="private ANN queries per GPU")
    ap.add_argument("--search-clients", type=int, default=0, help="serving measurement, thread form: concurrent clients per GPU, one host thread each (0 = skip)")
    ap.add_argument("--search-lanes", type=int, default=16, help="serving measurement, lock-step form: clients per lock-step group (graphann.SearchKNNLockstep); 0 = skip")
    ap.add_argument("--search-groups", type=int, default=4, help="lock-step groups per GPU, one host thread each")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from pacmann_b200 import cabi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    cabi.lib()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    parts = partitions(N_ROWS, BATCH)
    E = ENTRY_U64
    db_bytes = N_ROWS * E * 8
    total_hints = sum(p["hints"] for p in parts)
    n_prf = sum(p["set"] * (p["hints"] - p["mqpc"]) for p in parts)     # S*H' : backup group c skips chunk c
    b_hbm = db_bytes + total_hints * E * 8
    b_xor = n_prf * E * 8

    # ---- inputs: synthetic DB generated on the host, uploaded once (replicated per GPU) ----
    host_db = gen_db(N_ROWS, E)
    db = cabi.DB(host_db, device=local_rank)
    rk_all = []
    from pacmann_b200.keys import derive_key  # product-side key derivation (no oracle import here)
    for i in range(len(parts)):
        rk_all.append(cabi.expand_key(derive_key(SEED, 0, len(parts), i)))

    # ---- hint-set sharding: rank r owns hints [H*r/N, H*(r+1)/N) of every sub-PIR ----
    def shard(h):
        return h * rank // world, h * (rank + 1) // world

    my_hints = [shard(p["hints"]) for p in parts]
    my_count = sum(b - a for a, b in my_hints)
    # Job groups: with N > 1 the NCCL gather of one group's parities runs on a second stream while the next group's
    # kernel computes, so the exchange hides behind the math instead of following it.
    # (each group is one launch of whole 148-CTA rounds: at N = 8 a rank has ~5 rounds of work in total, so 2 groups)
    n_groups = 1 if world == 1 else min(4 if world <= 4 else 2, len(parts))
    bounds = [len(parts) * g // n_groups for g in range(n_groups + 1)]

    def group_count(r, g):
        return sum(parts[i]["hints"] * (r + 1) // world - parts[i]["hints"] * r // world for i in range(bounds[g], bounds[g + 1]))

    grp_pad = [max(group_count(r, g) for r in range(world)) for g in range(n_groups)]   # equal-size gather buffers
    max_count = sum(grp_pad)
    out_grp = [torch.zeros(grp_pad[g] * E, dtype=torch.int64, device=dev) for g in range(n_groups)]
    gather_grp = [[torch.empty_like(out_grp[g]) for _ in range(world)] if (world > 1 and rank == 0) else None
                  for g in range(n_groups)]

    def make_jobs(bases):
        """one pm_hint_job per sub-PIR; group g's parities are packed from bases[g]"""
        jobs = []
        for g in range(n_groups):
            off = 0
            for i in range(bounds[g], bounds[g + 1]):
                p, (a, b) = parts[i], my_hints[i]
                jobs.append(cabi.make_job(p["row0"], p["n_rows"], p["chunk"], p["set"], rk_all[i], a, b - a, p["primary"], p["mqpc"],
                                          parity_out=bases[g] + off * E * 8))
                off += b - a
        return jobs

    jobs_dev = make_jobs([t.data_ptr() for t in out_grp])
    stream = torch.cuda.Stream(device=dev)
    comm = torch.cuda.Stream(device=dev)

    # ---- N > 1, default exchange: no gather step at all.  Rank 0 owns the full parity table ([hints][E] per sub-PIR,
    # in hint order); every other rank maps it over CUDA IPC and its hint kernel stores its shard straight into rank 0's
    # HBM through NVLink peer memory while it computes.  One 1-element all-reduce per step is the completion signal.
    p2p_table, p2p_local, jobs_p2p, flag = None, None, None, None
    part_off = np.concatenate([[0], np.cumsum([p["hints"] for p in parts])]).astype(np.int64)
    if world > 1 and args.exchange == "p2p":
        try:
            handle = [None]
            if rank == 0:
                p2p_local = cabi.buf_alloc(int(part_off[-1]) * E * 8, local_rank)
                handle[0] = cabi.buf_ipc_export(p2p_local, local_rank)
            dist.broadcast_object_list(handle, src=0)
            p2p_table = p2p_local if rank == 0 else cabi.buf_ipc_open(handle[0], local_rank)
            jobs_p2p = [cabi.make_job(p["row0"], p["n_rows"], p["chunk"], p["set"], rk_all[i], a, b - a, p["primary"], p["mqpc"],
                                      parity_out=p2p_table + (int(part_off[i]) + a) * E * 8)
                        for i, (p, (a, b)) in enumerate(zip(parts, my_hints))]
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
        except Exception as exc:       # e.g. IPC not permitted in this container: fall back to the NCCL gather
            if rank == 0:
                print(f"bench.py: peer-memory exchange unavailable ({exc}); using NCCL gather", file=sys.stderr)
            jobs_p2p = None
        ok = torch.tensor([1 if jobs_p2p is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok[0]) == 0:
            jobs_p2p = None
    exchange = "single GPU" if world == 1 else ("peer-memory stores into rank 0 (CUDA IPC over NVLink) + 1-element all-reduce"
                                                if jobs_p2p is not None else "NCCL gather overlapped by job group")

    def step_device():
        if world == 1:
            cabi.hintgen_dev(db, jobs_dev, stream.cuda_stream)
            return
        if jobs_p2p is not None:
            cabi.hintgen_dev(db, jobs_p2p, stream.cuda_stream)
            dist.all_reduce(flag)      # stream-ordered after the kernel on every rank: all shards have landed when it returns
            return
        for g in range(n_groups):
            cabi.hintgen_dev(db, jobs_dev[bounds[g]:bounds[g + 1]], stream.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                dist.gather(out_grp[g], gather_grp[g], dst=0)
        stream.wait_stream(comm)      # the next step may overwrite the buffers only after the gathers have read them

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    torch.cuda.synchronize()   # buffers above were created on torch's default stream
    sampler = ClockSampler(local_rank)
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
        barrier()
        l0 = cabi.launch_count()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ek0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ek1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        e0.record(stream)
        for i in range(args.steps):
            ek0[i].record(stream)
            step_device()
            ek1[i].record(stream)
        e1.record(stream)
        barrier()
        clocks = sampler.result()
        launches = cabi.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    # duration of the dominant kernel alone (no gather), for the roofline: CUDA events on its own stream
    with torch.cuda.stream(stream):
        ks = []
        for _ in range(max(3, min(args.steps, 10))):
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream)
            cabi.hintgen_dev(db, jobs_dev, stream.cuda_stream)
            k1.record(stream)
            torch.cuda.synchronize()
            ks.append(k0.elapsed_time(k1))
    kern_ms = float(np.mean(ks))
    t = torch.tensor([ms_total, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t[0]) / args.steps
    kern_ms = float(t[1])
    value = db_bytes / (ms_step * 1e-3) / 1e9

    # ---- end-to-end through the host-buffer C-ABI (pinned host outputs, D2H inside the timed region) ----
    out_host = torch.empty(my_count * E, dtype=torch.int64).pin_memory()
    host_bases, acc = [], 0
    for g in range(n_groups):
        host_bases.append(out_host.data_ptr() + acc * E * 8)
        acc += group_count(rank, g)
    jobs_host = make_jobs(host_bases)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        cabi.hintgen(db, jobs_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        cabi.hintgen(db, jobs_host)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    h2d_bytes = len(parts) * 256 * world
    d2h_bytes = total_hints * E * 8

    # ---- spot check (outside every timed region): a few parities recomputed from the PRF definition ----
    verified = None
    if rank == 0:
        verified = spot_check(host_db, parts, rk_all, my_hints, out_host.numpy().view(np.uint64).reshape(-1, E))
        if jobs_p2p is not None:       # the table every rank wrote into: check hints of every rank's shard
            full = torch.empty(int(part_off[-1]) * E, dtype=torch.int64)
            cudart = C.CDLL("libcudart.so.12")
            assert cudart.cudaMemcpy(C.c_void_p(full.data_ptr()), C.c_void_p(p2p_table), C.c_size_t(full.numel() * 8), 2) == 0
            all_hints = [(0, p["hints"]) for p in parts]
            verified = verified and spot_check(host_db, parts, rk_all, all_hints, full.numpy().view(np.uint64).reshape(-1, E),
                                               extra=[p["hints"] * r // world for p in parts[:1] for r in range(1, world)])

    # ---- second half of the metric: end-to-end private-ANN queries/s on MS-MARCO-shaped data ----
    private_ann = None
    if not args.no_search:
        del host_db, out_host, jobs_host, jobs_dev, out_grp, gather_grp
        if world > 1:
            dist.barrier()
        if p2p_table is not None and rank != 0:
            cabi.buf_ipc_close(p2p_table, local_rank)
        if world > 1:
            dist.barrier()
        if p2p_local is not None:
            cabi.buf_free(p2p_local, local_rank)
        db.close()
        torch.cuda.empty_cache()
        private_ann = private_search(args, rank, world, local_rank, dist if world > 1 else None, dev)

    if rank == 0:
        peak, peak_src = measured_peak()
        b_hbm_rank = db_bytes + max_count * E * 8      # per GPU: whole DB read once + its share of the parities
        achieved = b_hbm_rank / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": "pir_hintgen_db_scan_gbs", "value": value, "unit": "GB/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_config(parts, world, exchange),
            "clocks": clocks, "gpu_launches": int(launches) * world,
            "e2e": {"value": db_bytes / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_s * 1e3,
                    "timer": "host perf_counter around the synchronous pm_hintgen() call, max over ranks",
                    "note": "DB is uploaded once at pm_db_create (as rawDB is built once in NewSimpleBatchPianoPIR); "
                            "per step only job descriptors go in and all parities come back to pinned host memory"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH if world == 1 else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel at N=1, one ncu --set full "
                                           "capture (profiles/r01_hintgen_msmarco_v3_ncu_full.csv)",
                         "peak_source": peak_src, "kernel": "hintgen_kernel<uint4,8,7,...>",
                         "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": b_hbm_rank,
                         "binding_term": "not HBM: L1 data-pipe wavefronts (row gather + AES T-table LDS) and ALU; see DESIGN.md",
                         "xor_gather_gbs": b_xor / world / (kern_ms * 1e-3) / 1e9,
                         "prf_per_s": n_prf / world / (kern_ms * 1e-3)},
            "verified_vs_oracle_prf": verified,
            "private_ann": private_ann,
            "reference_published": {"msmarco_prep_s": "9-10 (1 thread, reproduction/msmarco/README.md:26)"},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(parts)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def private_search(args, rank, world, local_rank, dist, dev):
    """End-to-end private graph search (graphann.SearchKNN over PIRGraphInfo, private-search.go) on MS-MARCO-shaped
    synthetic data: 3 201 821 x 192 fp32, degree-32 random graph (genRandomGraph, private-search.go:54-69), k = 100,
    step 20, parallel 3 (reproduction/msmarco/reproduce.sh:226-230).  Query batches shard across GPUs: every rank
    owns a replica of the DB and an independent client and answers its own queries.  The clients'
    periodic re-preprocessing ("maintenance", which the reference reports separately, private-search.go:219-240) is inside
    the timed region (4.9 ms every ~45 queries per client); the initial Preprocessing is reported on its own.  Rank 0 also times the CPU oracle on a few queries."""
    import torch
    from pacmann_b200 import cabi, graphann
    from pacmann_b200.keys import mix64
    n, dim, m, k, step, par = N_ROWS, 192, 32, 100, 20, 3
    rng = np.random.default_rng(SEED)
    vec = rng.standard_normal((n, dim), dtype=np.float32) * np.linspace(0.82, 0.29, dim, dtype=np.float32)
    graph = rng.integers(0, n, (n, m), dtype=np.int32)
    loop = graph == np.arange(n, dtype=np.int32)[:, None]
    graph[loop] = (graph[loop] + 1) % n
    nq = args.search_queries
    qall = vec[np.random.default_rng(SEED + 1).integers(0, n, nq * world)] + np.float32(0.25)
    queries = qall[rank * nq:(rank + 1) * nq]
    seed = SEED + 2
    # torch's own CPU thread pool is not needed here and its idle workers compete with the search driver threads for the
    # cores (measured: 2 880 vs 4 620 lock-step queries/s on a 16-core host)
    torch.set_num_threads(1)
    lanes, ngroups = max(0, args.search_lanes), max(1, args.search_groups)
    # host threads: the ranks of a node share its cores; a lock-step group = one driver thread + an OpenMP team for the
    # per-lane host work.  Oversubscribing the cores with spinning teams is ruinous, so both are sized to the share.
    cores_per_rank = max(1, (os.cpu_count() or 1) // max(1, world))
    ngroups = max(1, min(ngroups, cores_per_rank))
    os.environ["PM_HOST_THREADS"] = str(max(1, min(8, cores_per_rank // ngroups)))
    f = graphann.GraphANNFrontend(vec, graph, private=True, seed=seed, device=local_rank, group_lanes=max(1, lanes))
    t0 = time.perf_counter()
    f.Preprocess()
    setup_s = time.perf_counter() - t0
    pir = f.PIR
    prep_s = pir.PreprocessingTime()
    # warm-up on queries of its own: repeating a measured query would be answered from the client's local cache
    wq = vec[np.random.default_rng(SEED + 7).integers(0, n, 2)] + np.float32(0.25)
    f.SearchKNNBatch(wq, k, step, par)
    if dist is not None:
        dist.barrier()
    l0, s0 = cabi.launch_count(), pir.serverQueries
    t0 = time.perf_counter()
    f.SearchKNNBatch(queries, k, step, par)
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt_max = float(tt[0])
    out = {
        "workload": "MS-MARCO-shaped synthetic private graph search (BASELINE.json configs[2])", "n": n, "dim": dim, "m": m, "k": k,
        "step": step, "parallel": par, "queries": nq * world, "queries_per_gpu": nq, "n_gpus": world,
        "queries_per_s": nq * world / dt_max, "s_per_query_one_client": dt / nq,
        "pir_preprocessing_s": prep_s, "setup_s_pack_upload_preprocess": setup_s,
        "gpu_launches": int(cabi.launch_count() - l0), "server_subqueries": int(pir.serverQueries - s0),
        "pir_success_rate": f.succQueryNum / max(1, f.totalQueryNum),
        "client": "GPU-resident hint tables (pm_client_*); the timed region INCLUDES the client's re-preprocessing whenever its "
                  "query budget runs out (every ~45 queries: 4.9 ms on the GPU) -- the reference reports that maintenance "
                  "separately (private-search.go:219-240), here it is 3 % of the time",
        "pir_preprocessing_note": "wall clock of SimpleBatchPianoPIR.Preprocessing() with the resident client: key schedules, table "
                                  "init, offset index, hint kernel, replacement gather; nothing returns to the host",
        "reference_published": "0.0559 / 0.0640 s per query on SIFT1M, 1 thread (private-search-report.txt:19,44)",
    }
    # serving form: several clients (one per user: own keys, hint tables, search state) over the ONE resident DB replica,
    # driven from host threads; every client still answers its queries sequentially, exactly as a single one does
    K = args.search_clients
    if K > 1:
        import threading
        fs = [f] + [graphann.GraphANNFrontend(vec, graph, seed=seed + 100 + i, share_db_with=f) for i in range(K - 1)]
        for g in fs[1:]:
            g.Preprocess()
            g.SearchKNNBatch(wq[:1], k, step, par)
        per = max(8, nq // 2)
        qs = [vec[np.random.default_rng(SEED + 10 + rank * K + i).integers(0, n, per)] + np.float32(0.25) for i in range(K)]

        def work(i):
            fs[i].SearchKNNBatch(qs[i], k, step, par)

        th = [threading.Thread(target=work, args=(i,)) for i in range(K)]
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        mdt = time.perf_counter() - t0
        tm = torch.tensor([mdt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        out["multi_client"] = {"clients_per_gpu": K, "queries": K * per * world, "queries_per_s": K * per * world / float(tm[0]),
                               "s_per_query_per_client": mdt / per}
        del fs
    # serving form, lock step: groups of `lanes` independent clients whose hint tables live in one
    # pm_client; every search step of a group is ONE device call for all its lanes.  Each client still answers its own
    # queries sequentially and returns exactly what it would return alone (tests/test_graphann_gpu.py).  Several groups
    # run from host threads so that the host work of one overlaps the GPU work of another.
    if lanes > 1:
        import threading
        t0 = time.perf_counter()
        groups = []
        for gi in range(ngroups):
            lead = f if gi == 0 else graphann.GraphANNFrontend(vec, graph, seed=seed + 5000 * gi, share_db_with=f, group_lanes=lanes)
            if gi:
                lead.Preprocess()
            grp = [lead]
            for i in range(1, lanes):
                g = graphann.GraphANNFrontend(vec, graph, seed=seed + 5000 * gi + 1000 + i, lane_of=lead, lane=i)
                g.Preprocess()
                grp.append(g)
            groups.append(grp)
        gsetup = time.perf_counter() - t0
        per = max(4, nq // 16)
        lqs = [vec[np.random.default_rng(SEED + 50 + rank * ngroups + gi).integers(0, n, lanes * per)] + np.float32(0.25) for gi in range(ngroups)]
        # one host thread per group; each warms up in its own thread (one query per lane: OpenMP team, allocator arena)
        # and then waits for the common start
        start_b, done_b = threading.Barrier(ngroups + 1), threading.Barrier(ngroups + 1)

        wqs = [vec[np.random.default_rng(SEED + 900 + rank * ngroups + gi).integers(0, n, lanes)] + np.float32(0.25) for gi in range(ngroups)]

        def drive(gi):
            graphann.SearchKNNLockstep(groups[gi], wqs[gi], k, step, par)   # warm-up queries of their own (no cache hits later)
            start_b.wait()
            graphann.SearchKNNLockstep(groups[gi], lqs[gi], k, step, par)
            done_b.wait()

        th = [threading.Thread(target=drive, args=(gi,)) for gi in range(ngroups)]
        for t_ in th:
            t_.start()
        if dist is not None:
            dist.barrier()
        start_b.wait()
        l1 = cabi.launch_count()
        t0 = time.perf_counter()
        done_b.wait()
        ldt = time.perf_counter() - t0
        for t_ in th:
            t_.join()
        tl = torch.tensor([ldt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        nlq = ngroups * lanes * per
        out["lockstep"] = {"clients_per_gpu": ngroups * lanes, "groups_per_gpu": ngroups, "lanes_per_group": lanes, "queries": nlq * world,
                           "queries_per_s": nlq * world / float(tl[0]), "ms_per_step_per_group": ldt / (per * step) * 1e3,
                           "gpu_launches": int(cabi.launch_count() - l1), "group_setup_s": gsetup,
                           "host_threads_per_group": int(os.environ["PM_HOST_THREADS"]), "host_cores_per_rank": cores_per_rank,
                           "note": "every client keeps its own keys, hint tables, cache and search state; results identical to each client searching alone"}
        del groups
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as o
        raw = o.pack_db(vec, graph)
        o_pir = o.SimpleBatchPianoPIR(n, (dim + m) * 4, m, raw, 8)
        t0 = time.perf_counter()
        o_pir.preprocessing(key_seed=mix64(seed, 1), repl_seed=mix64(seed, 2), threads=os.cpu_count())
        cprep = time.perf_counter() - t0
        start = f.StartVertexIds()
        ncpu = 40    # ~2 s of single-thread CPU work; more would cross the client's query budget and mix re-preprocessing into the time
        cq = vec[np.random.default_rng(SEED + 3).integers(0, n, ncpu + 1)] + np.float32(0.25)
        o.search_knn_private(o_pir, vec, graph, start, cq[:1], k, step, par)
        t0 = time.perf_counter()
        o_ret, _, _ = o.search_knn_private(o_pir, vec, graph, start, cq[1:], k, step, par)
        cdt = time.perf_counter() - t0
        out["cpu_baseline"] = {"queries_per_s": ncpu / cdt, "s_per_query": cdt / ncpu, "cores": 1, "kind": "port",
                               "sample": f"{ncpu} queries, online part single-threaded as the reference",
                               "pir_preprocessing_s_all_cores": cprep}
    return out


def spot_check(host_db, parts, rk_all, my_hints, got, extra=()):
    """Outside every timed region: a handful of parities recomputed from the ORACLE's PRF (oracle/, CPU) + numpy XOR.
    The full-size oracle comparison of this very call is tests/test_hintgen_variants_gpu.py."""
    from oracle import oracle as o
    ok, off = True, 0
    rng = np.random.default_rng(1)
    for p, (a, b), rk in zip(parts, my_hints, rk_all):
        if b > a:
            for h in {a, b - 1, int(rng.integers(a, b))} | {x for x in extra if a <= x < b} | {x - 1 for x in extra if a < x <= b}:
                cs = np.arange(p["set"], dtype=np.uint64)
                offs = o.prf_batch(rk, np.full(p["set"], h, np.uint64), cs) & np.uint64(p["chunk"] - 1)
                rows = cs * np.uint64(p["chunk"]) + offs
                skip = -1 if h < p["primary"] else (h - p["primary"]) // p["mqpc"]
                sel = rows[(rows < p["n_rows"]) & (cs != skip if skip >= 0 else True)]
                want = np.bitwise_xor.reduce(host_db[p["row0"] + sel.astype(np.int64)], axis=0)
                ok = ok and bool((got[off + h - a] == want).all())
        off += b - a
    return ok


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(parts):
    """1-thread C oracle (as the reference runs: ThreadNum = 1, pianopir/batch-pir.go:16) on the WHOLE workload -- all 16
    sub-PIRs -- three times over: about 10 s of single-thread CPU work."""
    from oracle import oracle as o
    o.lib()
    reps = 3
    rows = sum(p["n_rows"] for p in parts)
    t = 0.0
    for i, p in enumerate(parts):
        sub = gen_db(N_ROWS, ENTRY_U64, p["row0"], p["n_rows"])
        pir = o.PianoPIR(p["n_rows"], ENTRY_U64 * 8, sub.reshape(-1), FAIL_LOG2)
        for _ in range(reps):
            t0 = time.perf_counter()
            pir.preprocessing(o.derive_key(SEED, 0, len(parts), i), repl_seed=i, threads=1)
            t += time.perf_counter() - t0
    return {"value": reps * rows * ENTRY_U64 * 8 / t / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": f"the full workload (16 sub-PIRs, {rows} rows, {rows * ENTRY_U64 * 8 / 1e9:.2f} GB) x {reps} repetitions, "
                      f"{t:.1f} s of single-thread CPU work",
            "seconds": t, "host_cpu": cpu_model(), "host_cores": os.cpu_count()}


if __name__ == "__main__":
    main()
