"""Multi-process (gloo, world_size 2 and 3) test of the hint-set sharding used by bench.py --gpus N: each rank
computes only its hint range of every sub-PIR, the shards are gathered on rank 0 and assembled, and the result must
equal the unsharded tables.  The per-shard compute is the CPU oracle here (no GPU in this container); on the GPU box
the same index arithmetic (pacmann_b200/sharding.py) drives pm_hintgen_dev + an NCCL gather."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, result_file):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o
    from pacmann_b200 import sharding
    from util import splitmix_db

    n, E, batch = 9001, 6, 8
    rows = splitmix_db(n, E, seed=99)
    full = o.SimpleBatchPianoPIR(n, E * 8, batch, rows.reshape(-1), 8)
    parts = full.partition_num
    subs = [full.sub(i) for i in range(parts)]
    hints = [s.primary_hint_num + s.set_size * s.max_query_per_chunk for s in subs]
    pad = sharding.padded_shard_len(hints, world)
    mine = np.zeros((pad, E), np.uint64)
    off = 0
    for i, s in enumerate(subs):
        a, b = sharding.shard_range(hints[i], rank, world)
        key = o.derive_key(7, 0, parts, i)
        s.preprocessing_range(key, a, b)
        tab = np.concatenate([s.table("primary_parity"), s.table("backup_parity").reshape(-1, E)])
        mine[off:off + b - a] = tab[a:b]
        off += b - a
    t = torch.from_numpy(mine.view(np.int64))
    gathered = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gathered, dst=0)
    ok = True
    if rank == 0:
        tables = sharding.assemble([g.numpy().view(np.uint64) for g in gathered], hints, world, E)
        full.preprocessing(key_seed=7, repl_seed=0, threads=2)
        for i in range(parts):
            s = full.sub(i)
            want = np.concatenate([s.table("primary_parity"), s.table("backup_parity").reshape(-1, E)])
            ok = ok and bool((tables[i] == want).all())
        open(result_file, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_hint_set_sharding_gathers_to_the_unsharded_tables(world, tmp_path):
    port = 29500 + os.getpid() % 2000 + world
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(world, port, str(result)), nprocs=world, join=True)
    assert result.read_text() == "ok"


def _table_worker(rank, world, port, result_file, mode, shm_name):
    """every rank computes its share (hint-set or partition sharding) and writes it into ONE shared table at the positions
    bench.py uses for its peer-memory / copy-engine exchange (here: a shared-memory file stands in for rank 0's HBM)"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o
    from pacmann_b200 import sharding
    from util import splitmix_db

    n, E, batch = 8003, 4, 8
    rows = splitmix_db(n, E, seed=98)
    full = o.SimpleBatchPianoPIR(n, E * 8, batch, rows.reshape(-1), 8)
    parts = full.partition_num
    subs = [full.sub(i) for i in range(parts)]
    hints = [s.primary_hint_num + s.set_size * s.max_query_per_chunk for s in subs]
    offs = np.concatenate([[0], np.cumsum(hints)]).astype(np.int64)
    table = np.memmap(shm_name, dtype=np.uint64, mode="r+", shape=(int(offs[-1]), E))
    relief = 0.3 if mode == "partition_relief" else 0.0
    ranges = sharding.hints_of("partition" if relief else mode, hints, rank, world, relief)
    local = np.zeros((sum(b - a for a, b in ranges), E), np.uint64)
    pos = 0
    for i, (a, b) in enumerate(ranges):
        if b > a:
            subs[i].preprocessing_range(o.derive_key(7, 0, parts, i), a, b)
            tab = np.concatenate([subs[i].table("primary_parity"), subs[i].table("backup_parity").reshape(-1, E)])
            local[pos:pos + b - a] = tab[a:b]
            pos += b - a
    runs = sharding.table_runs(ranges, offs)
    if mode == "partition":
        assert len(runs) == 1      # whole consecutive sub-PIRs: one copy per rank
    for tpos, lpos, cnt in runs:
        table[tpos:tpos + cnt] = local[lpos:lpos + cnt]
    table.flush()
    dist.barrier()
    if rank == 0:
        full.preprocessing(key_seed=7, repl_seed=0, threads=2)
        ok = True
        for i in range(parts):
            s = full.sub(i)
            want = np.concatenate([s.table("primary_parity"), s.table("backup_parity").reshape(-1, E)])
            ok = ok and bool((np.asarray(table[offs[i]:offs[i + 1]]) == want).all())
        open(result_file, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,world", [("partition", 2), ("partition", 4), ("hintset", 3), ("partition_relief", 4)])
def test_sharded_ranks_fill_one_table(mode, world, tmp_path):
    """bench.py --gpus N: partition sharding (rank g owns sub-PIRs [4g/N ...) whole) and hint-set sharding both write
    disjoint runs of ONE [hints][E] table that together equal the unsharded preprocessing"""
    from oracle import oracle as o
    from util import splitmix_db
    n, E, batch = 8003, 4, 8
    full = o.SimpleBatchPianoPIR(n, E * 8, batch, splitmix_db(n, E, seed=98).reshape(-1), 8)
    hints = sum(full.sub(i).primary_hint_num + full.sub(i).set_size * full.sub(i).max_query_per_chunk for i in range(full.partition_num))
    shm = tmp_path / "table.bin"
    np.zeros((hints, E), np.uint64).tofile(shm)
    result = tmp_path / "result.txt"
    port = 29800 + os.getpid() % 1000 + world
    mp.spawn(_table_worker, args=(world, port, str(result), mode, str(shm)), nprocs=world, join=True)
    assert result.read_text() == "ok"


def test_partition_sharding_needs_a_divisor():
    from pacmann_b200 import sharding
    with pytest.raises(ValueError):
        sharding.partition_owner_range(16, 0, 3)
    owned = [sharding.partition_owner_range(16, r, 8) for r in range(8)]
    assert owned[0] == (0, 2) and owned[-1] == (14, 16) and all(owned[r][1] == owned[r + 1][0] for r in range(7))


def test_shard_ranges_partition_the_hints():
    from pacmann_b200 import sharding
    for h in (0, 1, 7, 24416, 104448):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(h, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == h
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _scan_worker(rank, world, port, result_file):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o
    from pacmann_b200 import sharding

    n, d, nq = 10007, 32, 9
    rng = np.random.default_rng(5)
    rows = rng.integers(0, 2**32, (n, d), dtype=np.uint32)       # every rank draws the same table
    rows[::3] = 0xFFFFFFFF                                        # partial sums far beyond 2^32: the wrap must survive the all-reduce
    qs = rng.integers(0, 2**32, (nq, d), dtype=np.uint32)
    a, b = sharding.row_shard(n, rank, world)
    part = np.asarray(o.ip_scan(rows[a:b], qs), dtype=np.uint32)  # per-shard compute: the CPU oracle here, pm_ip_u32_scan on the GPU box
    t = torch.from_numpy(part.astype(np.int64))
    sharding.allreduce_checksums(dist, t)
    if rank == 0:
        want = np.asarray(o.ip_scan(rows, qs), dtype=np.uint32)
        open(result_file, "w").write("ok" if (t.numpy().astype(np.uint32) == want).all() else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_scan_allreduces_to_the_unsharded_checksums(world, tmp_path):
    """SURVEY 8e, linear scan: rows sharded over ranks, per-rank partial checksums, all-reduce mod 2^32."""
    result = tmp_path / "scan_result.txt"
    port = 29650 + world
    mp.spawn(_scan_worker, args=(world, port, str(result)), nprocs=world, join=True)
    assert result.read_text() == "ok"
