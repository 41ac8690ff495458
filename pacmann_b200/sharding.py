"""Hint-set sharding of PianoPIR preprocessing across GPUs: every hint's parity depends only on
(key, tag, DB), so rank r of N computes hints [H*r/N, H*(r+1)/N) of every sub-PIR over its own replica of the DB,
with no data-path exchange; the parities are then gathered on the consumer.  The linear-scan baseline (A11) is the one
piece with a real exchange step: rows are sharded, every rank scans its rows and the per-query checksums -- sums mod
2^32 -- are all-reduced.  Pure index arithmetic, shared by bench.py, scripts/scan_multi_gpu.py and the multi-process
tests."""


def shard_range(n_hints, rank, world):
    """half-open hint range owned by `rank`"""
    return n_hints * rank // world, n_hints * (rank + 1) // world


def partition_owner_range(n_parts, rank, world):
    """partition sharding (batch-pir.go:79-85: sub-PIRs own disjoint DB slices): the half-open range of sub-PIRs that `rank`
    owns WHOLE, together with only their rows.  world must divide n_parts."""
    if n_parts % world:
        raise ValueError(f"partition sharding needs world | n_parts ({world} does not divide {n_parts})")
    return n_parts * rank // world, n_parts * (rank + 1) // world


def hints_of(mode, hints_per_part, rank, world, relief=0.0):
    """per sub-PIR, the half-open hint range `rank` computes: mode "hintset" (every rank a slice of every sub-PIR, DB
    replicated) or "partition" (every rank all hints of its own sub-PIRs, DB sharded by rows).
    relief (partition mode): rank 0 is also the consumer -- the other ranks' parities land in its HBM while its kernel
    runs, which slows that kernel -- so the last `relief` of the hints of rank 0's sub-PIRs are computed by the other
    ranks instead, 1/(world-1) each (they then also keep rank 0's rows)."""
    if mode == "partition":
        lo, hi = partition_owner_range(len(hints_per_part), rank, world)
        out = [(0, h) if lo <= i < hi else (0, 0) for i, h in enumerate(hints_per_part)]
        if relief > 0 and world > 1:
            lo0, hi0 = partition_owner_range(len(hints_per_part), 0, world)
            for i in range(lo0, hi0):
                h = hints_per_part[i]
                keep = h - int(h * relief)
                if rank == 0:
                    out[i] = (0, keep)
                else:
                    out[i] = (keep + (h - keep) * (rank - 1) // (world - 1), keep + (h - keep) * rank // (world - 1))
        return out
    return [shard_range(h, rank, world) for h in hints_per_part]


def table_runs(ranges, part_offsets):
    """(table position, packed local position, length) in hints of the maximal contiguous runs of a rank's hints, given its
    per-sub-PIR ranges and the hint offset of every sub-PIR in the one [hints] table: what a rank copies into the
    consumer's table (one run per rank under partition sharding)"""
    runs, local = [], 0
    for i, (a, b) in enumerate(ranges):
        if b <= a:
            continue
        pos = part_offsets[i] + a
        if runs and runs[-1][0] + runs[-1][2] == pos and runs[-1][1] + runs[-1][2] == local:
            runs[-1] = (runs[-1][0], runs[-1][1], runs[-1][2] + b - a)
        else:
            runs.append((pos, local, b - a))
        local += b - a
    return runs


def shard_sizes(hints_per_part, world):
    """per-rank total hint counts over all sub-PIRs"""
    return [sum(shard_range(h, r, world)[1] - shard_range(h, r, world)[0] for h in hints_per_part) for r in range(world)]


def padded_shard_len(hints_per_part, world):
    """every rank's buffer is padded to the largest shard so a fixed-size gather works"""
    return max(shard_sizes(hints_per_part, world))


def assemble(gathered, hints_per_part, world, entry_u64):
    """gathered[r] = rank r's flat buffer (its shard of sub-PIR 0, then of sub-PIR 1, ...; padded).
    Returns one [n_hints][entry_u64] array per sub-PIR in hint-number order."""
    import numpy as np
    out = [np.zeros((h, entry_u64), np.uint64) for h in hints_per_part]
    for r in range(world):
        buf = np.asarray(gathered[r]).reshape(-1, entry_u64)
        off = 0
        for i, h in enumerate(hints_per_part):
            a, b = shard_range(h, r, world)
            out[i][a:b] = buf[off:off + (b - a)]
            off += b - a
    return out


def row_shard(n_rows, rank, world):
    """half-open row range owned by `rank` in the sharded linear scan"""
    return n_rows * rank // world, n_rows * (rank + 1) // world


def allreduce_checksums(dist, checksums_u32):
    """Combine per-rank partial checksums (uint32, wrapping sums over the rank's rows) into the checksums of the whole
    table.  `checksums_u32` is a torch tensor holding values in [0, 2^32) as int64 (torch has no uint32 collectives):
    an int64 sum over at most 2^31 ranks cannot overflow, the result is reduced mod 2^32 afterwards -- exactly the
    wrap-around of the reference's uint32 accumulator (graphann_test.go:268-273)."""
    dist.all_reduce(checksums_u32, op=dist.ReduceOp.SUM)
    checksums_u32 &= 0xFFFFFFFF
    return checksums_u32
