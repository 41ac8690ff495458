"""Shared helpers for the tests: seeded inputs and PianoPIR table bookkeeping."""
import numpy as np


def splitmix_db(n_rows, entry_u64, seed=1):
    """Deterministic pseudo-random table (numpy PCG64; any fixed generator works: oracle and CUDA path
    are fed the same array)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 2**64, size=(n_rows, entry_u64), dtype=np.uint64)


def oracle_parities(pir):
    """All hint parities of an oracle PianoPIR as one [P + S*M][E] array in hint-number order."""
    import numpy as np
    P, E = pir.primary_hint_num, pir.entry_u64
    prim = pir.table("primary_parity").reshape(P, E)
    back = pir.table("backup_parity").reshape(-1, E)
    return np.concatenate([prim, back], axis=0)
