"""Lock-step private search with the frontier on the GPU: one group of L clients, R rounds; queries/s and (under ncu) the
launch list of a round.  MS-MARCO shape by default."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="msmarco", choices=["sift", "msmarco"])
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--lanes", type=int, default=32)
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("--step", type=int, default=20)
a = ap.parse_args()
n = a.n or (1000000 if a.shape == "sift" else 3201821)
dim, m, k = (128, 32, 10) if a.shape == "sift" else (192, 32, 100)
rng = np.random.default_rng(1)
vec = (rng.integers(0, 256, (n, dim)).astype(np.float32) if a.shape == "sift"
       else rng.standard_normal((n, dim), dtype=np.float32) * np.linspace(0.82, 0.29, dim, dtype=np.float32))
graph = rng.integers(0, n, (n, m), dtype=np.int32)
from pacmann_b200 import cabi, graphann
group = graphann.make_client_group(vec, graph, a.lanes, seeds=[11 + i for i in range(a.lanes)])
qs = vec[np.random.default_rng(2).integers(0, n, a.lanes * (a.rounds + 1))] + np.float32(0.25)
graphann.SearchKNNLockstep(group, qs[:a.lanes], k, a.step, 3)
l0 = cabi.launch_count()
t0 = time.perf_counter()
graphann.SearchKNNLockstep(group, qs[a.lanes:], k, a.step, 3)
dt = time.perf_counter() - t0
nq = a.lanes * a.rounds
print(f"{a.shape}: {a.lanes} lanes x {a.rounds} rounds: {nq / dt:.0f} queries/s, {dt / (a.rounds * a.step) * 1e6:.0f} us per step, "
      f"{cabi.launch_count() - l0} launches, device stats {graphann.DeviceSearchStats()}")
