"""Where does a private-search step spend its time?  (MS-MARCO shape, resident client)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pacmann_b200 import cabi, pianopir, graphann

n, E, dim, m = 3201821, 112, 192, 32
rng = np.random.default_rng(1)
raw = rng.integers(0, 2**64, size=(n, E), dtype=np.uint64)
PIR = pianopir.NewSimpleBatchPianoPIR(n, E * 8, 32, raw, 8)
PIR.SetSeeds(1, 2)
PIR.EnableResidentClient()
t0 = time.perf_counter(); PIR.Preprocessing(); print("preprocessing call %.2f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter(); PIR.Preprocessing(); print("preprocessing call (2nd) %.2f ms, internal %.2f ms" % ((time.perf_counter() - t0) * 1e3, PIR.PreprocessingTime() * 1e3))
for it in range(3):
    batches = [rng.integers(0, n, 96).astype(np.uint64) for _ in range(200)]
    l0 = cabi.launch_count()
    t0 = time.perf_counter()
    for b in batches:
        PIR.Query(b)
    dt = (time.perf_counter() - t0) / len(batches)
    print("PIR.Query(96 ids): %.1f us per call, %.1f launches" % (dt * 1e6, (cabi.launch_count() - l0) / len(batches)))
vecs = rng.standard_normal((96, dim)).astype(np.float32)
q = rng.standard_normal(dim).astype(np.float32)
for it in range(2):
    t0 = time.perf_counter()
    for _ in range(500):
        cabi.l2_query(vecs, q)
    print("pm_l2_query(96 x 192): %.1f us per call" % ((time.perf_counter() - t0) / 500 * 1e6))
big = rng.standard_normal((2000, dim)).astype(np.float32)
t0 = time.perf_counter()
for _ in range(200):
    cabi.l2_query(big, q)
print("pm_l2_query(2000 x 192): %.1f us per call" % ((time.perf_counter() - t0) / 200 * 1e6))
