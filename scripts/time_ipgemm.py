"""Time the tensor-core linear scan (cfg5: 3 201 821 x 192 uint32, Q = 1000) on a dedicated stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pacmann_b200 import cabi
n, d, nq = 3201821, 192, 1000
torch.cuda.init()
rows = torch.randint(0, 2**31, (n, d), dtype=torch.int32, device="cuda")
db = cabi.DB(n_rows=n, entry_u64=d // 2, device=0, device_ptr=rows.data_ptr())
qs = torch.randint(0, 2**31, (nq, d), dtype=torch.int32, device="cuda")
cs = torch.empty(nq, dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.synchronize()
with torch.cuda.stream(stream):
    ts = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        cabi.check(cabi.lib().pm_ip_u32_scan_dev(db.h, d, qs.data_ptr(), nq, cs.data_ptr(), None, stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
macs = n * 1024 * d * 10  # ten u8 x u8 limb pairs per u32 product, queries padded to 1024
print(f"ip gemm: {min(ts):.3f} ms (all {[round(t, 3) for t in ts]})  int8 MAC/s {macs / min(ts) / 1e9:.1f} T  u32-equivalent {n * d * nq / min(ts) / 1e9:.1f} T MAC/s")
