"""Two HBM-bound kernels for an ncu capture: batched server Answer over 640-byte rows, and the one-query uint32 scan."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pacmann_b200 import cabi

torch.cuda.init()
stream = torch.cuda.Stream()
st = stream.cuda_stream
n, E, q, c, s = 1000000, 80, 96000, 512, 124
buf = torch.randint(-2**62, 2**62, (n * E,), dtype=torch.int64, device="cuda")
db = cabi.DB(n_rows=n, entry_u64=E, device=0, device_ptr=buf.data_ptr())
part = torch.randint(0, 16, (q,), device="cuda")
row0 = (part * 62500).to(torch.int64)
nrows = torch.full_like(row0, 62500)
chunk = torch.full((q,), c, dtype=torch.int32, device="cuda")
sets = torch.full((q,), s, dtype=torch.int32, device="cuda")
offs = torch.randint(0, c, (q, s), dtype=torch.int32, device="cuda")
out = torch.empty(q * E, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    cabi.check(cabi.lib().pm_answer_batch_dev(db.h, row0.data_ptr(), nrows.data_ptr(), chunk.data_ptr(), sets.data_ptr(), offs.data_ptr(), s, q, out.data_ptr(), st))
torch.cuda.synchronize()
n2, d = 3201821, 192
rows = torch.randint(0, 2**31, (n2, d), dtype=torch.int32, device="cuda")
db2 = cabi.DB(n_rows=n2, entry_u64=d // 2, device=0, device_ptr=rows.data_ptr())
qs = torch.randint(0, 2**31, (1, d), dtype=torch.int32, device="cuda")
cs = torch.empty(1, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    cabi.check(cabi.lib().pm_ip_u32_scan_dev(db2.h, d, qs.data_ptr(), 1, cs.data_ptr(), None, st))
torch.cuda.synchronize()
print("ok")
