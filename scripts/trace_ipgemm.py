"""Debug helper: run ONE tensor-core scan with a library built with `make EXTRA=-DPM_IPGEMM_TRACE`; the per-chunk clock
stamps of CTA 0 (TMA issue, full seen, TMEM store, sorted arrive, MMA sees sorted, MMA issued) go to stderr as CSV."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pacmann_b200 import cabi
n, d, nq = 3201821, 192, 1000
torch.cuda.init()
rows = torch.randint(0, 2**31, (n, d), dtype=torch.int32, device="cuda")
db = cabi.DB(n_rows=n, entry_u64=d // 2, device=0, device_ptr=rows.data_ptr())
qs = torch.randint(0, 2**31, (nq, d), dtype=torch.int32, device="cuda")
cs = torch.empty(nq, dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.synchronize()
cabi.check(cabi.lib().pm_ip_u32_scan_dev(db.h, d, qs.data_ptr(), nq, cs.data_ptr(), None, stream.cuda_stream))
torch.cuda.synchronize()
