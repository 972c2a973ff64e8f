"""One top-k search at BASELINE cfg4 size for profiler runs (ncu launch lists / full captures of the scan kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from news_recsys_b200.retrieval import TopkIndex  # noqa: E402

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N, D = 1_000_000, 128
c = torch.nn.functional.normalize(torch.randn(N, D, device="cuda"), dim=1)
idx = TopkIndex(c)
q = torch.nn.functional.normalize(torch.randn(Q, D, device="cuda"), dim=1)
for _ in range(2):
    s, i, st = idx.search(q, 100, want_status=True)
torch.cuda.synchronize()
print("fallback queries:", int(st.sum()))
