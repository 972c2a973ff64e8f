"""K7 (world = 1, no NVLink) vs nrx_adamw_dense_dev on the bench's flat buffer size; reps captured in one graph."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from news_recsys_b200 import _lib as L
from news_recsys_b200.parallel import PeerBuffer
DEV = torch.device("cuda", 0)
lib = L.load()
n = 2_606_000 // 4 * 4
hp = torch.tensor([1e-3, 0.1, 0.0316, 0.0], device=DEV)
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best
p, g, m, v = (torch.randn(n, device=DEV) * 0.01 for _ in range(4)); v.abs_()
t = timeit(lambda: L.check(lib.nrx_adamw_dense_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, hp.data_ptr(), 0.9, 0.999, 1e-8, 0.01, L.stream_ptr(DEV)), "adamw"))
print(f"nrx_adamw_dense_dev (torch memory)        {t:7.2f} us")
pb, gb, sb = PeerBuffer(4 * n, DEV), PeerBuffer(4 * n, DEV), PeerBuffer(1024, DEV, "<i4")
pp, gg = pb.tensor(), gb.tensor(); pp.copy_(p); gg.copy_(g)
t = timeit(lambda: L.check(lib.nrx_adamw_dense_dev(pp.data_ptr(), gg.data_ptr(), m.data_ptr(), v.data_ptr(), n, hp.data_ptr(), 0.9, 0.999, 1e-8, 0.01, L.stream_ptr(DEV)), "adamw"))
print(f"nrx_adamw_dense_dev (nrx_peer_alloc memory) {t:7.2f} us")
st = L.NrxPeerStep(); st.rank, st.world = 0, 1
st.p[0], st.g[0], st.sig[0] = pb.ptr, gb.ptr, sb.ptr
st.m, st.v, st.n, st.d_hparams = m.data_ptr(), v.data_ptr(), n, hp.data_ptr()
st.beta1, st.beta2, st.eps, st.weight_decay = 0.9, 0.999, 1e-8, 0.01
for cap in ("", "148", "296", "592", "1184"):
    if cap: os.environ["NRX_K7_CAP"] = cap
    t = timeit(lambda: L.check(lib.nrx_adamw_allreduce_peer(C.byref(st), L.stream_ptr(DEV)), "k7"))
    print(f"K7 world=1 cap={cap or 'default'}  {t:7.2f} us")
