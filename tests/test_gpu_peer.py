"""-m gpu: K7 (gradient all-reduce fused with AdamW over peer memory, include/nrx.h) on ONE GPU — the ranks are
separate buffers driven from separate streams, which exercises the same flag protocol as separate devices."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(world, n, seed=0):
    from news_recsys_b200 import _lib as L
    from news_recsys_b200.parallel import PeerBuffer
    g = torch.Generator().manual_seed(seed)
    p0 = torch.randn(n, generator=g)
    ranks = []
    for r in range(world):
        pb, gb, sb = PeerBuffer(4 * n, DEV), PeerBuffer(4 * n, DEV), PeerBuffer(4 * L.NRX_PEER_SIG_WORDS, DEV, "<i4")
        p, gr = pb.tensor(), gb.tensor()
        p.copy_(p0)
        gr.copy_(torch.randn(n, generator=g) * 0.1)
        ranks.append(dict(p=p, g=gr, sig=sb.tensor(), m=torch.zeros(n, device=DEV), v=torch.zeros(n, device=DEV)))
    hp = torch.tensor([1e-3, 1 - 0.9, (1 - 0.999) ** 0.5, 0.0], device=DEV)
    steps = []
    for r in range(world):
        st = L.NrxPeerStep()
        st.rank, st.world = r, world
        for j in range(world):
            st.p[j], st.g[j], st.sig[j] = ranks[j]["p"].data_ptr(), ranks[j]["g"].data_ptr(), ranks[j]["sig"].data_ptr()
        st.m, st.v, st.n = ranks[r]["m"].data_ptr(), ranks[r]["v"].data_ptr(), n
        st.d_hparams = hp.data_ptr()
        st.beta1, st.beta2, st.eps, st.weight_decay = 0.9, 0.999, 1e-8, 0.01
        steps.append(st)
    return L, p0, ranks, hp, steps


def _reference(L, p0, gs, hp):
    """nrx_adamw_dense_dev on the rank-ordered mean of the gradients."""
    lib = L.load()
    world = len(gs)
    g = torch.zeros_like(gs[0])
    for x in gs:
        g = g + x
    g = g * torch.tensor(1.0 / world, dtype=torch.float32, device=DEV)
    p = p0.to(DEV).clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    L.check(lib.nrx_adamw_dense_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), hp.data_ptr(),
                                    0.9, 0.999, 1e-8, 0.01, L.stream_ptr(torch.device(DEV))), "nrx_adamw_dense_dev")
    return p, m, v


@pytest.mark.parametrize("world,n", [(1, 4096), (2, 1 << 16), (3, 40004), (8, 1 << 17), (4, 8), (16, 1 << 16)])
def test_peer_step_equals_allreduce_then_adamw(world, n):
    from news_recsys_b200.parallel import shard_range
    L, p0, ranks, hp, steps = _setup(world, n)
    lib = L.load()
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for rep in range(2):   # second launch: flags carry the launch counter, moments carry over
        if rep == 0:
            ref_p, ref_m, ref_v = _reference(L, p0, [r["g"] for r in ranks], hp)
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                L.check(lib.nrx_adamw_allreduce_peer(C.byref(steps[r]), L.stream_ptr(torch.device(DEV))), "nrx_adamw_allreduce_peer")
        torch.cuda.synchronize()
        for r in range(world):
            assert int(ranks[r]["sig"][L_ERR].item()) == 0, "K7 timed out"
        if rep == 0:
            for r in range(world):
                assert torch.equal(ranks[r]["p"], ref_p), f"rank {r}: parameters differ from all-reduce + AdamW"
                lo, hi = shard_range(n // 4, r, world)
                assert torch.equal(ranks[r]["m"][4 * lo:4 * hi], ref_m[4 * lo:4 * hi])
                assert torch.equal(ranks[r]["v"][4 * lo:4 * hi], ref_v[4 * lo:4 * hi])
                rest = torch.ones(n, dtype=torch.bool, device=DEV)
                rest[4 * lo:4 * hi] = False
                assert float(ranks[r]["m"][rest].abs().sum()) == 0.0   # moments live on the owner only
        else:
            for r in range(1, world):
                assert torch.equal(ranks[r]["p"], ranks[0]["p"])
            assert not torch.equal(ranks[0]["p"], ref_p)


L_ERR = 130  # NRX_PEER_SIG_ERR


def test_missing_peer_is_fatal_and_sticky():
    """A peer that never arrives: the kernel gives up after the limit (no hang), leaves the parameters untouched, raises
    the error word on EVERY rank and bit 1 of the host-visible status word; every later launch — on the late rank too —
    returns at entry, so no rank can pair stale flags with new gradients (peer.cu)."""
    L, p0, ranks, hp, steps = _setup(2, 4096)
    lib = L.load()
    status = torch.zeros(2, dtype=torch.int32, device=DEV)
    for r in range(2):
        steps[r].timeout_ms = 300
        steps[r].status = status[r:r + 1].data_ptr()
    before = [ranks[r]["p"].clone() for r in range(2)]
    sp = L.stream_ptr(torch.device(DEV))
    L.check(lib.nrx_adamw_allreduce_peer(C.byref(steps[0]), sp), "nrx_adamw_allreduce_peer")   # rank 1 never launches
    flag = C.c_int32(0)
    L.check(lib.nrx_peer_status(ranks[0]["sig"].data_ptr(), C.byref(flag), sp), "nrx_peer_status")
    assert flag.value == 1
    assert int(status[0].item()) & 2
    assert int(ranks[1]["sig"][L_ERR].item()) == 1, "the error must be raised on the peer as well"
    assert torch.equal(ranks[0]["p"], before[0])
    # the late rank arrives now: it must NOT pass barrier A against rank 0's stale flag and update anything
    epoch1 = int(ranks[1]["sig"][128].item())
    L.check(lib.nrx_adamw_allreduce_peer(C.byref(steps[1]), sp), "nrx_adamw_allreduce_peer")
    L.check(lib.nrx_adamw_allreduce_peer(C.byref(steps[0]), sp), "nrx_adamw_allreduce_peer")
    torch.cuda.synchronize()
    assert int(status[1].item()) & 2 and int(status[0].item()) & 2
    for r in range(2):
        assert torch.equal(ranks[r]["p"], before[r]), f"rank {r}: a dead exchange must not touch the parameters"
        assert float(ranks[r]["m"].abs().sum()) == 0.0
    assert int(ranks[1]["sig"][128].item()) == epoch1


def test_peer_step_argument_checks():
    L, p0, ranks, hp, steps = _setup(1, 16)
    lib = L.load()
    steps[0].n = 6
    with pytest.raises(L.NrxError):
        L.check(lib.nrx_adamw_allreduce_peer(C.byref(steps[0]), None), "nrx_adamw_allreduce_peer")
    steps[0].n = 16
    steps[0].world = 99
    with pytest.raises(L.NrxError):
        L.check(lib.nrx_adamw_allreduce_peer(C.byref(steps[0]), None), "nrx_adamw_allreduce_peer")
