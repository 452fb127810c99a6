"""(Rewritten to time a CUDA-graph replay after the eager version turned out to measure the host launch path; this form has not
been run on a GPU yet.)  What do the aggregation launches cost when no row is alive (grids are sized by the bound B * n_r)?  And with every row alive?
Times topo_sccn_aggregate_fwd / _bwd on a batch of 64 complexes: all simplices inactive, ~35 % active (random), all active."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import topo_audio_autoencoder_b200 as T
from topo_audio_autoencoder_b200._lib import lib, check, ptr, ptr_array, stream
from topo_audio_autoencoder_b200.encoder_complex import active_sets
from topo_audio_autoencoder_b200.custom_sccn import BatchedComplex
from topo_audio_autoencoder_b200.rectifier import rectify_batch

dev = torch.device("cuda")
n, B, ch = 20, 64, 64
head = T.ComplexHead(n, embedding_dim=ch).to(dev)
tables = head._tables
N = tables.total
cnt = tables.counts


def run(label, probs):
    pos, act, counts, row_off = active_sets(probs, tables)
    cx = BatchedComplex(tables, probs, pos, act, counts, row_off, [B * c for c in cnt])
    xs = [torch.randn(B * c, ch, device=dev) for c in cnt]
    mk = lambda r: torch.zeros(B * cnt[r], ch, device=dev)
    down = [mk(0), mk(1), mk(2), None]
    up = [None, mk(1), mk(2), mk(3)]
    same = [mk(r) for r in range(4)]
    g_x = [mk(r) for r in range(4)]
    g_probs = torch.zeros_like(probs)
    view = cx.view(probs)

    def fwd():
        check(lib.topo_sccn_aggregate_fwd(tables.handle, C.byref(view), ch, ptr_array(xs, 4), ptr_array(down, 4), ptr_array(up, 4),
                                          ptr_array(same, 4), stream()))

    def bwd():
        check(lib.topo_sccn_aggregate_bwd(tables.handle, C.byref(view), ch, ptr_array(xs, 4), ptr_array([None, None, down[2], None], 4),
                                          ptr_array([None, None, up[2], up[3]], 4), ptr_array(down, 4), ptr_array(up, 4),
                                          ptr_array(same, 4), ptr_array(g_x, 4), ptr(g_probs), stream()))

    out = []
    for f in (fwd, bwd):
        # device time: 20 calls captured into one CUDA graph (eager launches from Python are bound by the host: ~15 us per call)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                f()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(20):
                f()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 20 * 1000)
    live = row_off[:, B].tolist()
    print(f"{label:12s} live rows {live}  aggregate_fwd {out[0]:7.1f} us  aggregate_bwd {out[1]:7.1f} us")


run("none alive", torch.zeros(B, N, device=dev))
g = torch.Generator(device="cpu").manual_seed(1)
p = (torch.rand(B, N, generator=g) < 0.8).float().to(dev) * torch.rand(B, N, generator=g).to(dev)
p[:, :n] = torch.rand(B, n, generator=g).to(dev) * 0.5 + 0.5
run("rectified", rectify_batch(p, head.constraints).detach())
run("all alive", torch.rand(B, N, generator=g).to(dev) * 0.5 + 0.5)
