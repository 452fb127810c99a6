"""Which torch (non-library) kernels run inside one eager stage step, with shapes and the autograd node that issued them."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import topo_audio_autoencoder_b200 as T
from bench import synthetic_inputs, SEED

regime = sys.argv[1] if len(sys.argv) > 1 else "full"
dev = torch.device("cuda")
torch.manual_seed(SEED)
kw = dict(gate="binary_gumbel", bias_on="probs") if regime == "full" else dict(gate="hard_concrete", bias_on="logits")
stage = T.ComplexStage(20, channels=64, n_layers=6, **kw).to(dev).train()
params = [p for p in stage.parameters() if p.requires_grad]
B = 64
lg_h, nz_h = synthetic_inputs(B, stage.head.total_simplices, regime, 0)
lg_d, nz_d = lg_h.to(dev), nz_h.to(dev)
g = torch.Generator().manual_seed(SEED)
ups = [torch.randn(B * c, 64, generator=g).to(dev) for c in stage.head._tables.counts]
ones = torch.ones(B, device=dev)


def step():
    for p in params:
        p.grad = None
    lg = lg_d.detach().requires_grad_(True)
    out = stage(lg, nz_d)
    torch.autograd.backward([out[f"rank_{r}"] for r in range(4)] + [out["vertex_penalty"], out["entropy_loss"]], ups + [ones, ones])


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=False) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    t = getattr(e, "self_device_time_total", 0.0)
    if t > 0:
        rows.append((t, e.count, e.key, str(e.input_shapes)[:110]))
tot = 0.0
for t, n, key, shp in sorted(rows, reverse=True):
    if key.startswith("aten::") or "elementwise" in key or "Memcpy" in key or "Memset" in key:
        print(f"{n:4d} x {t:9.1f} us  {key[:60]:60s} {shp}")
        if key.startswith("aten::"):
            tot += t
print("total aten self device time per step: %.1f us" % tot)
