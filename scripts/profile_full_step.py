"""GPU: where one full training step (config 4) spends its device time -- top kernels by total time (torch.profiler / CUPTI)."""
import sys
import torch
sys.path.insert(0, ".")
import topo_audio_autoencoder_b200 as T  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

if "--cudnn-benchmark" in sys.argv:
    torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = T.AudioAutoencoder(num_vertices=20, sccn_hidden_dim=64, gate="binary_gumbel", bias_on="probs")
tr = T.Trainer(model, device="cuda:0", accumulate_grad_batches=4)
g = torch.Generator().manual_seed(1)
mbs = [(torch.randn(64, 16, 4000, generator=g) * 0.3).cuda() for _ in range(4)]
for _ in range(3):
    tr.train_step(mbs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tr.train_step(mbs)
e1.record()
torch.cuda.synchronize()
print(f"step: {e0.elapsed_time(e1):.1f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.train_step(mbs)
    torch.cuda.synchronize()
rows = [(ev.device_time_total / 1e3, ev.count, ev.key) for ev in prof.key_averages() if ev.device_time_total > 0 and ev.device_type.name == "CUDA"]
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"kernel time total {tot:.1f} ms over {sum(r[1] for r in rows)} launches")
for t, n, k in rows[:40]:
    print(f"{t:9.2f} ms {100 * t / tot:5.1f}%  x{n:<5d} {k[:110]}")
