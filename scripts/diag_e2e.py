import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import topo_audio_autoencoder_b200 as T
torch.manual_seed(0)
dev = torch.device("cuda", 0)
stage = T.ComplexStage(20, channels=64, n_layers=6, gate="binary_gumbel", bias_on="probs").to(dev).train()
params = [p for p in stage.parameters()]
B, N = 64, stage.head.total_simplices
logits_h = torch.randn(B, N).pin_memory(); noise_h = (-torch.empty(2, B, N).exponential_().log()).pin_memory()
logits_d, noise_d = logits_h.to(dev), noise_h.to(dev)
ups = [torch.randn(B * c, 64, device=dev) for c in stage.head._tables.counts]
ones = torch.ones(B, device=dev)
def step(lg, nz):
    for p in params: p.grad = None
    lg = lg.detach().requires_grad_(True)
    out = stage(lg, nz)
    torch.autograd.backward([out[f"rank_{r}"] for r in range(4)] + [out["vertex_penalty"], out["entropy_loss"]], ups + [ones, ones])
    return out, lg
for i in range(3): step(logits_d, noise_d)
torch.cuda.synchronize()
def stats(tag):
    s = torch.cuda.memory_stats()
    print(tag, "alloc_retries", s["num_alloc_retries"], "device_allocs", s.get("num_device_alloc"), "device_frees", s.get("num_device_free"),
          "reserved GB", s["reserved_bytes.all.current"] / 1e9, "active GB peak", s["active_bytes.all.peak"] / 1e9)
stats("after warmup")
for mode in ("async", "sync_each", "host_io", "async"):
    t0 = time.perf_counter()
    for i in range(5):
        if mode == "host_io":
            lg = logits_h.to(dev, non_blocking=True); nz = noise_h.to(dev, non_blocking=True)
            out, lg = step(lg, nz)
            res = torch.cat([out["vertex_penalty"], out["entropy_loss"], lg.grad.sum().reshape(1)]).cpu()
        else:
            out, lg = step(logits_d, noise_d)
            if mode == "sync_each": torch.cuda.synchronize()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{mode:10s} cpu enqueue {1e3*(t1-t0)/5:.1f} ms/step, total {1e3*(t2-t0)/5:.1f} ms/step")
    stats(mode)
