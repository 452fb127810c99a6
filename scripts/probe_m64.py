"""GPU: where do the 64 accumulator rows of an M = 64 tcgen05.mma land in tensor memory?"""
import sys
import torch
from topo_audio_autoencoder_b200 import _lib
sys.path.insert(0, ".")
from topo_audio_autoencoder_b200._lib import lib, ptr, stream  # noqa: E402

g = torch.Generator().manual_seed(5)
rows = 128
a = torch.randn(rows, 64, generator=g).cuda()
b = torch.randn(rows, 64, generator=g).cuda()
out = torch.zeros(128, 64, device="cuda")
rc = _lib.load_debug().topo_debug_gemm_bf16x3(ptr(a), ptr(b), rows, 3, 16384, 1024, 2048, ptr(out), stream())
torch.cuda.synchronize()
want = (a.double().t() @ b.double()).float()          # [64, 64]
print("rc", rc)
for lane in range(128):
    row = out[lane]
    if row.abs().max().item() == 0:
        print(f"lane {lane:3d}: zero")
        continue
    d = (want - row.unsqueeze(0)).abs().max(dim=1).values
    j = int(d.argmin())
    print(f"lane {lane:3d}: matches accumulator row {j:2d} (err {d[j].item():.2e})")
