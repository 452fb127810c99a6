"""Pairwise spectral-distance sweep (config 5) on one B200: 4 s / 16 kHz synthetic clips, D = 645,864
spectrogram bins per clip.  Reports pairs/s of the tiled kernel, its FP32-instruction roofline, and the
oracle's per-pair CPU path on a bounded sample."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from topo_audio_autoencoder_b200 import precompute_distances as pd
from oracle import distance_oracle as do

n = int(os.environ.get("N_CLIPS", "1024"))
g = torch.Generator().manual_seed(511990)
audio = torch.randn(n, 1, 64000, generator=g) * 0.1
t0 = time.perf_counter()
spec, seg = pd.multiscale_spectrograms(audio.cuda())
prep = pd.PreparedSpectra(spec, seg)
torch.cuda.synchronize()
t_front = time.perf_counter() - t0
del spec
for _ in range(2):
    prep.rows(0, min(n, 128))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = prep.rows(0, n)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
pairs_computed = n * n                      # full rows (both triangles): no collective needed when row-sharded
d = prep.d
instr = pairs_computed * d * 4              # FADD, FFMA, FADD, FADD|.| per pair-element
# B200: 148 SMs x 128 FP32 lanes x 1.965 GHz
peak_instr = 148 * 128 * 1.965e9
m = 6
t0 = time.perf_counter()
do.batch_audio_distance(audio[:m], audio[m:2 * m])
t_cpu = (time.perf_counter() - t0) / m
print(json.dumps({
    "metric": "pairwise_spectral_distance_pairs_per_sec", "n_clips": n, "bins_per_clip": d, "ms": ms,
    "value": pairs_computed / (ms * 1e-3), "unique_pairs_per_sec": n * (n - 1) / 2 / (ms * 1e-3),
    "front_half_s": t_front, "fp32_instr_per_sec": instr / (ms * 1e-3), "frac_of_fp32_issue_peak": instr / (ms * 1e-3) / peak_instr,
    "cpu_oracle_pairs_per_sec": 1.0 / t_cpu, "cpu_threads": torch.get_num_threads(),
    "symmetric": bool(torch.equal(out, out.t()))}))
