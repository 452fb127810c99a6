"""GPU: isolated timing of the forward generations at the rank-3 size (310,080 rows, 2 messages, residual, LayerNorm)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from topo_audio_autoencoder_b200._lib import lib, check, ptr, stream, ImageJob, WEIGHT_IMAGE_BYTES  # noqa: E402
from topo_audio_autoencoder_b200.custom_sccn import _make_params  # noqa: E402

rows, ch, n_msgs = 310080, 64, 2
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g).cuda()  # noqa: E731
aggs = [rnd(rows, ch) for _ in range(n_msgs)]
ws = [rnd(ch, ch) * 0.2 for _ in range(n_msgs)]
scales = [torch.ones(1).cuda() for _ in range(n_msgs)]
x = rnd(rows, ch)
tensors = [rnd(ch, ch) * 0.2, rnd(ch) * 0.1, rnd(ch) * 0.3, rnd(1), 1 + 0.1 * rnd(ch), 0.1 * rnd(ch)]
pad = -(-rows // 128) * 128
images = torch.empty(WEIGHT_IMAGE_BYTES * (1 + 2 * n_msgs), dtype=torch.uint8, device="cuda")
jobs = (ImageJob * (1 + n_msgs))()
jobs[0].w, jobs[0].scale, jobs[0].att_w1, jobs[0].dst = None, None, ptr(tensors[0]), images.data_ptr()
for k in range(n_msgs):
    jobs[1 + k].w, jobs[1 + k].scale, jobs[1 + k].att_w1 = ptr(ws[k]), ptr(scales[k]), ptr(tensors[0])
    jobs[1 + k].dst = images.data_ptr() + WEIGHT_IMAGE_BYTES * (1 + 2 * k)
check(lib.topo_sccn_prepare_images(jobs, 1 + n_msgs, ch, stream()))
saved = ([torch.zeros(pad, ch, device="cuda") for _ in range(n_msgs)], [torch.zeros(pad, ch, device="cuda") for _ in range(n_msgs)],
         torch.zeros(3, rows, device="cuda"))
out = torch.zeros(rows, ch, device="cuda")
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for name, fn, tf in (("gen 1 (3xTF32, chained)", lib.topo_sccn_combine_fwd_tc, False), ("gen 2 (bf16x3, one product)", lib.topo_sccn_combine_fwd_tc2, True)):
    sv = saved if tf else ([torch.zeros(rows, ch, device="cuda") for _ in range(n_msgs)], [torch.zeros(rows, ch, device="cuda") for _ in range(n_msgs)],
                           torch.zeros(3, rows, device="cuda"))
    params = _make_params(ch, n_msgs, aggs, ws, scales, x, tensors, 1e-5, True, sv, tf, images)
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(C.byref(params), rows, None, ptr(out), stream()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name:30s} {min(ts[1:]):8.1f} us", flush=True)
