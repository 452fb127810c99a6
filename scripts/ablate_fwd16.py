"""GPU: differential timing of the second-generation forward at the rank-3 size (parts switched off one at a time)."""
# runs on libtopo_b200_debug.so (csrc/build.py builds it next to the product library): the hooks do not exist in the product
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from topo_audio_autoencoder_b200._lib import load_debug, check, ptr, stream  # noqa: E402
from topo_audio_autoencoder_b200.custom_sccn import _make_params  # noqa: E402

raw = lib = load_debug()          # kernels AND hooks from the debug twin (the hooks set globals of that library)
raw.topo_debug_fwd16_mask.argtypes = [C.c_int]
raw.topo_debug_fwd16_mask.restype = None
rows, ch, n_msgs = 310080, 64, 2
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g).cuda()  # noqa: E731
aggs = [rnd(rows, ch) for _ in range(n_msgs)]
ws = [rnd(ch, ch) * 0.2 for _ in range(n_msgs)]
scales = [torch.ones(1).cuda() for _ in range(n_msgs)]
x = rnd(rows, ch)
tensors = [rnd(ch, ch) * 0.2, rnd(ch) * 0.1, rnd(ch) * 0.3, rnd(1), 1 + 0.1 * rnd(ch), 0.1 * rnd(ch)]
pad = -(-rows // 128) * 128
saved = ([torch.zeros(pad, ch, device="cuda") for _ in range(n_msgs)], [torch.zeros(pad, ch, device="cuda") for _ in range(n_msgs)],
         torch.zeros(3, rows, device="cuda"))
params = _make_params(ch, n_msgs, aggs, ws, scales, x, tensors, 1e-5, True, saved, True)
out = torch.zeros(rows, ch, device="cuda")
flush = torch.empty(64 * 1024 * 1024, device="cuda")
names = {0: "full kernel", 1: "no saved_m/pre stores", 2: "no GELU", 4: "no image stores", 8: "no prefetch loads", 16: "no MMA",
         32: "no tile-end reads/stores", 64: "no TMEM loads", 256: "no residual loads", 1 | 32: "no stores at all",
         8 | 256: "no loads at all", 1 | 8 | 32 | 256: "no global traffic",
         2 | 4 | 16 | 64: "memory only", 1 | 2 | 4 | 8 | 16 | 32 | 64 | 256: "skeleton only"}
for mask, name in names.items():
    raw.topo_debug_fwd16_mask(mask)
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.topo_sccn_combine_fwd_tc2(C.byref(params), rows, None, ptr(out), stream()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"mask {mask:3d} {name:28s} {min(ts[1:]):8.1f} us", flush=True)
raw.topo_debug_fwd16_mask(0)

# set-up cost: one tile per CTA
for r_ in (128, 148 * 128):
    ts = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.topo_sccn_combine_fwd_tc2(C.byref(params), r_, None, ptr(out), stream()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"rows {r_:6d} (one tile per CTA): {min(ts[1:]):8.1f} us", flush=True)

# in-kernel timeline of CTA 0 (globaltimer, ns)
raw.topo_debug_fwd16_stamps.argtypes = [C.c_void_p]
raw.topo_debug_fwd16_stamps.restype = None
st = torch.zeros(8, dtype=torch.int64, device="cuda")
raw.topo_debug_fwd16_stamps(st.data_ptr())
for r_ in (128, rows):
    for it in range(3):
        check(lib.topo_sccn_combine_fwd_tc2(C.byref(params), r_, None, ptr(out), stream()))
        torch.cuda.synchronize()
    v = st.tolist()
    print(f"rows {r_}: set-up {v[1]-v[0]} ns, barrier {v[2]-v[1]}, first tile units {v[3]-v[2]}, tile end {v[4]-v[3]}, rest {v[5]-v[4]}, teardown {v[6]-v[5]}, total {v[6]-v[0]}")
raw.topo_debug_fwd16_stamps(None)
