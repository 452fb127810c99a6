"""GPU: in-kernel timeline (globaltimer) of the fused backward at the rank-3 size."""
# runs on libtopo_b200_debug.so (csrc/build.py builds it next to the product library): the hooks do not exist in the product
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from topo_audio_autoencoder_b200._lib import load_debug, check, ptr, stream, CombineGrads  # noqa: E402
from topo_audio_autoencoder_b200.custom_sccn import _make_params  # noqa: E402

raw = lib = load_debug()          # kernels AND hooks from the debug twin (the hooks set globals of that library)
raw.topo_debug_bwd_stamps.argtypes = [C.c_void_p]
raw.topo_debug_bwd_stamps.restype = None
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
check(lib.topo_sccn_combine_fwd_tc2(C.byref(params), rows, None, ptr(out), stream()))
g_out = rnd(rows, ch)
res = {"g_agg": [torch.zeros(rows, ch, device="cuda") for _ in range(n_msgs)], "wprod": [torch.zeros(ch, ch, device="cuda") for _ in range(n_msgs)],
       "g_x": torch.zeros(rows, ch, device="cuda"), "w1": torch.zeros(ch, ch, device="cuda"), "b1": torch.zeros(ch, device="cuda"),
       "w2": torch.zeros(ch, device="cuda"), "b2": torch.zeros(1, device="cuda"), "gamma": torch.zeros(ch, device="cuda"),
       "beta": torch.zeros(ch, device="cuda")}
grads = CombineGrads()
for k in range(n_msgs):
    grads.g_agg[k], grads.g_wprod[k] = ptr(res["g_agg"][k]), ptr(res["wprod"][k])
grads.g_x = ptr(res["g_x"])
grads.g_att_w1, grads.g_att_b1, grads.g_att_w2, grads.g_att_b2 = ptr(res["w1"]), ptr(res["b1"]), ptr(res["w2"]), ptr(res["b2"])
grads.g_ln_gamma, grads.g_ln_beta = ptr(res["gamma"]), ptr(res["beta"])
st = torch.zeros(64, dtype=torch.int64, device="cuda")
raw.topo_debug_bwd_stamps(st.data_ptr())
for r_ in (128, rows):
    ts = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.topo_sccn_combine_bwd_tc(C.byref(params), r_, None, ptr(g_out), C.byref(grads), stream()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    v = st.tolist()
    d = [v[i + 1] - v[i] for i in range(40) if v[i + 1] > 0 and v[i] > 0]
    print(f"rows {r_}: event time {min(ts[1:]):.1f} us; stamp deltas (ns): {d}")
raw.topo_debug_bwd_stamps(None)
