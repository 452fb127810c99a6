"""Diagnostic: localise a gradient discrepancy of the stage by loss component (GPU vs oracle)."""
import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import topo_audio_autoencoder_b200 as T
from oracle import gate_oracle as go, glue_oracle as glo, rectifier_oracle as ro
from oracle.sccn_oracle import OracleSCCN
from topo_audio_autoencoder_b200.encoder_complex import _PenaltiesFn

n, B, C, L = 9, 3, 64, 2
torch.manual_seed(2)
stage = T.ComplexStage(n, channels=C, n_layers=L, bias_on="logits").cuda().train()
head = stage.head
tab, off = ro.make_tables(n), glo.rank_offsets(n)
g = torch.Generator().manual_seed(511990)
logits = torch.randn(B, off[4], generator=g)
u = torch.rand(B, off[4], generator=g).clamp_(1e-6, 1 - 1e-6)
ref = OracleSCCN(C, 3, L).train()
ref.load_state_dict({k: v.detach().cpu() for k, v in stage.sccn.state_dict().items()})
emb = [tuple(t.detach().cpu() for t in (getattr(head, nm)[0].weight, getattr(head, nm)[1].weight, getattr(head, nm)[1].bias)) for nm in head._embedding_names]
loc = torch.relu(torch.cat([p.detach().cpu() for p in (head.vertex_bias, head.edge_bias, head.triangle_bias, head.tetra_bias)]))

def ours(w_sccn, w_vp, w_ent):
    lg = logits.cuda().requires_grad_(True)
    out = stage(lg, u.cuda(), sync=True)
    loss = w_sccn * sum(out[f"rank_{r}"].pow(2).sum() for r in range(4)) + w_vp * out["vertex_penalty"].sum() + w_ent * out["entropy_loss"].sum()
    loss.backward()
    return lg.grad.cpu(), out

def oracle(w_sccn, w_vp, w_ent):
    lc = logits.clone().requires_grad_(True)
    z = go.hard_concrete(lc, u, head.sampler.current_temp, -0.1, 1.1, loc, off)
    loss = 0.0
    for b in range(B):
        e, (adj, inc), rect = glo.complex_from_probs(z[b], n, head.vertex_bias.detach().cpu(), tab, emb, False)
        o = ref({f"rank_{r}": e[f"rank_{r}"] for r in range(4)}, inc, adj)
        loss = loss + w_sccn * sum(o[f"rank_{r}"].pow(2).sum() for r in range(4)) + w_vp * glo.vertex_penalty(rect[0], 8, 16) + w_ent * glo.entropy_loss(*rect)
    loss.backward()
    return lc.grad

for name, w in (("sccn only", (1, 0, 0)), ("vp only", (0, 1, 0)), ("ent only", (0, 0, 1)), ("all", (1, 0.3, 0.7))):
    a, out = ours(*w)
    b = oracle(*w)
    err = (a - b).abs()
    i = err.argmax().item()
    print(f"{name:10s} max|err|={err.max().item():.3e} max|ref|={b.abs().max().item():.3e} at sample {i // off[4]} idx {i % off[4]} ours={a.flatten()[i].item():.6e} ref={b.flatten()[i].item():.6e}")
    print("   vertex grads ours", a[0, :9].tolist())
    print("   vertex grads ref ", b[0, :9].tolist())

# penalties alone
rect = out["rectified"].detach().clone().requires_grad_(True)
vp, ent = _PenaltiesFn.apply(rect, head._tables, 8, 16)
(0.3 * vp.sum() + 0.7 * ent.sum()).backward()
rc = out["rectified"].detach().cpu().clone().requires_grad_(True)
l = 0.0
for b in range(B):
    parts = torch.split(rc[b], tab.sizes)
    l = l + 0.3 * glo.vertex_penalty(parts[0], 8, 16) + 0.7 * glo.entropy_loss(*parts)
l.backward()
print("penalties joint: max err", (rect.grad.cpu() - rc.grad).abs().max().item(), "vp", vp.tolist())
print(" ours", rect.grad[0, :12].tolist()); print(" ref ", rc.grad[0, :12].tolist())
