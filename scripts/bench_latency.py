"""Per-sample drop-in latency (the reference trains with batch size 1, trainer.py:93): complex stage fwd+bwd at B = 1 and
B = 8 through (a) the batched API (ComplexStage, eager and CUDA-graph replay) and (b) the reference-signature path
(generate_complex -> explicit sparse operators -> GradientSCCN.forward, one clip at a time)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import topo_audio_autoencoder_b200 as T
from topo_audio_autoencoder_b200.graph import GraphedStep

n, C, L = 20, 64, 6
out = {}
for regime in ("full", "sparse"):
    kw = dict(gate="binary_gumbel", bias_on="probs") if regime == "full" else dict(gate="hard_concrete", bias_on="logits")
    torch.manual_seed(511990)
    stage = T.ComplexStage(n, channels=C, n_layers=L, **kw).cuda().train()
    N = stage.head.total_simplices
    g = torch.Generator().manual_seed(1)
    for B in (1, 8):
        logits = torch.randn(B, N, generator=g).cuda()
        noise = ((-torch.empty(2, B, N).exponential_(generator=g).log()) if regime == "full"
                 else torch.rand(B, N, generator=g).clamp_(1e-6, 1 - 1e-6)).cuda()
        ups = [torch.randn(B * c, C, generator=g).cuda() for c in stage.head._tables.counts] + [torch.ones(B).cuda()] * 2

        def eager():
            lg = logits.clone().requires_grad_(True)
            o = stage(lg, noise)
            torch.autograd.backward([o[f"rank_{r}"] for r in range(4)] + [o["vertex_penalty"], o["entropy_loss"]], ups)

        graphed = GraphedStep(stage, logits, noise, ups)

        def timeit(fn, reps=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        out[f"{regime}/B={B}/batched_eager_ms"] = timeit(eager)
        out[f"{regime}/B={B}/batched_graph_ms"] = timeit(lambda: graphed.replay(logits, noise))
    # reference-signature path, one clip
    head = stage.head
    lg1 = torch.randn(N, generator=g).cuda()
    nz1 = ((-torch.empty(2, N).exponential_(generator=g).log()) if regime == "full" else torch.rand(N, generator=g).clamp_(1e-6, 1 - 1e-6)).cuda()

    def per_sample():
        l1 = lg1.clone().requires_grad_(True)
        emb, mats = head.generate_complex(l1, nz1)
        o = stage.sccn(emb, mats.incidences, mats.adjacencies)
        sum(v.sum() for v in o.values() if v is not None).backward()

    out[f"{regime}/reference_signature_one_clip_ms"] = timeit(per_sample, reps=5)
print(json.dumps(out))
