"""GPU: the full model around the hot path (BASELINE.json config 4) -- stock front-end -> complex stage -> decoder consumer
-> loss -> optimizer step -- and the batched path against the reference-signature per-sample path."""
import pytest
import torch

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


def _model(n=9, gate="hard_concrete", bias_on="logits"):
    import topo_audio_autoencoder_b200 as T
    torch.manual_seed(7)
    return T.AudioAutoencoder(num_vertices=n, gate=gate, bias_on=bias_on).cuda().train()


def test_batched_model_equals_the_per_sample_reference_signature_path():
    """forward of a batch (matrix-free SCCN on compact rows + batched decoder consumer) against one clip at a time through
    generate_complex -> explicit sparse operators -> AudioDecoder.forward(feature_embeddings, complex_matrices, length),
    the call sequence of the reference (audio2complex.py:46-51).  One clip of the batch has an empty complex."""
    model = _model()
    enc, dec = model.encoder, model.decoder
    N, B = enc.total_simplices, 4
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(B, N, generator=g).cuda()
    logits[2, :enc.num_vertices] = -60.0                                  # clip 2: every vertex gate closed
    noise = torch.rand(B, N, generator=g).clamp_(1e-6, 1 - 1e-6).cuda()
    bands = torch.randn(B, 16, 4000, generator=g).cuda()
    enc.logits = lambda x: logits                                         # plant the logits behind the front-end
    out, div, valid = model(bands, noise)
    assert valid.tolist() == [True, True, False, True] and out.shape == (3, 16, 4000)
    assert div["diversity"].shape == (B,) and div["binary_entropy"].shape == (B,)
    # the same batch through the stage only, to compare the hot path's own output before the decoder amplifies anything
    rect = enc.rectified_batch(logits, noise)
    cx = enc.batched_complex(rect, sync=True)
    xs = dec.sccn.forward_complex(cx, enc.embed(cx))
    rows, sccn_err = [], 0.0
    for b in (0, 1, 3):
        emb, mats = enc.generate_complex(logits[b], noise[b])
        per = dec.sccn(emb, mats.incidences, mats.adjacencies)
        for r in range(4):
            mine = torch.split(xs[r][:int(cx.host_counts[:, r].sum())], cx.host_counts[:, r].tolist())[b]
            scale = max(per[f"rank_{r}"].abs().max().item(), 1e-6)
            sccn_err = max(sccn_err, (mine - per[f"rank_{r}"]).abs().max().item() / scale)
        rows.append(dec(emb, mats, 4000))
    assert enc.generate_complex(logits[2], noise[2]) == (None, None, None)
    # matrix-free bf16x3 tensor-core execution vs explicit CSR operators, six layers with LayerNorm: fp32-equivalent
    assert sccn_err < 1e-4, f"SCCN outputs of the two paths differ by {sccn_err:.2e} of their scale"
    want = torch.cat(rows)
    # the decoder tail (stock PyTorch on both sides; batched == per-sample to 5e-6 in tests/test_decoder_tail.py) passes
    # those rows x0.1 through LayerNorms, which amplify the SCCN's fp32 rounding differences
    assert_close("full/batched-vs-per-sample", out, want, rtol=1e-3, atol=5e-3 * want.abs().max().item())


@pytest.mark.parametrize("gate", ["hard_concrete", "binary_gumbel"])
def test_training_step_runs_and_updates_both_parameter_groups(gate, tmp_path):
    import topo_audio_autoencoder_b200 as T
    model = _model(gate=gate, bias_on="logits" if gate == "hard_concrete" else "probs")
    tr = T.Trainer(model, checkpoint_dir=str(tmp_path), device="cuda", accumulate_grad_batches=2)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    tr.set_epoch(0)
    loss = tr.train_step([(torch.randn(3, 16, 4000, generator=g) * 0.3).cuda() for _ in range(2)])
    assert torch.isfinite(loss) and torch.isfinite(tr.last_grad_norm)
    after = model.state_dict()
    changed = {k for k in before if not torch.equal(before[k], after[k])}
    assert any(k.startswith("encoder.to_simplices") for k in changed), "front-end did not train"
    assert any(k.startswith("decoder.sccn.") for k in changed), "SCCN did not train"
    assert any(k.startswith("decoder.cross_attention") for k in changed), "decoder tail did not train"
    path = tr.save_checkpoint("epoch_0_iter_0")
    tr2 = T.Trainer(_model(gate=gate, bias_on="logits" if gate == "hard_concrete" else "probs"), checkpoint_dir=str(tmp_path))
    tr2.load_checkpoint(path)
    for k, v in tr2.model.state_dict().items():
        assert torch.equal(v, after[k]), k
