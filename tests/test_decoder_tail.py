"""CPU: the decoder consumer (stock PyTorch, reference decoder.py:131-175) against the fixture the reference's own
AudioDecoder.forward produced (oracle/make_golden_glue.py), and the batched consumer of compact rows against the
per-sample path -- including the compaction contract: rows past a sample's active count must never be read."""
import numpy as np
import pytest
import torch

from oracle.param_fill import fill_by_name
from tests.helpers import assert_close, bits_equal, load_golden


def _tail(seed):
    from topo_audio_autoencoder_b200.decoder import DecoderTail
    return fill_by_name(DecoderTail(64, 250, 16), seed).train()


def test_tail_reproduces_reference_decoder_forward():
    fx = load_golden("ref_decoder_tail")
    tail = _tail(int(fx["seed"]))
    xs = {f"rank_{r}": torch.from_numpy(fx[f"x_rank_{r}"]).requires_grad_(True) for r in range(4)}
    y = tail(xs)
    assert bits_equal(y, torch.from_numpy(fx["out"])), "decoder tail output differs from the reference's"
    named = list(tail.named_parameters())
    grads = torch.autograd.grad(y, list(xs.values()) + [p for _, p in named], torch.from_numpy(fx["up"]), allow_unused=True)
    for r in range(4):
        assert_close(f"decoder-tail/dx{r}", grads[r], torch.from_numpy(fx[f"gx_rank_{r}"]), rtol=1e-6, atol=1e-7)
    for (name, _), g in zip(named, grads[4:]):
        want = float(fx[f"gpnorm_{name}"])
        got = 0.0 if g is None else g.double().norm().item()
        assert abs(got - want) <= 1e-5 * max(want, 1e-12) + 1e-9, name


def test_state_dict_keys_match_reference_decoder():
    """every non-SCCN key of the reference AudioDecoder (read from the fixture's gradient-norm names) exists here"""
    fx = load_golden("ref_decoder_tail")
    want = sorted(k[len("gpnorm_"):] for k in fx.files if k.startswith("gpnorm_"))
    assert sorted(k for k, _ in _tail(1).named_parameters()) == want


@pytest.mark.parametrize("full", [False, True])
def test_batched_consumer_equals_per_sample(full):
    tail = _tail(5)
    g = torch.Generator().manual_seed(3)
    if full:
        counts = torch.tensor([[6, 15, 20, 15]] * 3)
    else:
        counts = torch.tensor([[4, 3, 0, 0], [6, 15, 20, 15], [1, 0, 2, 1], [4, 9, 7, 0]])
    tot = counts.sum(0).tolist()
    # compact buffers sized by the allocation bound: the rows past the live total hold garbage that must not be read
    xs = [torch.cat([torch.randn(tot[r], 64, generator=g), torch.full((5, 64), float("nan"))]) for r in range(4)]
    leaves = [x.clone().requires_grad_(True) for x in xs]
    out = tail.forward_batched(leaves, counts)
    up = torch.randn(out.shape, generator=g)
    gb = torch.autograd.grad(out, leaves, up)
    starts = torch.cumsum(counts, 0) - counts
    leaves2 = [x.clone().requires_grad_(True) for x in xs]
    outs = []
    for b in range(counts.shape[0]):
        sample = {f"rank_{r}": (leaves2[r][starts[b, r]:starts[b, r] + counts[b, r]] if counts[b, r] else None) for r in range(4)}
        outs.append(tail(sample))
    ref = torch.cat(outs)
    gs = torch.autograd.grad(ref, leaves2, up)
    assert torch.isfinite(out).all()
    # the same stock PyTorch modules on both sides; padded batched attention sums the keys in a different order
    assert_close(f"decoder-tail/batched/full={full}", out, ref, rtol=1e-5, atol=5e-6)
    for r in range(4):
        assert torch.equal(gb[r][tot[r]:], torch.zeros(5, 64)), "dead rows received gradient"
        assert_close(f"decoder-tail/batched/dx{r}/full={full}", gb[r][:tot[r]], gs[r][:tot[r]], rtol=1e-4, atol=1e-4)


def test_empty_inputs_are_refused():
    tail = _tail(5)
    xs = [torch.randn(8, 64) for _ in range(4)]
    with pytest.raises(ValueError):
        tail.forward_batched(xs, torch.tensor([[0, 1, 1, 1]]))
    with pytest.raises(ValueError):
        tail.forward_batched(xs, torch.tensor([[2, 0, 0, 0]]))
