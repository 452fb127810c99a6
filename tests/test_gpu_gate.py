"""GPU parity: Hard Concrete (builder's spec, parity unpinned) and BinaryGumbel vs the oracle."""
import pytest
import torch

from oracle import gate_oracle as go
from oracle.glue_oracle import rank_offsets
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("training,ste", [(True, False), (True, True), (False, False)])
@pytest.mark.parametrize("n,batch", [(20, 3), (7, 1)])
def test_hard_concrete_forward_backward(training, ste, n, batch):
    import topo_audio_autoencoder_b200 as T
    off = rank_offsets(n)
    N = off[4]
    g = torch.Generator().manual_seed(511990)
    logits = torch.randn(batch, N, generator=g)
    u = torch.rand(batch, N, generator=g).clamp_(1e-6, 1 - 1e-6)
    params = torch.tensor([0.66, -0.1, 1.1, 2.0, 1.0, 1.0, 1.5])
    up = torch.randn(batch, N, generator=g)

    lc, pc = logits.clone().requires_grad_(True), params.clone().requires_grad_(True)
    z_ref = go.hard_concrete(lc, u, pc[0], pc[1], pc[2], pc[3:], off, training=training, ste=ste)
    (z_ref * up).sum().backward()

    lg, pg = logits.cuda().requires_grad_(True), params.cuda().requires_grad_(True)
    z = T.hard_concrete(lg, u.cuda() if training else None, pg, off, training=training, ste=ste)
    (z * up.cuda()).sum().backward()

    assert_close(f"hard-concrete/z/train={training}/ste={ste}/n={n}", z, z_ref)
    # exact zeros / ones agree except for elements whose pre-clamp value is within float rounding of the clamp edge
    mism = ((z.detach().cpu() == 0) != (z_ref.detach() == 0)) | ((z.detach().cpu() == 1) != (z_ref.detach() == 1))
    assert mism.sum().item() <= 1, f"{mism.sum().item()} clamp-edge disagreements"
    if n == 20:
        assert (z.detach() == 0).any() and (z.detach() == 1).any()
    assert_close(f"hard-concrete/dlogits/train={training}/ste={ste}/n={n}", lg.grad, lc.grad)
    # parameter gradients are sums over B*N elements: compare relative to their accumulated magnitude
    assert_close(f"hard-concrete/dparams/train={training}/ste={ste}/n={n}", pg.grad, pc.grad, rtol=1e-4, atol=1e-4)


def test_hard_concrete_module_has_trainer_contract():
    import topo_audio_autoencoder_b200 as T
    off = rank_offsets(6)
    hc = T.HardConcrete(off).cuda()
    hc.current_temp = 0.5                       # trainer.py:266 writes this attribute
    hc.train()
    z = hc(torch.randn(2, off[4], device="cuda"))
    assert z.shape == (2, off[4]) and z.min() >= 0 and z.max() <= 1
    hc.eval()
    z1, z2 = hc(torch.ones(1, off[4], device="cuda")), hc(torch.ones(1, off[4], device="cuda"))
    assert torch.equal(z1, z2)


def test_binary_gumbel_training_branch():
    import topo_audio_autoencoder_b200 as T
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(6195, generator=g)
    gumbels = -torch.empty(2, 6195).exponential_(generator=g).log()
    up = torch.randn(6195, generator=g)
    for temp in (1.0, 0.1):
        lc = logits.clone().requires_grad_(True)
        p_ref = go.binary_gumbel_train(lc, gumbels, temp)
        (p_ref * up).sum().backward()
        gate = T.BinaryGumbel().cuda().train()
        gate.set_temperature(temp)
        lg = logits.cuda().requires_grad_(True)
        p = gate(lg, gumbels.cuda())
        (p * up.cuda()).sum().backward()
        assert_close(f"binary-gumbel/p/temp={temp}", p, p_ref)
        assert_close(f"binary-gumbel/dlogits/temp={temp}", lg.grad, lc.grad)
        assert torch.equal(p.detach().cpu() == 0, p_ref.detach() == 0)
