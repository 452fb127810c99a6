"""GPU parity straight against the REFERENCE's outputs: the fixtures tests/golden/ref_*.npz were produced by the
reference's own encoder.py / precompute_distances.py / custom_sccn.py (oracle/make_golden_glue.py; only the absent
third-party imports are stubbed), so nothing in these comparisons is a hand restatement except the fp64 anchor used
for multi-layer chains.  Tolerance: the north star's rtol 1e-5 / atol 1e-6 unless a comment says otherwise."""
import copy
import pickle
import wave

import numpy as np
import pytest
import torch

from oracle.param_fill import fill_by_name, value_for
from oracle.sccn_oracle import OracleSCCN
from tests.helpers import (NAMES, assert_close, assert_fp32_equivalent, load_golden, load_sccn_case, ref_sccn_cases,
                           run_sccn_case)

pytestmark = pytest.mark.gpu


def test_binary_gumbel_against_reference():
    import topo_audio_autoencoder_b200 as T
    fx = load_golden("ref_gumbel")
    for i in range(int(fx["n_cases"])):
        gate = T.BinaryGumbel().train()
        gate.set_temperature(float(fx[f"c{i}_temp"]))
        lg = torch.from_numpy(fx[f"c{i}_logits"]).cuda().requires_grad_(True)
        out = gate(lg, torch.from_numpy(fx[f"c{i}_gumbels"]).cuda())
        (g,) = torch.autograd.grad(out, lg, torch.from_numpy(fx[f"c{i}_up"]).cuda())
        assert_close(f"ref/gumbel/{i}/out", out, torch.from_numpy(fx[f"c{i}_out"]))
        assert_close(f"ref/gumbel/{i}/grad", g, torch.from_numpy(fx[f"c{i}_grad"]))
    gate = T.BinaryGumbel()
    gate.set_temperature(0.001)
    assert np.float32(gate.current_temp) == fx["min_temp_after_floor"]


@pytest.mark.parametrize("case", ["ref_glue_n6", "ref_glue_n9"])
def test_encoder_glue_against_reference(case):
    import topo_audio_autoencoder_b200 as T
    fx = load_golden(case)
    n, ch, seed = int(fx["n_vertices"]), int(fx["channels"]), int(fx["seed"])
    head = T.ComplexHead(n, embedding_dim=ch, min_active_vertices=int(fx["min_active"]),
                         max_active_vertices=int(fx["max_active"]), gate="binary_gumbel", bias_on="probs")
    with torch.no_grad():
        for name, p in head.named_parameters():
            if name.split(".")[0] in head._embedding_names:
                p.copy_(value_for(seed, name, p))
        head.vertex_bias.fill_(float(fx["vertex_bias"]))
    head = head.cuda()
    # split_simplices (encoder.py:291-297): exact
    parts = head.split_simplices(torch.from_numpy(fx["split_in"]).cuda())
    for k, p in zip(NAMES, parts):
        assert torch.equal(p.cpu(), torch.from_numpy(fx[f"split_{k}"])), f"split {k}"
    # get_active_simplex_embeddings (encoder.py:227-263)
    probs = [torch.from_numpy(fx[f"prob_{k}"]).cuda().requires_grad_(True) for k in NAMES]
    emb = head.get_active_simplex_embeddings(*probs, "cuda")
    for r, k in enumerate(NAMES):
        assert emb["active_indices"][k].dtype == torch.int64
        assert np.array_equal(emb["active_indices"][k].cpu().numpy(), fx[f"active_{k}"]), f"active indices {k}"
        assert_close(f"ref/{case}/emb{r}", emb[f"rank_{r}"], torch.from_numpy(fx[f"emb_{r}"]))
    tables = [t for nm in head._embedding_names for t in (getattr(head, nm)[0].weight, getattr(head, nm)[1].weight, getattr(head, nm)[1].bias)]
    ups = [torch.from_numpy(fx[f"emb_up_{r}"]).cuda() for r in range(4)]
    grads = torch.autograd.grad([emb[f"rank_{r}"] for r in range(4)], probs + tables, ups, allow_unused=True)
    names = [f"prob_{k}" for k in NAMES] + [f"{t}_{w}" for t in ("vtab", "etab", "ttab", "qtab") for w in ("weight", "ln_w", "ln_b")]
    for nm, g, leaf in zip(names, grads, probs + tables):
        g = torch.zeros_like(leaf) if g is None else g
        # table / LayerNorm gradients are sums over the active rows: relative to their accumulated magnitude
        loose = not nm.startswith("prob_")
        assert_close(f"ref/{case}/embgrad/{nm}", g, torch.from_numpy(fx[f"embgrad_{nm}"]),
                     rtol=1e-5, atol=2e-6 if loose else 1e-6)
    # penalties (encoder.py:199-225)
    for i, (v, want, wg) in enumerate(zip(fx["vp_in"], fx["vp_out"], fx["vp_grad"])):
        vl = torch.from_numpy(v).cuda().requires_grad_(True)
        p = head.compute_vertex_penalty(vl)
        (g,) = torch.autograd.grad(p, vl, allow_unused=True)
        g = torch.zeros_like(vl) if g is None else g
        assert_close(f"ref/{case}/vertex-penalty/{i}", p, torch.tensor(want))
        assert_close(f"ref/{case}/vertex-penalty-grad/{i}", g, torch.from_numpy(wg))
    pl = [torch.from_numpy(fx[f"prob_{k}"]).cuda().requires_grad_(True) for k in NAMES]
    ent = head.compute_entropy_loss(*pl)
    assert_close(f"ref/{case}/entropy", ent, torch.tensor(fx["entropy"]))
    for k, g in zip(NAMES, torch.autograd.grad(ent, pl)):
        assert_close(f"ref/{case}/entropy-grad/{k}", g, torch.from_numpy(fx[f"entgrad_{k}"]))


def _write_wav(path, x):
    pcm = np.round(np.clip(x, -1, 1) * 32767.0).astype(np.int16)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(pcm.tobytes())


def test_distances_against_reference(tmp_path):
    from topo_audio_autoencoder_b200 import precompute_distances as pd
    fx = load_golden("ref_distance")
    a, b = torch.from_numpy(fx["bmd_a"]).cuda(), torch.from_numpy(fx["bmd_b"]).cuda()
    for norm in ("L1", "L2"):
        for rel in (0, 1):
            assert_close(f"ref/bmd/{norm}/{rel}", pd.batch_mean_difference(a, b, norm=norm, relative=bool(rel)),
                         torch.from_numpy(fx[f"bmd_{norm}_{rel}"]))
    d = pd.BatchAudioDistance()(torch.from_numpy(fx["bad_x"]).cuda(), torch.from_numpy(fx["bad_y"]).cuda())["spectral_distance"]
    # sums of 10^5 fp32 terms per pair: 2e-5 relative (cuFFT vs pocketfft in front, different summation order behind)
    assert_close("ref/batch-audio-distance", d, torch.from_numpy(fx["bad_out"]), rtol=2e-5, atol=1e-6)

    # compute_distances (precompute_distances.py:51-153): same wav files on disk -> same two output files
    audio, lengths, order = fx["cd_audio"], fx["cd_lengths"], [str(s) for s in fx["cd_file_order"]]
    adir, sdir = tmp_path / "audio", tmp_path / "out"
    adir.mkdir(); sdir.mkdir()
    for name, row, ln in zip(order, audio, lengths):
        _write_wav(adir / name, row[0, :ln])
    pd.compute_distances(adir, sdir, batch_size=4)
    got = torch.load(sdir / "distance_matrix.pt")
    with open(sdir / "neighbors.pkl", "rb") as f:
        nb = pickle.load(f)
    f2i = nb["__file_to_idx__"]
    assert sorted(f2i.values()) == list(range(len(order)))
    ours_of_ref = [f2i[str(adir / name)] for name in order]           # reference row -> our row (glob order may differ)
    perm = torch.tensor(ours_of_ref)
    want = torch.from_numpy(fx["cd_matrix"])
    # zero-padded clips: every pair carries |log(1e-7) - log(s)| terms of ~16 per bin, so the entries are ~100 and their
    # fp32 sums over 10^5 bins agree to a few 1e-5 relative
    assert_close("ref/compute-distances/matrix", got[perm][:, perm], want, rtol=5e-5, atol=1e-6)
    assert got.dtype == torch.float32 and tuple(got.shape) == (len(order), len(order))
    assert torch.equal(got, got.t()) and (torch.diagonal(got) == 0).all()
    for ref_row, name in enumerate(order):
        rec = nb[str(adir / name)]
        assert sorted(rec.keys()) == [str(s) for s in fx["cd_neighbor_keys"]]
        assert rec["index"] == ours_of_ref[ref_row]
        want_names = [str(adir / order[j]) for j in fx["cd_sorted_idx"][ref_row].tolist()]
        assert rec["sorted_neighbors"] == want_names, f"neighbour order of {name}"
        np.testing.assert_allclose(rec["sorted_distances"], fx["cd_sorted_vals"][ref_row], rtol=5e-5, atol=1e-6)


@pytest.mark.parametrize("case", ref_sccn_cases())
def test_sccn_against_reference_forward(case):
    """GradientSCCN.forward / GradientSCCNLayer.forward as the reference itself ran them (custom_sccn.py:62-162):
    outputs, feature gradients, operator-value gradients and parameter gradients."""
    import topo_audio_autoencoder_b200 as T
    fx = load_golden(case)
    ch, max_rank, n_layers, seed = int(fx["channels"]), int(fx["max_rank"]), int(fx["n_layers"]), int(fx["seed"])
    ours = fill_by_name(T.GradientSCCN(ch, max_rank, n_layers), seed).cuda()
    anchor = fill_by_name(OracleSCCN(ch, max_rank, n_layers), seed).double()
    ours.train(bool(fx["train"])); anchor.train(bool(fx["train"]))
    feats, inc, adj, ups = load_sccn_case(fx, "cuda")
    out, gf, gm, gp = run_sccn_case(ours, feats, inc, adj, ups)
    f64 = lambda d: {k: (v.double().cpu() if v is not None else None) for k, v in d.items()}      # noqa: E731
    o64, gf64, gm64, gp64 = run_sccn_case(anchor, f64(feats), f64(inc), f64(adj), f64(ups))
    for k, v in out.items():
        assert (v is None) == bool(fx[f"outnone_{k}"]), (case, k)
        if v is not None:
            assert_fp32_equivalent(f"ref/{case}/out/{k}", v, torch.from_numpy(fx[f"out_{k}"]), o64[k])
    for k, g in gf.items():
        if f"gx_{k}" in fx.files:
            assert_fp32_equivalent(f"ref/{case}/dx/{k}", g, torch.from_numpy(fx[f"gx_{k}"]), gf64[k])
    for (kind, k), g in gm.items():
        if f"g{kind}_{k}" in fx.files:
            if g is None:               # an operator of an empty rank: the reference's gradient is an empty tensor
                assert fx[f"g{kind}_{k}"].size == 0, (case, kind, k)
                continue
            assert_fp32_equivalent(f"ref/{case}/d{kind}/{k}", g, torch.from_numpy(fx[f"g{kind}_{k}"]), gm64[(kind, k)])
    floor = 5e-6 * max(float(np.abs(fx[f]).max()) for f in fx.files if f.startswith("gp_"))
    for k, g in gp.items():
        if f"gp_{k}" in fx.files:
            if g is None:               # parameters of an empty rank: the reference reports an all-zero gradient
                assert not fx[f"gp_{k}"].any(), (case, k)
                continue
            assert_fp32_equivalent(f"ref/{case}/dparam/{k}", g, torch.from_numpy(fx[f"gp_{k}"]), gp64[k], floor=floor)
        else:
            assert g is None or g.abs().max().item() == 0, (case, k)
