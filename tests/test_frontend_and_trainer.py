"""The callers either side of the hot path (stock PyTorch in the reference): front-end, loss, dataset sampling."""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle.param_fill import fill_by_name, tensor_by_name
from tests.helpers import assert_close, load_golden


def test_grouped_front_end_reproduces_the_reference_modules():
    """frontend.ConvFrontEnd (16 band stacks as one grouped convolution) against the reference's own modules run band by
    band (oracle/make_golden_glue.gen_frontend), same name-keyed parameters."""
    from topo_audio_autoencoder_b200.frontend import ConvFrontEnd
    fx = load_golden("ref_frontend")
    seed, n = int(fx["seed"]), int(fx["n_vertices"])
    fe = fill_by_name(ConvFrontEnd(n), seed).eval()
    ref_names = [str(s) for s in fx["param_names"]]
    mine = sorted(k for k, _ in fe.named_parameters())
    assert set(mine) <= set(ref_names), "every front-end parameter exists under the same name in the reference encoder"
    x = tensor_by_name(seed, "frontend_input", (2, 16, 4000), "normal", 0.3)
    with torch.no_grad():
        bands = fe._bands(x)
        logits = fe(x)
    assert_close("frontend/bands", bands[:, :, :8], torch.from_numpy(fx["bands_out"]), rtol=1e-5, atol=2e-6)
    assert_close("frontend/logits", logits, torch.from_numpy(fx["logits"]), rtol=1e-4, atol=1e-5)


def test_dataset_sampling_follows_the_reference_index_arithmetic(tmp_path):
    """nsyth_dataset.py:50-69 on a neighbors.pkl in compute_distances' schema, full lists and top-k lists."""
    from topo_audio_autoencoder_b200.nsyth_dataset import NSynthDataset
    n = 300
    keys = [f"clip{i:03d}" for i in range(n)]
    root = tmp_path / "tensors"
    root.mkdir()
    for i, k in enumerate(keys):
        torch.save(torch.full((1, 8), float(i)), root / f"{k}.pt")
    g = torch.Generator().manual_seed(0)
    full = {}
    for i, k in enumerate(keys):
        order = [keys[j] for j in torch.randperm(n, generator=g).tolist() if j != i]
        full[k] = {"sorted_neighbors": order, "sorted_distances": sorted(torch.rand(n - 1, generator=g).tolist()), "index": i}
    full["__file_to_idx__"] = {k: i for i, k in enumerate(keys)}
    with open(tmp_path / "full.pkl", "wb") as f:
        pickle.dump(full, f)
    data = {k: {} for k in keys[:250]}                     # a training split smaller than the neighbour lists
    ds = NSynthDataset(data, str(root), train=True, neighbors_path=str(tmp_path / "full.pkl"))
    assert ds.current_negative_offset == 250
    item = ds[3]
    assert item.shape == (12, 1, 8)
    order = full["clip003"]["sorted_neighbors"]
    idx = lambda name: float(name[4:])                     # noqa: E731
    assert item[0, 0, 0] == 3.0 and item[1, 0, 0].item() in {idx(nm) for nm in order[:10]}
    assert [v.item() for v in item[2:, 0, 0]] == [idx(nm) for nm in order[240:250]]           # range(offset - 10, offset)
    ds.set_epoch(5)
    assert ds.current_negative_offset == int(250 * 0.9 ** 5)
    ds.set_epoch(50)
    assert ds.current_negative_offset == 100 and ds.negative_window("clip003") == order[90:100]
    ds.train = False
    assert ds[7].shape == (1, 8)
    # top-k file: positives unchanged, negatives from the far end of what the file holds
    top = {k: (dict(v, sorted_neighbors=v["sorted_neighbors"][:32], sorted_distances=v["sorted_distances"][:32])
               if k != "__file_to_idx__" else v) for k, v in full.items()}
    with open(tmp_path / "top.pkl", "wb") as f:
        pickle.dump(top, f)
    ds2 = NSynthDataset(data, str(root), train=True, neighbors_path=str(tmp_path / "top.pkl"))
    assert ds2.negative_window("clip003") == order[22:32]


def test_spectral_loss_matches_the_pair_distance_of_the_sweep():
    """loss.spectral_distance on single-clip batches equals the distance oracle's per-pair value (same formula, the
    batch mean of one element), and AutoencoderLoss adds the weighted penalties (loss.py:39-43)."""
    from oracle import distance_oracle as do
    from topo_audio_autoencoder_b200.loss import AutoencoderLoss, spectral_distance
    g = torch.Generator().manual_seed(2)
    x, y = torch.randn(1, 1, 8192, generator=g) * 0.1, torch.randn(1, 1, 8192, generator=g) * 0.1
    want = do.batch_audio_distance(x, y)[0]
    assert_close("loss/spectral", spectral_distance(x, y).reshape(1), want.reshape(1), rtol=1e-5, atol=1e-6)
    fn = AutoencoderLoss(binary_entropy_penalty=0.01, complexity_penalty=0.1)
    total = fn(x, y, {"binary_entropy": torch.tensor(-0.05), "diversity": torch.tensor(3.0)})
    assert_close("loss/total", total.reshape(1), (want + 0.01 * -0.05 + 0.1 * 3.0).reshape(1), rtol=1e-5, atol=1e-6)
    assert set(fn.loss_components) == {"spectral_loss", "binary_entropy_loss", "diversity_loss", "total_loss"}
    xg = x.clone().requires_grad_(True)
    fn(xg, y, {"binary_entropy": 0.0, "diversity": 0.0}).backward()
    assert torch.isfinite(xg.grad).all() and xg.grad.abs().sum() > 0
