"""GPU parity: tiled pairwise spectral distance vs the oracle restatement of precompute_distances.py.
STFT parity is unpinned (rave absent); the pair reduction after the STFT is what is pinned."""
import pytest
import torch

from oracle import distance_oracle as do
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu


def test_pair_reduction_against_oracle_small():
    from topo_audio_autoencoder_b200 import precompute_distances as pd
    g = torch.Generator().manual_seed(511990)
    audio = torch.randn(11, 1, 6000, generator=g) * 0.1
    want = do.pairwise_matrix(audio, batch_size=4)
    got = pd.pairwise_spectral_distances(audio.cuda())
    assert_close("distance/matrix-11x6000", got, want, rtol=2e-5, atol=1e-6)
    assert torch.equal(got, got.t()), "mirrored exactly (precompute_distances.py:114-115)"
    assert (torch.diagonal(got) == 0).all()
    v_c, i_c = do.neighbour_order(want)
    v_g, i_g = pd.neighbour_order(got.cpu())
    assert torch.equal(i_c, i_g), "neighbour ordering"
    # BatchAudioDistance: first argument supplies the normaliser
    bad = pd.BatchAudioDistance()
    d = bad(audio[:4].cuda(), audio[4:8].cuda())["spectral_distance"]
    assert_close("distance/batch-audio-distance", d, do.batch_audio_distance(audio[:4], audio[4:8]), rtol=2e-5, atol=1e-6)


def test_full_length_clips_and_row_sharding():
    """4 s / 16 kHz clips (D = 645,864 bins): a few pairs against the oracle, and row-block shards
    reassemble to the single-rank matrix bit for bit."""
    from topo_audio_autoencoder_b200 import precompute_distances as pd
    g = torch.Generator().manual_seed(1)
    audio = torch.randn(70, 1, 64000, generator=g) * 0.1
    spec, seg = pd.multiscale_spectrograms(audio.cuda())
    assert spec.shape[1] == 645864 and seg == [129150, 128763, 128757, 129129, 130065]
    prep = pd.PreparedSpectra(spec, seg)
    full = prep.rows(0, 70)
    want = do.batch_audio_distance(audio[[0, 3, 65]], audio[[1, 69, 68]])
    got = torch.stack([full[0, 1], full[3, 69], full[65, 68]])
    assert_close("distance/full-length-pairs", got, want, rtol=2e-5, atol=1e-6)
    shards = [prep.rows(*pd.shard_rows(70, r, 4)) for r in range(4)]
    assert torch.equal(torch.cat(shards), full)
    # both orientations of a pair are reduced; they agree to fp32 rounding and the dense product mirrors the upper one
    assert_close("distance/orientation", full, full.t(), rtol=2e-6, atol=1e-6)
    sym = pd.mirror_upper(full)
    assert torch.equal(sym, sym.t()) and torch.equal(torch.triu(sym, 1), torch.triu(full, 1)) and (torch.diagonal(full) == 0).all()


def test_streaming_topk_sweep_equals_the_dense_matrix():
    """Config 5 machinery at a size the dense path can check: resident row blocks against column blocks prepared on the
    fly (audio in pinned host memory), running top-k merge, row-sharded; nothing of size N x N is formed."""
    from topo_audio_autoencoder_b200 import precompute_distances as pd
    g = torch.Generator().manual_seed(9)
    n, k = 37, 6
    audio = torch.randn(n, 1, 6000, generator=g) * 0.1
    dense = pd.pairwise_spectral_distances(audio.cuda())
    want_v, want_i = pd.neighbour_order(dense)
    for cache_bytes in (0, 48 << 30):             # columns re-prepared per row block / prepared once
        vals, idx = pd.spectral_topk(audio.pin_memory(), k, row_block=16, col_block=8, cache_bytes=cache_bytes)
        # the pair reduction is identical; the spectra come from cuFFT plans of different batch sizes (a block of clips
        # instead of the whole collection), which differ in the last bits
        assert_close("distance/topk-values", vals, want_v[:, :k], rtol=2e-6, atol=1e-6)
        assert torch.equal(idx, want_i[:, :k])
    shards = [pd.spectral_topk(audio.cuda(), k, row_block=5, col_block=16, rank=r, world_size=3) for r in range(3)]
    assert_close("distance/topk-sharded", torch.cat([s[0] for s in shards]), want_v[:, :k], rtol=2e-6, atol=1e-6)
    assert torch.equal(torch.cat([s[1] for s in shards]), want_i[:, :k])
    # asymmetric normaliser across block borders: entry (i, j) and (j, i) of different blocks agree exactly
    rows, cols = pd.prepare_block(audio[3:9], "cuda"), pd.prepare_block(audio[20:31], "cuda")
    a, b = rows.block(cols, 3, 20), cols.block(rows, 20, 3)
    assert torch.equal(a, b.t())
    assert_close("distance/block-vs-dense", a, dense[3:9, 20:31], rtol=2e-6, atol=1e-6)
    # the same prepared spectra through both entry points: bit-identical (tile position does not change a pair's arithmetic)
    spec, seg = pd.multiscale_spectrograms(audio.cuda())
    whole = pd.PreparedSpectra(spec, seg)
    part_r, part_c = pd.PreparedSpectra(spec[3:9].contiguous(), seg), pd.PreparedSpectra(spec[20:31].contiguous(), seg)
    assert torch.equal(part_r.block(part_c, 3, 20), whole.rows(3, 9, 20, 31))


def test_compute_distances_top_k_files(tmp_path):
    import pickle
    import wave
    import numpy as np
    from topo_audio_autoencoder_b200 import precompute_distances as pd
    g = torch.Generator().manual_seed(4)
    adir, sdir = tmp_path / "a", tmp_path / "o"
    adir.mkdir(); sdir.mkdir()
    for j in range(9):
        x = (torch.randn(5000 + 300 * j, generator=g) * 0.1).clamp(-1, 1)
        with wave.open(str(adir / f"c{j}.wav"), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
            w.writeframes(np.round(x.numpy() * 32767).astype(np.int16).tobytes())
    dense = pd.compute_distances(adir, sdir)
    with open(sdir / "neighbors.pkl", "rb") as f:
        full = pickle.load(f)
    res = pd.compute_distances(adir, sdir, top_k=3)
    with open(sdir / "neighbors.pkl", "rb") as f:
        top = pickle.load(f)
    assert (sdir / "topk_distances.pt").exists() and tuple(res["indices"].shape) == (9, 3)
    assert top["__file_to_idx__"] == full["__file_to_idx__"]
    for name, rec in full.items():
        if name == "__file_to_idx__":
            continue
        assert top[name]["sorted_neighbors"] == rec["sorted_neighbors"][:3]
        assert top[name]["sorted_distances"] == rec["sorted_distances"][:3]
        assert top[name]["index"] == rec["index"]
    assert torch.equal(torch.load(sdir / "distance_matrix.pt"), dense)


def test_gram_term_keeps_fp32_accuracy_for_near_duplicates():
    """The squared-difference term is mean x^2 + mean y^2 - 2 <x, y> / len with <x, y> on the tensor cores: duplicates and
    near-duplicates (the pairs a nearest-neighbour search is about) are where that form cancels."""
    from topo_audio_autoencoder_b200 import precompute_distances as pd
    g = torch.Generator().manual_seed(21)
    base = torch.randn(6, 1, 64000, generator=g) * 0.1
    audio = torch.cat([base, base[:2].clone(), base[2:4] * 1.0005, base[4:6] + 1e-4 * torch.randn(2, 1, 64000, generator=g)])
    got = pd.pairwise_spectral_distances(audio.cuda()).cpu()
    pairs = [(0, 6), (1, 7), (2, 8), (3, 9), (4, 10), (5, 11), (0, 1), (6, 9)]
    want = do.batch_audio_distance(audio[[a for a, _ in pairs]], audio[[b for _, b in pairs]])
    mine = torch.stack([got[a, b] for a, b in pairs])
    # exact duplicates: the true distance is 0; the two ways of summing x^2 (prepare kernel vs tensor cores) leave ~1e-6
    assert mine[:2].abs().max().item() < 5e-6, mine[:2]
    assert_close("distance/near-duplicates", mine[2:], want[2:], rtol=2e-5, atol=3e-6)
    assert (got >= 0).all()


def test_edge_sizes_of_the_sweep():
    """one clip, two clips, k larger than the collection, block sizes that do not divide anything, short clips"""
    from topo_audio_autoencoder_b200 import precompute_distances as pd
    g = torch.Generator().manual_seed(5)
    one = torch.randn(1, 1, 3000, generator=g) * 0.1
    d1 = pd.pairwise_spectral_distances(one.cuda())
    assert tuple(d1.shape) == (1, 1) and d1.item() == 0.0
    v, i = pd.spectral_topk(one.cuda(), 4)
    assert tuple(v.shape) == (1, 0) and tuple(i.shape) == (1, 0)
    audio = torch.randn(131, 1, 2500, generator=g) * 0.1          # 131 clips: two 128-clip image blocks, the second nearly empty
    dense = pd.pairwise_spectral_distances(audio.cuda())
    want = do.pairwise_matrix(audio[:9], batch_size=4)
    assert_close("distance/edge/131-clips", dense[:9, :9], want, rtol=2e-5, atol=2e-6)
    assert torch.isfinite(dense).all() and torch.equal(dense, dense.t())
    v, i = pd.spectral_topk(audio.cuda(), 500, row_block=50, col_block=77)      # k clamps to n - 1
    assert tuple(v.shape) == (131, 130)
    wv, wi = pd.neighbour_order(dense)
    assert_close("distance/edge/full-order", v, wv, rtol=2e-6, atol=1e-6)
    assert (torch.sort(i, dim=1).values == torch.sort(wi, dim=1).values).all(), "every other clip exactly once"
    two = audio[:2].cuda()
    d2 = pd.pairwise_spectral_distances(two)
    assert_close("distance/edge/two", d2[0, 1].reshape(1), do.batch_audio_distance(audio[:1], audio[1:2]), rtol=2e-5, atol=2e-6)


def test_fixed_point_logs_cover_silence_and_loud_clips():
    """The L1 term runs on Q6.20 fixed-point logs (window [-40, 24) in log units, csrc/distance.cu): digital silence
    (log(1e-7) = -16.1 in every bin), very quiet, ordinary and very loud clips against the fp32 oracle, and the L1 term on
    its own against an fp64 sum of the same fp32 logs."""
    from topo_audio_autoencoder_b200 import precompute_distances as pd
    g = torch.Generator().manual_seed(77)
    base = torch.randn(4, 1, 16000, generator=g) * 0.1
    audio = torch.cat([torch.zeros(1, 1, 16000), base[:1] * 1e-4, base[1:2], base[2:3] * 1e3, base[3:4] * 3e4])
    got = pd.pairwise_spectral_distances(audio.cuda()).cpu()
    want = do.pairwise_matrix(audio, batch_size=4)
    assert_close("distance/fixed-point-window", got, want, rtol=2e-5, atol=2e-6)
    # the L1 term alone: d(i, j) - relative-L2 term, both sides from the same spectra
    spec, seg = pd.multiscale_spectrograms(audio.cuda())
    logs = torch.log(spec.double() + pd.LOG_EPSILON)
    assert logs.min().item() > -40 and logs.max().item() < 24
    off = 0
    l1 = torch.zeros(5, 5, dtype=torch.float64, device="cuda")
    l2 = torch.zeros(5, 5, dtype=torch.float64, device="cuda")
    for n_s in seg:
        x, lx = spec[:, off:off + n_s].double(), logs[:, off:off + n_s].float().double()
        l1 += (lx[:, None, :] - lx[None, :, :]).abs().mean(-1)
        msq = ((x[:, None, :] - x[None, :, :]) ** 2).mean(-1)
        norm = (x ** 2).mean(-1)
        idx = torch.arange(5, device="cuda")
        l2 += msq / (torch.where(idx[:, None] <= idx[None, :], norm[:, None], norm[None, :]) + 1e-7)      # the lower-index clip's
        off += n_s
    ref = (l1 + l2).float().cpu()
    ref.fill_diagonal_(0)
    assert_close("distance/fixed-point-vs-fp64", got, ref, rtol=5e-6, atol=2e-6)
    with pytest.raises(RuntimeError, match="fixed-point window"):
        pd.PreparedSpectra(spec, seg, log_eps=1e-20)
