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
    assert torch.equal(full, full.t())
