"""Oracle: pairwise multi-scale spectral distance.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  * batch_mean_difference   precompute_distances.py:11-31
  * batch_audio_distance    precompute_distances.py:33-49 (BatchAudioDistance.forward)
  * pairwise_matrix         precompute_distances.py:89-115 (upper-triangle sweep; the row clip is
                            always the first argument, so the asymmetric normaliser comes from
                            the lower-index clip; result mirrored)
  * neighbour_order         precompute_distances.py:121-126 (sort rows, drop self)

``multiscale_stft`` -- STFT PARITY UNPINNED.  ``MultiScaleSTFT`` is acids-rave
(``rave.core``, absent, no version pinned).  Restated from its published behaviour: one
``torchaudio.transforms.Spectrogram(n_fft=s, win_length=s, hop_length=s//4, power=None)`` per
scale (Hann window, centred, reflect padding), then ``abs()`` because the reference asks for
``magnitude=True`` (precompute_distances.py:65).  Only the scale list and epsilon are in-repo.
"""
from __future__ import annotations

import torch

SCALES = (2048, 1024, 512, 256, 128)   # precompute_distances.py:65
LOG_EPS = 1e-7                         # precompute_distances.py:66


def multiscale_stft(x: torch.Tensor, scales=SCALES):
    """x [B, 1, T] -> list of magnitude spectrograms [B, s/2+1, frames]."""
    x = x.reshape(-1, x.shape[-1])
    out = []
    for s in scales:
        win = torch.hann_window(s, dtype=x.dtype, device=x.device)
        spec = torch.stft(x, n_fft=s, hop_length=s // 4, win_length=s, window=win, center=True,
                          pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
        out.append(spec.abs())
    return out


def batch_mean_difference(target, value, norm="L1", relative=False):
    """precompute_distances.py:11-31."""
    dims = list(range(1, target.dim()))
    d = target - value
    if norm == "L1":
        d = d.abs().mean(dim=dims)
        ref = target.abs().mean(dim=dims)
    elif norm == "L2":
        d = (d * d).mean(dim=dims)
        ref = (target * target).mean(dim=dims)
    else:
        raise ValueError(f"Norm must be either L1 or L2, got {norm}")
    return d / (ref + 1e-7) if relative else d


def spectral_distance(specs_x, specs_y, log_eps=LOG_EPS):
    """precompute_distances.py:39-49 on precomputed spectrogram lists."""
    total = 0.0
    for sx, sy in zip(specs_x, specs_y):
        lin = batch_mean_difference(sx, sy, norm="L2", relative=True)
        log = batch_mean_difference(torch.log(sx + log_eps), torch.log(sy + log_eps), norm="L1")
        total = total + lin + log
    return total


def batch_audio_distance(x, y, scales=SCALES):
    """precompute_distances.py:33-49: x, y [B, 1, T] -> [B]."""
    return spectral_distance(multiscale_stft(x, scales), multiscale_stft(y, scales))


def pairwise_matrix(audio, batch_size=32, scales=SCALES):
    """precompute_distances.py:64-115 without file I/O: audio [N, 1, T] -> [N, N]."""
    n = audio.shape[0]
    dist = torch.zeros(n, n)
    rows, cols = torch.triu_indices(n, n, offset=1)
    for b in range(0, len(rows), batch_size):
        r, c = rows[b:b + batch_size], cols[b:b + batch_size]
        d = batch_audio_distance(audio[r], audio[c], scales)
        dist[r, c] = d
        dist[c, r] = d
    return dist


def neighbour_order(dist):
    """precompute_distances.py:121-126."""
    vals, idx = torch.sort(dist, dim=1)
    return vals[:, 1:], idx[:, 1:]
