"""Drop-in for the reference's ``precompute_distances.py``.

    batch_mean_difference(target, value, norm, relative)      reference :11-31
    BatchAudioDistance.forward(x, y) -> {'spectral_distance'} reference :33-49
    compute_distances(audio_dir, save_path, batch_size=32)    reference :51-153

The reference recomputes both multi-scale STFTs for every one of the N(N-1)/2 pairs and fills the
matrix in a Python loop.  Here the N spectrograms are computed once (cuFFT through torch.stft --
the front half, a library call) and the pair reduction runs in csrc/distance.cu, tiled, for a block
of rows at a time, so ranks can shard the matrix by row blocks without any collective.
"""
from __future__ import annotations

import pickle
from pathlib import Path
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from ._lib import lib, check, ptr, stream, i64_array

SCALES = (2048, 1024, 512, 256, 128)     # reference :65
LOG_EPSILON = 1e-7                       # reference :66


def batch_mean_difference(target: torch.Tensor, value: torch.Tensor, norm: str = "L1", relative: bool = False):
    """reference :11-31 (any shape; kept as tensor algebra -- the hot pair loop does not go through it)."""
    dims = list(range(1, target.dim()))
    diff = target - value
    if norm == "L1":
        diff = diff.abs().mean(dim=dims)
        ref = target.abs().mean(dim=dims)
    elif norm == "L2":
        diff = (diff * diff).mean(dim=dims)
        ref = (target * target).mean(dim=dims)
    else:
        raise ValueError(f"Norm must be either L1 or L2, got {norm}")
    return diff / (ref + 1e-7) if relative else diff


def multiscale_spectrograms(audio: torch.Tensor, scales: Sequence[int] = SCALES):
    """audio [N, 1, T] or [N, T] (CUDA) -> (spec [N, D] with the scales' magnitude spectrograms
    flattened and concatenated, segment lengths).  Same transform as rave's MultiScaleSTFT with
    magnitude=True: Hann window, hop = scale // 4, centred, reflect padding."""
    x = audio.reshape(audio.shape[0], -1)
    parts, seg = [], []
    for s in scales:
        win = torch.hann_window(s, dtype=x.dtype, device=x.device)
        mag = torch.stft(x, n_fft=s, hop_length=s // 4, win_length=s, window=win, center=True, pad_mode="reflect",
                         normalized=False, onesided=True, return_complex=True).abs()
        parts.append(mag.reshape(x.shape[0], -1))
        seg.append(parts[-1].shape[1])
    return torch.cat(parts, dim=1).contiguous(), seg


class PreparedSpectra:
    """Padded spectrogram rows, their logs and per-scale mean squares, resident on the device."""

    def __init__(self, spec: torch.Tensor, seg_len: Sequence[int], log_eps: float = LOG_EPSILON):
        spec = spec.contiguous()
        self.n, self.d = spec.shape
        self.seg_len = [int(s) for s in seg_len]
        self.seg_c = i64_array(self.seg_len)
        self.dp = int(lib.topo_distance_padded_size(self.seg_c, len(self.seg_len)))
        if self.dp < 0:
            raise ValueError("between 1 and 8 scales are supported")
        dev = spec.device
        self.spec_p = torch.empty(self.n, self.dp, dtype=torch.float32, device=dev)
        self.logspec_p = torch.empty(self.n, self.dp, dtype=torch.float32, device=dev)
        self.sq_mean = torch.empty(self.n, len(self.seg_len), dtype=torch.float32, device=dev)
        check(lib.topo_distance_prepare(ptr(spec), self.n, self.d, self.seg_c, len(self.seg_len), float(log_eps),
                                        ptr(self.spec_p), ptr(self.logspec_p), ptr(self.sq_mean), stream()))

    def rows(self, row_begin: int, row_end: int, col_begin: int = 0, col_end: Optional[int] = None) -> torch.Tensor:
        col_end = self.n if col_end is None else col_end
        out = torch.empty(row_end - row_begin, col_end - col_begin, dtype=torch.float32, device=self.spec_p.device)
        check(lib.topo_distance_rows(ptr(self.spec_p), ptr(self.logspec_p), ptr(self.sq_mean), self.n, self.seg_c,
                                     len(self.seg_len), row_begin, row_end, col_begin, col_end, ptr(out), stream()))
        return out


def pairwise_spectral_distances(audio: torch.Tensor, scales: Sequence[int] = SCALES, rank: int = 0,
                                world_size: int = 1) -> torch.Tensor:
    """The rows of the N x N distance matrix owned by `rank` (contiguous row blocks), all columns."""
    spec, seg = multiscale_spectrograms(audio, scales)
    prep = PreparedSpectra(spec, seg)
    lo, hi = shard_rows(prep.n, rank, world_size)
    return prep.rows(lo, hi)


def shard_rows(n: int, rank: int, world_size: int):
    per = (n + world_size - 1) // world_size
    return min(rank * per, n), min((rank + 1) * per, n)


class BatchAudioDistance(nn.Module):
    """reference :33-49: x, y [B, 1, T] -> {'spectral_distance': [B]} (x supplies the normaliser)."""

    def __init__(self, scales: Sequence[int] = SCALES, log_epsilon: float = LOG_EPSILON):
        super().__init__()
        self.scales, self.log_epsilon = tuple(scales), log_epsilon

    @torch.no_grad()
    def forward(self, x: torch.Tensor, y: torch.Tensor):
        b = x.shape[0]
        spec, seg = multiscale_spectrograms(torch.cat([x, y], dim=0), self.scales)
        prep = PreparedSpectra(spec, seg, self.log_epsilon)
        block = prep.rows(0, b, b, 2 * b)          # rows = x clips (lower index), columns = y clips
        return {"spectral_distance": torch.diagonal(block).clone()}


def load_wav(path) -> torch.Tensor:
    """One audio file -> float32 [channels, T] in [-1, 1), as ``torchaudio.load`` returns it (reference :77).  torchaudio
    needs an audio backend (torchcodec / soundfile) that is not always installed; 16-bit and 32-bit PCM and float32 wav
    files are then read with the standard library instead."""
    try:
        import torchaudio
        wav, _ = torchaudio.load(path)
        return wav
    except (ImportError, RuntimeError, OSError):
        pass
    import wave
    import numpy as np
    with wave.open(str(path), "rb") as w:
        ch, width, frames = w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(frames)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    else:
        raise ValueError(f"{path}: unsupported sample width {width} bytes")
    return torch.from_numpy(x.reshape(-1, ch).T.copy())


def neighbour_order(distances: torch.Tensor):
    """reference :121-126: full ascending order of every row, self dropped."""
    vals, idx = torch.sort(distances, dim=1)
    return vals[:, 1:], idx[:, 1:]


def compute_distances(audio_dir: Path, save_path: Path, batch_size: int = 32):
    """reference :51-153.  Same inputs and the same two output files (distance_matrix.pt,
    neighbors.pkl with 'sorted_neighbors' / 'sorted_distances' / 'index' per file and the
    '__file_to_idx__' map); ``batch_size`` is accepted for signature parity and unused."""
    audio_files = list(Path(audio_dir).glob("*.wav"))
    n_files = len(audio_files)
    file_to_idx = {str(f): i for i, f in enumerate(audio_files)}
    wavs, max_len = [], 0
    for f in audio_files:
        wav = load_wav(f)
        wavs.append(wav.unsqueeze(0))
        max_len = max(max_len, wav.shape[-1])
    audio = torch.cat([torch.nn.functional.pad(w, (0, max_len - w.shape[2])) for w in wavs], dim=0).cuda()
    distances = pairwise_spectral_distances(audio[:, :1]).cpu()
    sorted_vals, sorted_idx = neighbour_order(distances)
    neighbors = {
        str(audio_files[i]): {
            "sorted_neighbors": [str(audio_files[j]) for j in sorted_idx[i].tolist()],
            "sorted_distances": sorted_vals[i].tolist(),
            "index": i,
        } for i in range(n_files)
    }
    neighbors["__file_to_idx__"] = file_to_idx
    save_path = Path(save_path)
    torch.save(distances, save_path / "distance_matrix.pt")
    with open(save_path / "neighbors.pkl", "wb") as f:
        pickle.dump(neighbors, f)
    return distances
