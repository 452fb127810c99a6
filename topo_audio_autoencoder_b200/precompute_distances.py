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
    """Operands of the pair reduction for one block of clips, resident on the device: the logs as Q6.20 fixed point, k-major
    inside 64-clip blocks (integer pipe, L1 term), the bf16x3 operand image of the spectra (tensor cores, Gram term;
    csrc/distance.cu) and per-scale mean squares."""

    def __init__(self, spec: torch.Tensor, seg_len: Sequence[int], log_eps: float = LOG_EPSILON):
        spec = spec.contiguous()
        self.n, self.d = spec.shape
        self.seg_len = [int(s) for s in seg_len]
        self.seg_c = i64_array(self.seg_len)
        n_scales = len(self.seg_len)
        self.dp = int(lib.topo_distance_padded_size(self.seg_c, n_scales))
        if self.dp < 0:
            raise ValueError("between 1 and 8 scales are supported")
        dev = self.device = spec.device
        self.logq = torch.empty(int(lib.topo_distance_logq_words(self.n, self.seg_c, n_scales)), dtype=torch.int32, device=dev)
        image_bytes = int(lib.topo_distance_image_bytes(self.n, self.seg_c, n_scales))
        # rows past n in the last 128-clip block must read as zero; cudaMalloc'ed tensors are at least 512-byte aligned and
        # the caching allocator hands out 512-byte multiples: over-allocate to reach the 1024-byte alignment of the image
        raw = (torch.zeros if self.n % 128 else torch.empty)(image_bytes + 1024, dtype=torch.uint8, device=dev)
        shift = (-raw.data_ptr()) % 1024
        self._image_raw = raw
        self.image = raw[shift:shift + image_bytes]
        self.sq_mean = torch.empty(self.n, n_scales, dtype=torch.float32, device=dev)
        check(lib.topo_distance_prepare(ptr(spec), self.n, self.d, self.seg_c, n_scales, float(log_eps),
                                        self.logq.data_ptr(), self.image.data_ptr(), ptr(self.sq_mean), stream()))

    def _workspace(self, n_rows: int, n_cols: int) -> torch.Tensor:
        words = int(lib.topo_distance_workspace_floats(n_rows, n_cols, len(self.seg_len)))
        return torch.empty(max(words, 1), dtype=torch.float32, device=self.device)

    def block(self, cols: "PreparedSpectra", row_global0: int = 0, col_global0: int = 0,
              out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
        """All pairs (row clip of self, column clip of `cols`) -> [self.n, cols.n].  The two blocks are slices
        [row_global0, ...) and [col_global0, ...) of one collection: the collection-wide indices decide which clip of a
        pair supplies the normaliser (the lower one, reference :89, :106-110) and where the zero diagonal is."""
        if cols.seg_len != self.seg_len:
            raise ValueError("row and column blocks were prepared with different scale segments")
        if out is None:
            out = torch.empty(self.n, cols.n, dtype=torch.float32, device=self.device)
        # `out` may be a column window of a wider buffer (the running top-k merge): unit column stride, any row stride
        if not (out.is_cuda and out.dtype == torch.float32 and out.dim() == 2 and out.stride(1) == 1
                and out.shape[0] >= self.n and out.shape[1] >= cols.n):
            raise ValueError("out must be a CUDA float32 [rows, >= columns] view with unit column stride")
        if workspace is None:
            workspace = self._workspace(self.n, cols.n)
        check(lib.topo_distance_block(self.logq.data_ptr(), self.image.data_ptr(), ptr(self.sq_mean), self.n, int(row_global0),
                                      cols.logq.data_ptr(), cols.image.data_ptr(), ptr(cols.sq_mean), cols.n, int(col_global0),
                                      self.seg_c, len(self.seg_len), ptr(workspace), out.data_ptr(), out.stride(0), stream()))
        return out

    def rows(self, row_begin: int, row_end: int, col_begin: int = 0, col_end: Optional[int] = None) -> torch.Tensor:
        col_end = self.n if col_end is None else col_end
        out = torch.empty(row_end - row_begin, col_end - col_begin, dtype=torch.float32, device=self.device)
        workspace = self._workspace(row_end - row_begin, col_end - col_begin)
        check(lib.topo_distance_rows(self.logq.data_ptr(), self.image.data_ptr(), ptr(self.sq_mean), self.n, self.seg_c,
                                     len(self.seg_len), row_begin, row_end, col_begin, col_end, ptr(workspace), ptr(out), stream()))
        return out


def pairwise_spectral_distances(audio: torch.Tensor, scales: Sequence[int] = SCALES, rank: int = 0,
                                world_size: int = 1) -> torch.Tensor:
    """The rows of the N x N distance matrix owned by `rank` (contiguous row blocks), all columns."""
    spec, seg = multiscale_spectrograms(audio, scales)
    prep = PreparedSpectra(spec, seg)
    lo, hi = shard_rows(prep.n, rank, world_size)
    out = prep.rows(lo, hi)
    if world_size == 1:
        out = mirror_upper(out)
    return out


def mirror_upper(distances: torch.Tensor) -> torch.Tensor:
    """reference :113-115: only the pairs i < j are computed (x = the lower-index clip) and mirrored.  The kernels reduce
    both orientations of a pair; they agree to fp32 rounding (the six bf16 part products of the Gram term are accumulated
    in an order that depends on which clip is the row), so the exact mirror of the upper triangle is taken here.  A
    row-sharded rank holds only its own rows and keeps both orientations as computed."""
    upper = torch.triu(distances, diagonal=1)
    return upper + upper.t()


def shard_rows(n: int, rank: int, world_size: int):
    per = (n + world_size - 1) // world_size
    return min(rank * per, n), min((rank + 1) * per, n)


def prepare_block(audio_block: torch.Tensor, device, scales: Sequence[int] = SCALES) -> PreparedSpectra:
    """Front half for one block of clips (host or device tensor [n, 1, T] / [n, T]): H2D copy if needed, multi-scale
    STFT (cuFFT through torch.stft -- a library call), padded spectra + logs + mean squares (csrc/distance.cu)."""
    x = audio_block.to(device, non_blocking=True)
    spec, seg = multiscale_spectrograms(x, scales)
    return PreparedSpectra(spec, seg)


@torch.no_grad()
def spectral_topk(audio: torch.Tensor, k: int, row_block: int = 1024, col_block: int = 4096, rank: int = 0,
                  world_size: int = 1, device=None, scales: Sequence[int] = SCALES, cache_bytes: int = 48 << 30):
    """Streaming sweep for collections whose spectra do not fit in HBM (config 5: 100k clips x 2.58 MB x 2 arrays).

    The rows owned by `rank` (contiguous row blocks, no collective on the compute path) are processed `row_block` clips
    at a time: their spectra stay resident while every `col_block` of the collection is prepared on the fly from the
    audio -- which may live in (pinned) host memory -- and reduced against them; only the k nearest neighbours of each
    row survive a block (running top-k merge), so nothing of size N x N is ever formed.  When all column spectra fit in
    `cache_bytes` they are prepared once instead.

    Launch geometry: one launch reduces row_block x col_block pairs in 128 x 64 (L1 term) and 128 x 128 x 5 scales (Gram
    term) tiles, so the blocks are chosen large enough for several CTAs per SM (1024 x 4096: 512 and 1280 CTAs on 148 SMs);
    a row block of 1024 clips keeps 6.7 GB resident, a column block of 4096 clips 27 GB.

    -> (distances [rows, k] ascending, neighbour indices [rows, k] int64, collection-wide; self excluded,
        reference :121-126), on `device`."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = audio.shape[0]
    k = min(int(k), max(n - 1, 0))
    lo, hi = shard_rows(n, rank, world_size)
    vals = torch.empty(hi - lo, k, dtype=torch.float32, device=device)
    idx = torch.empty(hi - lo, k, dtype=torch.int64, device=device)
    if hi == lo or k == 0:
        return vals, idx
    col_starts = list(range(0, n, col_block))
    probe = prepare_block(audio[:1], device, scales)
    per_clip = probe.dp * 10                          # fp32 logs + three bf16 parts
    cached = None
    if n * per_clip <= cache_bytes:
        cached = [prepare_block(audio[c0:c0 + col_block], device, scales) for c0 in col_starts]
    scratch = torch.empty(min(row_block, hi - lo), k + col_block, dtype=torch.float32, device=device)
    scratch_i = torch.empty(min(row_block, hi - lo), k + col_block, dtype=torch.int64, device=device)
    for r0 in range(lo, hi, row_block):
        r1 = min(r0 + row_block, hi)
        rows = prepare_block(audio[r0:r1], device, scales)
        m = r1 - r0
        best_v = scratch[:m, :k].fill_(float("inf"))
        best_i = scratch_i[:m, :k].fill_(-1)
        row_ids = torch.arange(r0, r1, device=device).unsqueeze(1)
        for ci, c0 in enumerate(col_starts):
            cols = cached[ci] if cached is not None else prepare_block(audio[c0:c0 + col_block], device, scales)
            width = cols.n
            d = scratch[:m, k:k + width]
            rows.block(cols, r0, c0, out=d)
            col_ids = torch.arange(c0, c0 + width, device=device).unsqueeze(0)
            d.masked_fill_(row_ids == col_ids, float("inf"))                     # drop self (reference :124-126)
            scratch_i[:m, k:k + width] = col_ids
            v, sel = torch.topk(scratch[:m, :k + width], k, dim=1, largest=False, sorted=True)
            i_sel = torch.gather(scratch_i[:m, :k + width], 1, sel)
            best_v.copy_(v)
            best_i.copy_(i_sel)
        vals[r0 - lo:r1 - lo] = best_v
        idx[r0 - lo:r1 - lo] = best_i
    return vals, idx


def gather_topk(vals: torch.Tensor, idx: torch.Tensor, n: int):
    """The ONE exchange of the sharded sweep: every rank's [rows, k] results -> [n, k] on every rank (NCCL all-gather of
    equally padded shards; a no-op without an initialised process group)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return vals, idx
    world = dist.get_world_size()
    per = (n + world - 1) // world
    pv = torch.full((per, vals.shape[1]), float("inf"), dtype=vals.dtype, device=vals.device)
    pi = torch.full((per, idx.shape[1]), -1, dtype=idx.dtype, device=idx.device)
    pv[:vals.shape[0]], pi[:idx.shape[0]] = vals, idx
    gv = torch.empty(world * per, vals.shape[1], dtype=vals.dtype, device=vals.device)
    gi = torch.empty(world * per, idx.shape[1], dtype=idx.dtype, device=idx.device)
    dist.all_gather_into_tensor(gv, pv)
    dist.all_gather_into_tensor(gi, pi)
    return gv[:n], gi[:n]


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


def _load_padded(audio_files) -> torch.Tensor:
    """reference :72-85: load every file, zero-pad to the longest, stack -> [N, 1, T] (host)."""
    wavs, max_len = [], 0
    for f in audio_files:
        wav = load_wav(f)
        wavs.append(wav.unsqueeze(0))
        max_len = max(max_len, wav.shape[-1])
    return torch.cat([torch.nn.functional.pad(w, (0, max_len - w.shape[2])) for w in wavs], dim=0)[:, :1]


def compute_distances(audio_dir: Path, save_path: Path, batch_size: int = 32, top_k: Optional[int] = None,
                      rank: int = 0, world_size: int = 1):
    """reference :51-153.  Same inputs and, by default, the same two output files (distance_matrix.pt: dense [N, N]
    fp32; neighbors.pkl: per file 'sorted_neighbors' / 'sorted_distances' / 'index', plus the '__file_to_idx__' map);
    ``batch_size`` is accepted for signature parity and unused.

    ``top_k`` (this repo's addition for collections where a dense matrix and fully sorted rows are impractical -- 40 GB
    and 10^10 list entries at 100k clips): the streaming sweep of ``spectral_topk`` instead; neighbors.pkl then holds the
    k nearest neighbours per file in the same schema and distance_matrix.pt is replaced by topk_distances.pt
    ({'distances': [N, k], 'indices': [N, k]}).  With ``world_size`` > 1 every rank sweeps its row block and the results
    meet in one all-gather; rank 0 writes the files."""
    audio_files = list(Path(audio_dir).glob("*.wav"))
    n_files = len(audio_files)
    file_to_idx = {str(f): i for i, f in enumerate(audio_files)}
    audio = _load_padded(audio_files)
    save_path = Path(save_path)
    if top_k is None:
        distances = pairwise_spectral_distances(audio.cuda()).cpu()
        sorted_vals, sorted_idx = neighbour_order(distances)
        result = distances
    else:
        vals, idx = spectral_topk(audio.pin_memory() if audio.numel() else audio, top_k, rank=rank, world_size=world_size)
        vals, idx = gather_topk(vals, idx, n_files)
        sorted_vals, sorted_idx = vals.cpu(), idx.cpu()
        result = {"distances": sorted_vals, "indices": sorted_idx}
    if rank != 0:
        return result
    neighbors = {
        str(audio_files[i]): {
            "sorted_neighbors": [str(audio_files[j]) for j in sorted_idx[i].tolist()],
            "sorted_distances": sorted_vals[i].tolist(),
            "index": i,
        } for i in range(n_files)
    }
    neighbors["__file_to_idx__"] = file_to_idx
    torch.save(result, save_path / ("distance_matrix.pt" if top_k is None else "topk_distances.pt"))
    with open(save_path / "neighbors.pkl", "wb") as f:
        pickle.dump(neighbors, f)
    return result
