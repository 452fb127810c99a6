"""The training step of the reference's ``Trainer`` (trainer.py:49-469), batched and data-parallel.

Kept from the reference, name for name: two Adam groups (encoder 1e-3, decoder 1e-4; :84-87), loss scaled by the number
of accumulated micro-batches (:283-285), global-norm clipping at ``gradient_clip_val`` then ``optimizer.step()`` every
``accumulate_grad_batches`` micro-batches (:288-293), the invalid-state penalty (:278-279), the temperature schedule
written to ``model.encoder.sampler.current_temp`` (:266-269), and the checkpoint dictionary (:417-432).

Data parallel: one process per GPU, batch-sharded, parameters replicated.  The model is wrapped in
``DistributedDataParallel``: gradients are accumulated locally for the first micro-batches (``no_sync``) and the LAST
micro-batch's backward launches the bucketed NCCL all-reduce of the full gradient (~18 M fp32, 25 MB buckets) while the
rest of that backward is still running; clipping sees the reduced gradients.  Out of scope: early stopping,
grid search, gradient-norm reports, audio sample dumps, DataLoader plumbing.
"""
from __future__ import annotations

from pathlib import Path
from typing import Iterable, Optional, Sequence

import torch
import torch.distributed as dist
from torch.optim import Adam

from .loss import AutoencoderLoss


class Trainer:
    def __init__(self, model, checkpoint_dir: Optional[str] = None, encoder_lr: float = 1e-3, decoder_lr: float = 1e-4,
                 initial_reg_factor: float = 0.00001, invalid_state_penalty: float = 100.0, device: str = "cuda",
                 seed: int = 511990, initial_temp: float = 5.0, min_temp: float = 0.1, temp_decay: float = 0.95,
                 gradient_clip_val: float = 10.0, accumulate_grad_batches: int = 4, bucket_cap_mb: int = 25,
                 cudnn_benchmark: bool = True):
        torch.manual_seed(seed)                                                                   # :462-469
        self.device = torch.device(device)
        if cudnn_benchmark and self.device.type == "cuda":
            # The stock front-end and decoder convolutions see the same shapes every step: let cuDNN time its algorithms once
            # (process-wide PyTorch switch; fp32 results stay within the spread of cuDNN's own algorithms).  Measured on a
            # B200: 22 ms of a 178 ms step, mostly the grouped convolutions' data gradient.
            torch.backends.cudnn.benchmark = True
        self.model = model.to(self.device)
        self.checkpoint_dir = Path(checkpoint_dir) if checkpoint_dir is not None else None
        if self.checkpoint_dir is not None:
            self.checkpoint_dir.mkdir(parents=True, exist_ok=True)
        self.optimizer = Adam([{"params": self.model.encoder.parameters(), "lr": encoder_lr},     # :81-87
                               {"params": self.model.decoder.parameters(), "lr": decoder_lr}],
                              fused=self.device.type == "cuda")          # one multi-tensor kernel per group on the GPU, same update
        self.loss_fn = AutoencoderLoss(binary_entropy_penalty=initial_reg_factor, min_entropy_penalty=0.01,
                                       complexity_penalty=0.1)                                   # :97-101
        self.invalid_state_penalty = invalid_state_penalty
        self.accumulation_steps = accumulate_grad_batches
        self.gradient_clip_val = gradient_clip_val
        self.current_temp, self.min_temp, self.temp_decay = initial_temp, min_temp, temp_decay
        self.metrics = {"train_losses": [], "val_losses": []}
        self.ddp = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from torch.nn.parallel import DistributedDataParallel
            ids = [self.device.index] if self.device.type == "cuda" else None
            self.ddp = DistributedDataParallel(self.model, device_ids=ids, bucket_cap_mb=bucket_cap_mb,
                                               gradient_as_bucket_view=True, find_unused_parameters=True)
        self.last_grad_norm: Optional[torch.Tensor] = None

    # ---- trainer.py:266-269 ----
    def set_epoch(self, epoch: int) -> float:
        temp = max(self.min_temp, self.current_temp * (self.temp_decay ** epoch))
        self.model.encoder.sampler.current_temp = temp
        return temp

    def micro_batch_loss(self, bands: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """forward + loss of one micro-batch (trainer.py:276-284), already divided by the accumulation count"""
        net = self.ddp if self.ddp is not None else self.model
        out, diversity, valid = net(bands, noise)
        n_bad = int((~valid).sum())
        penalty = self.invalid_state_penalty * n_bad / max(bands.shape[0], 1)
        if out is None:
            # no clip of the micro-batch decoded: the penalty is a constant, the structural penalties still train the gate
            loss = penalty + 0.0 * (diversity["diversity"].sum() + diversity["binary_entropy"].sum())
        else:
            v = valid.to(bands.device)
            div = {k: t[v] for k, t in diversity.items()}
            loss = self.loss_fn(out, bands[v], div, record=False) * (float(valid.sum()) / bands.shape[0]) + penalty
        return loss / self.accumulation_steps

    def train_step(self, micro_batches: Sequence, noises: Optional[Sequence] = None) -> torch.Tensor:
        """One optimizer step over ``accumulate_grad_batches`` micro-batches (trainer.py:271-293).  -> summed loss (device)."""
        assert len(micro_batches) == self.accumulation_steps
        total = None
        for i, bands in enumerate(micro_batches):
            last = i == len(micro_batches) - 1
            noise = None if noises is None else noises[i]
            if self.ddp is not None and not last:
                with self.ddp.no_sync():                       # accumulate locally; only the last backward reduces
                    loss = self.micro_batch_loss(bands, noise)
                    loss.backward()
            else:
                loss = self.micro_batch_loss(bands, noise)
                loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        self.last_grad_norm = torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.gradient_clip_val)   # :290
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        return total

    # ---- trainer.py:417-453 ----
    def save_checkpoint(self, name: str, checkpoint_dir: Optional[Path] = None) -> Path:
        checkpoint_dir = self.checkpoint_dir if checkpoint_dir is None else Path(checkpoint_dir)
        checkpoint = {
            "model_state_dict": self.model.state_dict(),
            "optimizer_state_dict": self.optimizer.state_dict(),
            "metrics": self.metrics,
            "hyperparameters": {"encoder_lr": self.optimizer.param_groups[0]["lr"],
                                "decoder_lr": self.optimizer.param_groups[1]["lr"],
                                "complexity_penalty": self.loss_fn.complexity_penalty},
        }
        path = checkpoint_dir / f"{name}.pt"
        torch.save(checkpoint, path)
        return path

    def load_checkpoint(self, name, checkpoint_dir: Optional[Path] = None) -> None:
        checkpoint_dir = self.checkpoint_dir if checkpoint_dir is None else Path(checkpoint_dir)
        path = name if isinstance(name, Path) else checkpoint_dir / f"{name}.pt"
        checkpoint = torch.load(path, map_location=self.device, weights_only=False)
        self.model.load_state_dict(checkpoint["model_state_dict"])
        self.optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
        self.metrics = checkpoint["metrics"]
        params = checkpoint["hyperparameters"]
        self.optimizer.param_groups[0]["lr"] = params["encoder_lr"]
        self.optimizer.param_groups[1]["lr"] = params["decoder_lr"]
        self.loss_fn.complexity_penalty = params["complexity_penalty"]
