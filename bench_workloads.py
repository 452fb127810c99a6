"""Secondary workloads of bench.py (same JSON contract as the headline stage workload):

  --workload distance    BASELINE.json configs[4]: the pairwise multi-scale spectral distance sweep
                         (reference precompute_distances.py:51-153), row-sharded over the GPUs, streaming top-k.
  --workload full_step   BASELINE.json configs[3]: the full autoencoder training step (reference trainer.py:260-311),
                         batch-sharded data-parallel with the bucketed gradient all-reduce overlapped with the backward.
"""
from __future__ import annotations

import datetime
import json
import os
import time

import torch

from bench import SEED, ClockSampler, peaks

INT_LANES_PER_SM = 64          # VABSDIFF: one warp instruction per two clocks and SM sub-partition (scripts/probes/sad_rate.cu)


def _dist_env():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


# ==================================================================================================
# distance sweep
# ==================================================================================================
D_METRIC, D_UNIT = "pairwise_spectral_distance_pairs_per_sec", "pairs/s"


def _distance_config(args, world):
    return {"workload": f"pairwise multi-scale spectral distance sweep: {args.clips} synthetic clips of {args.clip_samples} samples "
                        f"(4 s @ 16 kHz NSynth-shaped, randn x 0.1), 5 STFT scales -> 645,864 magnitude bins per clip; "
                        f"every rank sweeps its own row blocks against ALL clips and keeps the {args.top_k} nearest neighbours per row",
            "rows_per_gpu_per_step": args.row_block, "columns": args.clips, "top_k": args.top_k,
            "parallelism": f"row-sharded over {world} GPU(s), no collective on the compute path, one all-gather of the [N, k] results",
            "l2": "the column spectra swept per step (clips x 5.2 MB) exceed the 126 MB L2 by orders of magnitude"}


def _clips(n, samples, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 1, samples, generator=g) * 0.1


def run_distance(args):
    import torch.distributed as dist
    import topo_audio_autoencoder_b200 as T
    from topo_audio_autoencoder_b200 import precompute_distances as pd

    world, rank, local = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    n, R, k = args.clips, args.row_block, args.top_k
    audio_h = _clips(n, args.clip_samples, SEED).pin_memory()       # identical collection on every rank
    # column spectra: prepared once when they fit (pd.spectral_topk does the same), in blocks of 512 clips
    col_block = 4096          # one launch = row_block x 4096 pairs: 128 x 64 L1 tiles -> 512 CTAs per 1024 rows, 3.5 per SM
    cols = [pd.prepare_block(audio_h[c0:c0 + col_block], dev) for c0 in range(0, n, col_block)]
    d_bins, dp = cols[0].d, cols[0].dp
    lo, hi = pd.shard_rows(n, rank, world)
    n_blocks = max(1, (hi - lo) // R)
    scratch = torch.empty(R, k + col_block, dtype=torch.float32, device=dev)
    scratch_i = torch.empty(R, k + col_block, dtype=torch.int64, device=dev)
    res_host = torch.empty(R, k, dtype=torch.float32).pin_memory()
    idx_host = torch.empty(R, k, dtype=torch.int64).pin_memory()

    def sweep(rows, r0):
        best_v, best_i = scratch[:, :k].fill_(float("inf")), scratch_i[:, :k].fill_(-1)
        row_ids = torch.arange(r0, r0 + R, device=dev).unsqueeze(1)
        for ci, c in enumerate(cols):
            c0, w = ci * col_block, c.n
            d = scratch[:, k:k + w]
            rows.block(c, r0, c0, out=d)
            col_ids = torch.arange(c0, c0 + w, device=dev).unsqueeze(0)
            d.masked_fill_(row_ids == col_ids, float("inf"))
            scratch_i[:, k:k + w] = col_ids
            v, sel = torch.topk(scratch[:, :k + w], k, dim=1, largest=False, sorted=True)
            best_i.copy_(torch.gather(scratch_i[:, :k + w], 1, sel))
            best_v.copy_(v)
        return best_v, best_i

    def step(i, host_io):
        r0 = lo + (i % n_blocks) * R
        src = audio_h[r0:r0 + R] if host_io else audio_d[r0 - lo:r0 - lo + R]
        rows = pd.prepare_block(src, dev)                # front half of the row block: (H2D,) STFT, padded spectra + logs
        v, ix = sweep(rows, r0)
        if host_io:
            res_host.copy_(v, non_blocking=True)
            idx_host.copy_(ix, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    audio_d = audio_h[lo:hi].to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(steps, host_io):
        evs = []
        barrier()
        for i in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(i, host_io)
            e1.record()
            evs.append((e0, e1))
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for i in range(max(args.warmup, 3)):
        step(i, False)
    T.lib.reset_counts()
    step(0, False)
    launches_per_step = T.lib.kernel_launches()
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(args.steps, False)
    clocks = sampler.stop() if sampler else None
    timed(1, True)
    ms_e2e = timed(args.steps, True)
    pairs_per_step = world * R * n
    value = pairs_per_step * args.steps / (ms * 1e-3)
    e2e_value = pairs_per_step * args.steps / (ms_e2e * 1e-3)

    # per-kernel device time of one step (CUPTI through torch.profiler: the three kernels of a block call run on two streams)
    roofline = None
    if rank == 0 and not args.no_profile_pass:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(0, False)
            torch.cuda.synchronize()
        kt = {}
        for ev in prof.key_averages():
            for name in ("gram_kernel", "l1_kernel", "combine_kernel", "distance_prepare_kernel"):
                if name in ev.key:
                    kt[name] = kt.get(name, 0.0) + ev.device_time_total / 1e3          # ms
        sm = torch.cuda.get_device_properties(dev).multi_processor_count
        mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
        pair_elems = float(R) * n * d_bins
        l1_ms, gram_ms = kt.get("l1_kernel", 0.0), kt.get("gram_kernel", 0.0)
        int_peak = sm * INT_LANES_PER_SM * mhz * 1e6 / 1e12                             # T instructions / s
        roofline = {"kernel": "l1_kernel (integer pipe, L1-of-logs term)", "bound": "int_alu", "unit": "T integer instructions/s",
                    "achieved": pair_elems / (l1_ms * 1e-3) / 1e12 if l1_ms else None, "peak": int_peak,
                    "frac": (pair_elems / (l1_ms * 1e-3) / 1e12 / int_peak) if l1_ms else None,
                    "peak_source": f"nominal issue rate of the integer pipe: {sm} SMs x 64 lanes x {mhz:.0f} MHz (MEASURED_PEAKS.json has "
                                   "no ALU figure; scripts/probes/sad_rate.cu measured 62.2 of the 64 per clock and SM)",
                    "traffic": None, "kernel_ms_per_step": kt, "share_of_step": l1_ms / (ms / args.steps),
                    "algorithmic_instructions_per_pair_element": 1,
                    "note": "1 integer instruction per pair-element (VABSDIFF.U32 with accumulate on Q6.20 logs; the FP32 form needs 2: "
                            "FADD, FADD|.|): not a GEMM, not HBM-bound (operands are re-read from L2 by the TMA engine)",
                    "gram_kernel": {"bound": "tensor", "unit": "TFLOP/s (bf16, six part products per fp32 product)",
                                    "achieved": 12.0 * pair_elems / (gram_ms * 1e-3) / 1e12 if gram_ms else None,
                                    "peak": json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
                                    if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) else 1400.0,
                                    "operand_bytes_streamed": 2.0 * 6.0 * pair_elems / 128.0,
                                    "note": "128 x 128 tiles: 96 KB of operand image per 64-bin chunk -> bound by L2 / HBM operand streaming, "
                                            "not by the tensor pipe; it is ~1/4 of the step and overlaps nothing yet"}}
        if roofline["gram_kernel"]["achieved"]:
            roofline["gram_kernel"]["frac"] = roofline["gram_kernel"]["achieved"] / roofline["gram_kernel"]["peak"]
            roofline["gram_kernel"]["operand_GBps"] = roofline["gram_kernel"]["operand_bytes_streamed"] / (gram_ms * 1e-3) / 1e9
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = _distance_cpu(args, pairs=24)
    if rank == 0:
        line = {"metric": D_METRIC, "value": value, "unit": D_UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": _distance_config(args, world),
                "e2e": {"value": e2e_value, "unit": D_UNIT, "h2d_bytes_per_step": R * args.clip_samples * 4,
                        "d2h_bytes_per_step": R * k * 12, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "extrapolation_100k_clips_8_gpus_s": (1e5 * 1e5 / 8) / (value / world) if value else None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _distance_cpu(args, pairs):
    """the oracle's per-pair path (both STFTs recomputed for every pair, precompute_distances.py:36-37) on the host cores"""
    from oracle import distance_oracle as do
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    audio = _clips(2 * pairs, args.clip_samples, SEED)
    do.batch_audio_distance(audio[:2], audio[2:4])
    t0 = time.perf_counter()
    for b in range(0, pairs, 8):                       # batches of pairs, as the reference does (batch_size pairs per call)
        do.batch_audio_distance(audio[b:b + 8], audio[pairs + b:pairs + b + 8])
    t = time.perf_counter() - t0
    return {"value": pairs / t, "unit": D_UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{pairs} pairs of {args.clip_samples}-sample clips, 8 pairs per call, both multi-scale STFTs recomputed per "
                      f"pair as precompute_distances.py:36-37 does; os.cpu_count()={os.cpu_count()}"}


def run_distance_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    per = 16
    for _ in range(min(args.warmup, 1)):
        _distance_cpu(args, 8)
    t0 = time.perf_counter()
    recs = [_distance_cpu(args, per) for _ in range(args.steps)]
    t = time.perf_counter() - t0
    value = sum(per / (per / r["value"]) for r in recs) / len(recs)
    line = {"impl": "reference", "metric": D_METRIC, "value": value, "unit": D_UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": _distance_config(args, world),
            "cpu_baseline": dict(recs[-1], value=value),
            "e2e": {"value": value, "unit": D_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ==================================================================================================
# full training step
# ==================================================================================================
F_METRIC, F_UNIT = "full_training_step_samples_per_sec", "samples/s"


def _full_config(args, world):
    return {"workload": f"full autoencoder training step: stock conv front-end on randn [B, 16, 4000] band signals (PQMF is third-party "
                        f"and absent: stated deviation, SURVEY 8(d)) -> complex stage ({args.vertices} vertices, C={args.channels}, "
                        f"{args.layers} SCCN layers, regime={args.regime}) -> decoder consumer -> multi-scale spectral loss + penalties -> "
                        f"backward; {args.micro_batches} micro-batches accumulated, ONE gradient reduction, clip 10, Adam (2 groups)",
            "clips_per_gpu_per_micro_batch": args.batch, "micro_batches_per_step": args.micro_batches,
            "parallelism": f"dp{world}: batch-sharded, DistributedDataParallel buckets (25 MB) all-reduced over NCCL while the last "
                           f"micro-batch's backward is still running",
            "l2": "activations of one micro-batch (GBs) exceed the 126 MB L2; no explicit flush"}


def run_full_step(args):
    import torch.distributed as dist
    import topo_audio_autoencoder_b200 as T

    world, rank, local = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    torch.manual_seed(SEED)
    kw = dict(gate="binary_gumbel", bias_on="probs") if args.regime == "full" else dict(gate="hard_concrete", bias_on="logits")
    model = T.AudioAutoencoder(num_vertices=args.vertices, sccn_hidden_dim=args.channels, **kw)
    n_params = model.num_params()
    tr = T.Trainer(model, device=str(dev), accumulate_grad_batches=args.micro_batches)
    B, M = args.batch, args.micro_batches
    g = torch.Generator().manual_seed(SEED + rank)
    bands_h = [(torch.randn(B, 16, 4000, generator=g) * 0.3).pin_memory() for _ in range(M)]
    bands_d = [b.to(dev) for b in bands_h]

    def step(host_io):
        mb = [b.to(dev, non_blocking=True) for b in bands_h] if host_io else bands_d
        loss = tr.train_step(mb)
        return loss.item() if host_io else loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(steps, host_io):
        evs = []
        barrier()
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(host_io)
            e1.record()
            evs.append((e0, e1))
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for _ in range(max(args.warmup, 3)):
        step(False)
    T.lib.reset_counts()
    step(False)
    launches = T.lib.kernel_launches()
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(args.steps, False)
    clocks = sampler.stop() if sampler else None
    timed(1, True)
    ms_e2e = timed(args.steps, True)
    per_step = world * B * M
    value, e2e_value = per_step * args.steps / (ms * 1e-3), per_step * args.steps / (ms_e2e * 1e-3)
    breakdown = None
    if not args.no_profile_pass:
        # every rank takes this extra step (its last backward is a collective); rank 0 alone reports
        T.lib.start_timing()
        step(False)
        stats = T.lib.stop_timing()
    if rank == 0 and not args.no_profile_pass:
        lib_ms = sum(v[1] for v in stats.values())
        breakdown = {"libtopo_b200_kernels_ms_per_step": lib_ms, "share_of_step": lib_ms / (ms / args.steps),
                     "note": "the rest is the stock PyTorch front-end, decoder tail, loss (cuFFT), optimizer and the all-reduce"}
    if rank == 0:
        hbm, src = peaks()
        line = {"metric": F_METRIC, "value": value, "unit": F_UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(_full_config(args, world), parameters=n_params, gradient_bytes_reduced_per_step=4 * n_params),
                "e2e": {"value": e2e_value, "unit": F_UNIT, "h2d_bytes_per_step": M * B * 16 * 4000 * 4, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches * args.steps, "clocks": clocks, "roofline": None, "cpu_baseline": None,
                "breakdown": breakdown}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_full_step_reference(args):
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"impl": "reference", "unavailable": "the reference's training step raises before it reaches the decoder "
                          "(SURVEY.md 0.1) and needs rave / TopoModelX, which are absent; the stage-level CPU arm is "
                          "`bench.py --impl reference` (default workload)"}), flush=True)
