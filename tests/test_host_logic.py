"""CPU: host-side logic of the drop-in layer (index plumbing, sharding, gradient all-reduce over gloo)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_csr_pattern_and_its_transpose():
    from topo_audio_autoencoder_b200.custom_sccn import CsrPattern
    g = torch.Generator().manual_seed(0)
    dense = (torch.rand(7, 5, generator=g) < 0.4).float() * torch.rand(7, 5, generator=g)
    sp = dense.to_sparse().coalesce()
    pat = CsrPattern(sp.indices(), sp.shape)
    vals = sp.values()
    crow, col = pat.row_ptr.long(), pat.col.long()
    rebuilt = torch.zeros_like(dense)
    for r in range(7):
        for e in range(crow[r], crow[r + 1]):
            rebuilt[r, col[e]] = vals[e]
    assert torch.equal(rebuilt, dense)
    crow_t, col_t, vals_t = pat.row_ptr_t.long(), pat.col_t.long(), vals[pat.perm]
    rebuilt_t = torch.zeros(5, 7)
    for r in range(5):
        for e in range(crow_t[r], crow_t[r + 1]):
            rebuilt_t[r, col_t[e]] = vals_t[e]
    assert torch.equal(rebuilt_t, dense.t())


def test_index_sets_to_device_layout_matches_active_set_kernel_layout():
    from topo_audio_autoencoder_b200.complex_builder import index_sets_to_device_layout
    from topo_audio_autoencoder_b200.rectifier import _Tables
    t = _Tables(5, upload=False)
    act = {"vertices": torch.tensor([0, 2, 4]), "edges": torch.tensor([1, 3, 9]), "triangles": torch.tensor([], dtype=torch.long),
           "tetra": torch.tensor([4])}
    pos, idx, counts_dev, counts = index_sets_to_device_layout(act, t, "cpu")
    assert counts == [3, 3, 0, 1] and counts_dev.tolist() == counts
    o = t.offsets
    assert pos[o[0]:o[1]].tolist() == [0, -1, 1, -1, 2]
    assert idx[o[1]:o[1] + 3].tolist() == [1, 3, 9] and idx[o[1] + 3] == -1
    assert pos[o[3] + 4] == 0 and (pos[o[2]:o[3]] == -1).all()


def test_row_and_batch_sharding_cover_everything_once():
    from topo_audio_autoencoder_b200.dist import shard_batch
    from topo_audio_autoencoder_b200.precompute_distances import shard_rows
    for total in (0, 1, 7, 64, 100000):
        for world in (1, 2, 3, 8):
            spans = [shard_batch(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
            rows = [shard_rows(total, r, world) for r in range(world)]
            assert sum(e - b for b, e in rows) == total and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from topo_audio_autoencoder_b200.dist import allreduce_gradients, shard_batch
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Linear(4, 1))     # identical replicas
    data = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10
    lo, hi = shard_batch(8, rank, world)
    # per-rank mean over its shard; averaged over ranks == mean over the global batch (equal shards)
    model(data[lo:hi]).mean().backward()
    allreduce_gradients(model.parameters())
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    if rank == 0:
        torch.save(flat, out)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world_size_2_gloo(tmp_path):
    out = str(tmp_path / "grads.pt")
    mp.spawn(_dp_worker, args=(2, 29541, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Linear(4, 1))
    data = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10
    model(data).mean().backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)


def test_sm_partition_between_concurrent_rank_launches():
    """custom_sccn._sm_shares: every live launch gets at least one SM, never more than it has tiles, the shares
    use the whole GPU when there is enough work, and the launches finish together (the four ranks of a layer run
    side by side on these partitions)."""
    from topo_audio_autoencoder_b200.custom_sccn import _sm_shares
    tiles = [10, 95, 570, 2423]                                   # 128-row tiles of the full 20-vertex complex, 64 clips
    costs = [t * (n + 1.2) for t, n in zip(tiles, (2, 3, 3, 2))]
    s = _sm_shares(costs, 148, tiles)
    assert sum(s) == 148 and min(s) >= 1
    assert s[3] > s[2] > s[1] >= s[0]
    finish = [-(-t // k) * c / t for t, k, c in zip(tiles, s, costs)]
    assert max(finish[1:]) / min(finish[1:]) < 1.15, finish       # the three big launches end within 15 % of each other
    assert _sm_shares([0, 5, 0, 5], 148, [0, 5, 0, 5]) == [0, 5, 0, 5]      # one CTA per tile is the most a launch can use
    small = _sm_shares([1, 1, 1, 10 ** 9], 8, [1, 1, 1, 1000])
    assert sum(small) == 8 and small[:3] == [1, 1, 1]
    assert _sm_shares([0, 0, 0, 0], 148) == [0, 0, 0, 0]


def test_tile_fragment_layout_is_a_permutation_of_each_tile():
    """layout.cuh: float4 index = tile * 2048 + (q * 4 + j) * 128 + r; the chunk-map view addresses the same
    elements (the test mirrors the two device functions)."""
    import numpy as np
    rows = 3 * 128
    seen = np.zeros(rows * 16, dtype=np.int64)
    for row in range(rows):
        row0, r = (row // 128) * 128, row % 128
        for q in range(4):
            base = (row0 >> 7) * 2048 + q * 512 + r                # tf_index(row0, q, r)
            for j in range(4):
                seen[base + j * 128] += 1
        for c in range(8):                                         # tf_index_chunk(row0, r, c) and + 128
            base = (row0 >> 7) * 2048 + ((c >> 1) * 4 + (c & 1) * 2) * 128 + r
            q, j0 = c // 2, (c % 2) * 2
            assert base == (row0 >> 7) * 2048 + (q * 4 + j0) * 128 + r
    assert (seen == 1).all()


def test_row_map_lane_transpose_is_an_involution_with_the_documented_access_pattern():
    """layout.cuh rowmap_transpose4: two butterfly steps over every group of four lanes.  Mirrors the device function on a
    [32 lanes][4 slots] table of (row, chunk) labels: afterwards lane 4g + j holds chunk j of rows 4g .. 4g + 3 in slot order
    (so access number i of a warp touches 64 contiguous bytes of rows 4g + i), and applying it twice is the identity."""
    def transpose4(v):            # v[lane][slot]
        v = [list(lane) for lane in v]
        for dist_, pairs in ((2, ((0, 2), (1, 3))), (1, ((0, 1), (2, 3)))):
            for lo, hi_slot in pairs:
                send = [v[l][lo] if (l & dist_) else v[l][hi_slot] for l in range(32)]     # what every lane hands over
                recv = [send[l ^ dist_] for l in range(32)]                                 # __shfl_xor_sync
                for l in range(32):
                    if l & dist_:
                        v[l][lo] = recv[l]
                    else:
                        v[l][hi_slot] = recv[l]
        return v

    start = [[(lane, chunk) for chunk in range(4)] for lane in range(32)]      # row map: lane l holds chunks 0..3 of row l
    t = transpose4(start)
    for lane in range(32):
        g, j = lane // 4, lane % 4
        assert t[lane] == [(4 * g + i, j) for i in range(4)]
    assert transpose4(t) == start
    # one warp access i: the bytes each row receives are one contiguous 64-byte run (chunks 0..3 of 16 bytes)
    for i in range(4):
        by_row = {}
        for lane in range(32):
            row, chunk = t[lane][i]
            by_row.setdefault(row, []).append(chunk)
        assert len(by_row) == 8 and all(sorted(c) == [0, 1, 2, 3] for c in by_row.values())


def _gather_topk_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from topo_audio_autoencoder_b200.precompute_distances import gather_topk, shard_rows
    n, k = 7, 3
    lo, hi = shard_rows(n, rank, world)
    vals = torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1) + torch.arange(k, dtype=torch.float32) * 0.1
    idx = (torch.arange(lo, hi).unsqueeze(1) * 10 + torch.arange(k)).to(torch.int64)
    gv, gi = gather_topk(vals, idx, n)
    q.put((rank, gv.numpy().copy(), gi.numpy().copy()))
    dist.destroy_process_group()


def test_sharded_sweep_results_meet_in_one_gather():
    """distance sweep, world size 2 over gloo: row shards of unequal size are padded, gathered once and trimmed"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_topk_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_v = torch.arange(7, dtype=torch.float32).unsqueeze(1) + torch.arange(3, dtype=torch.float32) * 0.1
    want_i = (torch.arange(7).unsqueeze(1) * 10 + torch.arange(3)).to(torch.int64)
    for _, gv, gi in got:
        assert torch.equal(torch.from_numpy(gv), want_v) and torch.equal(torch.from_numpy(gi), want_i)


class _TinyAutoencoder(torch.nn.Module):
    """The interface Trainer needs (encoder with a sampler, decoder, forward -> (out, diversity, valid)) in pure torch."""

    def __init__(self):
        super().__init__()
        self.encoder = torch.nn.Linear(4000, 8)
        self.encoder.sampler = type("S", (), {"current_temp": 1.0})()
        self.decoder = torch.nn.Linear(8, 4000)

    def forward(self, bands, noise=None):
        z = self.encoder(bands)
        out = self.decoder(torch.tanh(z))
        valid = torch.ones(bands.shape[0], dtype=torch.bool)
        div = {"binary_entropy": z.mean(dim=(1, 2)), "diversity": z.pow(2).mean(dim=(1, 2))}
        return out, div, valid


def _trainer_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from topo_audio_autoencoder_b200.trainer import Trainer
    torch.manual_seed(0)
    tr = Trainer(_TinyAutoencoder(), device="cpu", accumulate_grad_batches=4)
    assert tr.ddp is not None
    g = torch.Generator().manual_seed(100)
    data = torch.randn(2 * 4, 2, 2, 4000, generator=g) * 0.1            # [rank * micro, B, bands, T]
    mine = [data[rank * 4 + i] for i in range(4)]
    loss = tr.train_step(mine)
    q.put((rank, loss.item(), {k: v.numpy().copy() for k, v in tr.model.state_dict().items()}, float(tr.last_grad_norm)))
    dist.destroy_process_group()


def test_data_parallel_training_step_accumulates_locally_and_reduces_once():
    """trainer.py:271-293 on two gloo ranks: four micro-batches per rank, one gradient reduction, clipping after it.
    Both replicas end with identical parameters, equal to a single process that saw all eight micro-batches."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7) % 2000
    procs = [ctx.Process(target=_trainer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, _, sd0, n0), (_, _, sd1, n1) = got
    assert abs(n0 - n1) < 1e-6 * max(n0, 1.0), "clipping must see the reduced gradients on every rank"
    sd0 = {k: torch.from_numpy(v) for k, v in sd0.items()}
    sd1 = {k: torch.from_numpy(v) for k, v in sd1.items()}
    for k in sd0:
        assert torch.equal(sd0[k], sd1[k]), f"replicas diverged at {k}"
    # single-process reference: the mean over ranks of the per-rank accumulated gradient
    from topo_audio_autoencoder_b200.trainer import Trainer
    torch.manual_seed(0)
    tr = Trainer(_TinyAutoencoder(), device="cpu", accumulate_grad_batches=4)
    g = torch.Generator().manual_seed(100)
    data = torch.randn(2 * 4, 2, 2, 4000, generator=g) * 0.1
    for rank in range(2):
        for i in range(4):
            (tr.micro_batch_loss(data[rank * 4 + i]) / 2).backward()
    torch.nn.utils.clip_grad_norm_(tr.model.parameters(), tr.gradient_clip_val)
    tr.optimizer.step()
    for k, v in tr.model.state_dict().items():
        assert torch.allclose(v, sd0[k], rtol=1e-5, atol=1e-6), k


def test_checkpoint_dictionary_has_the_reference_schema(tmp_path):
    from topo_audio_autoencoder_b200.trainer import Trainer
    tr = Trainer(_TinyAutoencoder(), checkpoint_dir=str(tmp_path), device="cpu")
    assert [g["lr"] for g in tr.optimizer.param_groups] == [1e-3, 1e-4]            # trainer.py:84-87
    tr.train_step([torch.randn(2, 2, 4000) * 0.1 for _ in range(4)])
    path = tr.save_checkpoint("epoch_0_iter_0")
    ck = torch.load(path, weights_only=False)
    assert sorted(ck) == ["hyperparameters", "metrics", "model_state_dict", "optimizer_state_dict"]      # :421-431
    assert sorted(ck["hyperparameters"]) == ["complexity_penalty", "decoder_lr", "encoder_lr"]
    tr.optimizer.param_groups[0]["lr"] = 5.0
    tr.load_checkpoint("epoch_0_iter_0")
    assert tr.optimizer.param_groups[0]["lr"] == 1e-3
    assert tr.set_epoch(3) == max(0.1, 5.0 * 0.95 ** 3) and tr.model.encoder.sampler.current_temp == tr.set_epoch(3)
