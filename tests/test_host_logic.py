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
