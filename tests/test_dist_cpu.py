"""World-size-2 gloo tests of the N>1 host logic: gene partition, NCCL-id broadcast plumbing, and the exchange step's
arithmetic (block statistics summed over ranks == statistics of the whole matrix)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gene_block_partition():
    from insider_b200.dist import gene_block
    for P in (1, 31, 32, 33, 5000, 44477, 56200):
        for world in (1, 2, 4, 8):
            blocks = [gene_block(P, world, r) for r in range(world)]
            assert blocks[0][0] == 0
            assert sum(n for _, n in blocks) == P
            for (a, n), (b, _) in zip(blocks, blocks[1:]):
                assert a + n == b
                assert b % 32 == 0 or b == P                 # block boundaries on mask-word boundaries
    assert gene_block(44477, 8, 7) == (38944, 5533)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from insider_b200.dist import allreduce_stats_numpy, broadcast_bytes, env_rank, gene_block
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert env_rank() == (rank, world, rank)
        uid = bytes(range(128)) if rank == 0 else None
        got = broadcast_bytes(uid, 128, 0)
        assert got == bytes(range(128))
        rng = np.random.default_rng(0)                       # same data on every rank
        N, P, K = 20, 100, 4
        Y = rng.normal(size=(N, P)); V = rng.normal(size=(K, P)); M = rng.random((N, P)) < 0.9
        j0, n = gene_block(P, world, rank)
        sl = slice(j0, j0 + n)
        B_loc = (Y[:, sl] * M[:, sl]) @ V[:, sl].T            # B_k = sum_j m_kj y_kj v_j over the rank's genes
        G_loc = np.einsum("ij,aj,bj->iab", M[:, sl].astype(float), V[:, sl], V[:, sl])
        B = allreduce_stats_numpy(B_loc); G = allreduce_stats_numpy(G_loc)
        np.testing.assert_allclose(B, (Y * M) @ V.T, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(G, np.einsum("ij,aj,bj->iab", M.astype(float), V, V), rtol=1e-12, atol=1e-12)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_exchange():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
