"""CPU tests (gloo, world_size 2) of the multi-GPU host-side logic: the unique-id exchange over torch.distributed, the
block-cyclic layout helpers, and a NumPy re-enactment of the distributed sweep of gpx_dist.cu in which each rank only
touches the block rows it owns and the all-gather / broadcast / all-reduce steps go through gloo."""
import os
import socket

import numpy as np
import pytest

from gpy_b200 import dist as gdist


def test_block_layout_and_chunk_positions():
    Npad, nblk, npr = gdist.block_layout(16384, 512, 8)
    assert (Npad, nblk, npr) == (16384, 32, 4)
    Npad, nblk, npr = gdist.block_layout(3000, 256, 4)
    assert (Npad, nblk, npr) == (3072, 12, 3)
    for G in (1, 2, 3, 8):
        npr = -(-13 // G)
        pos = [gdist.chunk_position(R, G, npr) for R in range(13)]
        assert len(set(pos)) == 13                                   # a permutation into G*npr slots
        for R in range(13):                                          # each rank's chunks are contiguous
            assert pos[R] // npr == gdist.block_owner(R, G)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = gdist.exchange_unique_id(lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        # ---- distributed sweep in NumPy: T = tile, NB = block, G ranks; mirrors gpx_dist.cu::dist_exact_eval --------
        T, NB, n = 2, 4, 24
        rng = np.random.default_rng(0)
        X = rng.uniform(-3, 3, (n, 2))
        K = np.exp(-0.5 * ((X[:, None] - X[None]) ** 2).sum(-1) / 2.0) + 0.05 * np.eye(n)
        y = rng.standard_normal(n)
        nblk = n // NB
        npr = -(-nblk // world)
        own = lambda R: R % world == rank
        S = np.zeros((n, n))
        for R in range(nblk):                                         # owned block rows only
            if own(R):
                S[R * NB:(R + 1) * NB] = np.tril(K)[R * NB:(R + 1) * NB]
        P = np.zeros((world * npr, NB, NB))
        for k in range(nblk):
            sl = slice(k * NB, (k + 1) * NB)
            Bc = np.zeros((NB, NB))
            if own(k):
                L = np.linalg.cholesky(S[sl, sl] + np.tril(S[sl, sl], -1).T)
                Li = np.linalg.inv(L)
                S[sl, sl] = np.tril(L, -1) + np.triu(Li.T)           # diag: lower L (strict) | upper U_kk incl. diagonal
                Ldiag = np.diag(L).copy()
                Bc[:] = Li
                P[gdist.chunk_position(k, world, npr)] = Li.T
                logdet_local = 2 * np.log(Ldiag).sum()
            else:
                logdet_local = 0.0
            t = torch.from_numpy(Bc); dist.broadcast(t, src=k % world); Bc = t.numpy()
            for R in range(nblk):
                if R != k and own(R):
                    P[gdist.chunk_position(R, world, npr)] = S[R * NB:(R + 1) * NB, sl] @ Bc.T
            mine = torch.from_numpy(P[rank * npr:(rank + 1) * npr].copy())
            outs = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(outs, mine)
            P = np.concatenate([o_.numpy() for o_ in outs])
            chunk = lambda R: P[gdist.chunk_position(R, world, npr)]
            for R in range(nblk):                                     # copy-back rules of copyback_kernel
                if R == k:
                    continue
                if (R > k and own(R)) or (R < k and k % world == rank):
                    S[R * NB:(R + 1) * NB, sl] = chunk(R)
            for c in range(k + 1, nblk):                              # owned rows: [0..k] (inverse region) and [c..] (SYRK)
                for R in list(range(0, k + 1)) + list(range(c, nblk)):
                    if own(R):
                        S[R * NB:(R + 1) * NB, c * NB:(c + 1) * NB] -= chunk(R) @ chunk(c).T
            if k == 0:
                ld_acc = 0.0
            ld_acc += logdet_local
        # U is column-distributed: rank owns column blocks k = rank (mod world); K^-1 contribution over its k-range
        Kinv_part = np.zeros((n, n))
        tvec = np.zeros(n)
        for k in range(nblk):
            if k % world != rank:
                continue
            sl = slice(k * NB, (k + 1) * NB)
            Ucol = np.zeros((n, NB))
            Ucol[:k * NB] = S[:k * NB, sl]
            Ucol[sl] = np.triu(S[sl, sl])
            Kinv_part += Ucol @ Ucol.T
            tvec[sl] = Ucol.T @ y
        tt = torch.from_numpy(tvec); dist.all_reduce(tt); tvec = tt.numpy()
        kk = torch.from_numpy(Kinv_part); dist.all_reduce(kk)
        ldt = torch.tensor([ld_acc]); dist.all_reduce(ldt)
        Kinv_ref = np.linalg.inv(K)
        assert np.abs(kk.numpy() - Kinv_ref).max() < 1e-9 * np.abs(Kinv_ref).max()
        assert abs(float(ldt[0]) - np.linalg.slogdet(K)[1]) < 1e-10
        assert abs(tvec @ tvec - y @ Kinv_ref @ y) < 1e-9
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sweep():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
