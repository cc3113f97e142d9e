"""CPU tests (gloo, world_size 2) of the multi-GPU host-side logic: the unique-id exchange over torch.distributed, the
block-cyclic layout helpers, and a NumPy re-enactment of the distributed sweep of gpx_dist.cu in its MEMORY-DISTRIBUTED
layout (every rank stores only its block rows of the workspace and its column blocks of U), in which each rank only
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
        # memory-distributed storage, as in gpx_dist.cu: SL = the OWNED block rows only (local block row R // G), all columns;
        # SU = the OWNED column blocks of U = L^-T only (local column block k // G), all rows. Nothing of size n x n per rank.
        SL = np.zeros((npr * NB, n))
        SU = np.zeros((n, npr * NB))
        rows = lambda R: slice(gdist.local_slot(R, world) * NB, (gdist.local_slot(R, world) + 1) * NB)
        for R in range(nblk):
            if own(R):
                SL[rows(R)] = np.tril(K)[R * NB:(R + 1) * NB]
        P = np.zeros((world * npr, NB, NB))
        for k in range(nblk):
            sl = slice(k * NB, (k + 1) * NB)
            Bc = np.zeros((NB, NB))
            if own(k):
                D = SL[rows(k), sl]
                L = np.linalg.cholesky(D + np.tril(D, -1).T)
                Li = np.linalg.inv(L)
                SL[rows(k), sl] = np.tril(L, -1) + np.triu(Li.T)      # diag: lower L (strict) | upper U_kk incl. diagonal
                Ldiag = np.diag(L).copy()
                Bc[:] = Li
                P[gdist.chunk_position(k, world, npr)] = Li.T         # assemble: U_kk into the owner's chunk
                logdet_local = 2 * np.log(Ldiag).sum()
            else:
                logdet_local = 0.0
            t = torch.from_numpy(Bc); dist.broadcast(t, src=k % world); Bc = t.numpy()
            for R in range(nblk):
                if R != k and own(R):
                    P[gdist.chunk_position(R, world, npr)] = SL[rows(R), sl] @ Bc.T
            mine = torch.from_numpy(P[rank * npr:(rank + 1) * npr].copy())
            outs = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(outs, mine)
            P = np.concatenate([o_.numpy() for o_ in outs])
            chunk = lambda R: P[gdist.chunk_position(R, world, npr)]
            for R in range(nblk):                                     # filing rules of copyback_kernel
                if R > k and own(R):
                    SL[rows(R), sl] = chunk(R)                        # L panel rows -> row-owned workspace
                if R <= k and k % world == rank:                      # U(:, k) incl. U_kk -> column-owned storage
                    SU[R * NB:(R + 1) * NB, rows(k)] = np.triu(chunk(R)) if R == k else chunk(R)
            for c in range(k + 1, nblk):                              # owned rows: [0..k] (inverse region) and [c..] (SYRK)
                for R in list(range(0, k + 1)) + list(range(c, nblk)):
                    if own(R):
                        SL[rows(R), c * NB:(c + 1) * NB] -= chunk(R) @ chunk(c).T
            if k == 0:
                ld_acc = 0.0
            ld_acc += logdet_local
        # U is column-distributed: rank owns column blocks k = rank (mod world); K^-1 contribution over its k-range
        Kinv_part = np.zeros((n, n))
        tvec = np.zeros(n)
        for k in range(nblk):
            if k % world != rank:
                continue
            Ucol = SU[:, rows(k)]
            Kinv_part += Ucol @ Ucol.T
            tvec[k * NB:(k + 1) * NB] = Ucol.T @ y
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


def _sparse_worker(rank, world, port, q):
    """Row-sharded VarDTC as gpx_sparse_eval runs it with a communicator: each rank holds N/world data rows; num_data,
    trYYT, A = tmp tmp^T and tmp Y (tmp = Lm^-1 psi1^T of the local rows) are all-reduced, every rank repeats the M x M
    algebra, the Knm gradient pieces are all-reduced. Checked against the oracle on the whole data set."""
    import torch
    import torch.distributed as dist
    from oracle import gpy_oracle as o
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, M, D, P = 403, 37, 3, 2
        rng = np.random.default_rng(5)
        X = rng.uniform(-3, 3, (N, D))
        Y = np.stack([np.sin(X).sum(1) + 0.1 * rng.standard_normal(N) for _ in range(P)], 1)
        Z = X[rng.permutation(N)[:M]] + 0.01 * rng.standard_normal((M, D))
        ls, var, noise = np.array([1.1, 1.6, 2.0]), 1.3, 0.07
        lml0, g0, Zg0, res0 = o.sparse_eval(X, Y, Z, "matern52", True, var, ls, noise)
        rows = gdist.shard_rows(N, rank, world)
        Xl, Yl = X[rows], Y[rows]
        kern = o.StationaryOracle("matern52", D, var, ls, True)
        beta = 1.0 / max(noise, 1e-8)

        def allsum(a):
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
            dist.all_reduce(t)
            return t.numpy()

        psi1 = kern.K(Xl, Z)
        ntot, trYYT = allsum(np.array([float(Xl.shape[0]), float((Yl * Yl).sum())]))
        assert int(round(ntot)) == N
        # replicated M x M algebra, in the product form of gpx_sparse.cu; the N-dependent pieces are formed from
        # tmp = Lm^-1 psi1^T of this rank's rows and all-reduced
        Kmm = kern.K(Z) + 1e-8 * np.eye(M)
        Lm = np.linalg.cholesky(Kmm); Lmi = np.linalg.inv(Lm); Um = Lmi.T
        tmp = Lmi @ psi1.T
        Ar = allsum(tmp @ tmp.T)
        t = allsum(tmp @ Yl)
        B = np.eye(M) + beta * Ar
        LB = np.linalg.cholesky(B); LBi = np.linalg.inv(LB); UB = LBi.T
        v = LBi @ (beta * t); w = UB @ v; C = Um @ w
        DBi = P * (UB @ UB.T) + w @ w.T
        dKmm = Um @ (-0.5 * DBi - 0.5 * P * B + P * np.eye(M)) @ Um.T
        W2 = beta * Um @ (P * np.eye(M) - DBi) @ Um.T
        data_fit, trA, sumADB = float((v * v).sum()), beta * np.trace(Ar), beta * float((Ar * DBi).sum())
        psi0_sum = var * ntot
        lml = (-0.5 * ntot * P * (np.log(2 * np.pi) - np.log(beta)) - 0.5 * beta * trYYT
               - 0.5 * P * (beta * psi0_sum - trA) - P * np.log(np.diag(LB)).sum() + 0.5 * data_fit)
        dR = (-0.5 * ntot * P * beta + 0.5 * trYYT * beta ** 2 + 0.5 * P * (psi0_sum * beta ** 2 - trA * beta)
              + beta * (0.5 * sumADB - data_fit))
        # local Knm pieces -> all-reduce
        dKnm = (beta * Yl) @ C.T + psi1 @ W2
        dv1, dl1 = kern.update_gradients_full(dKnm, Xl, Z)
        dZ1 = kern.gradients_X(dKnm.T, Z, Xl)
        pieces = allsum(np.concatenate([[dv1], np.atleast_1d(dl1), dZ1.ravel()]))
        dv2, dl2 = kern.update_gradients_full(dKmm, Z, None)
        dvar = -0.5 * P * beta * ntot + pieces[0] + dv2
        dlen = pieces[1:1 + D] + np.atleast_1d(dl2)
        Zg = pieces[1 + D:].reshape(M, D) + kern.gradients_X(dKmm, Z)
        g = np.concatenate([[dvar], dlen, [dR]])
        assert abs(lml - lml0) < 1e-8 * max(1.0, abs(lml0))
        np.testing.assert_allclose(g, g0, rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(Zg, Zg0, rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(C, res0["woodbury_vector"], rtol=1e-6, atol=1e-8)
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sparse_row_shards():
    import torch.multiprocessing as mp
    assert [gdist.shard_rows(10, r, 3) for r in range(3)] == [slice(0, 4), slice(4, 7), slice(7, 10)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sparse_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def _sparse_het_worker(rank, world, port, q):
    """Row-sharded VarDTC with one noise variance per data point, in the form gpx_sparse_eval_het computes it: tmp = Lm^-1
    psi1^T of the local rows with its COLUMNS scaled by sqrt(beta_n), A and tmp Y all-reduced, the M x M algebra with
    beta = 1, dL_dKnm^T columns scaled by beta_n, per-point noise gradients from the three column reductions
    (s1 = |tmp[:, n]|^2 before the scaling, r = w . tmp[:, n], s2 = |Q psi1^T[:, n]|^2 with Q = LB^-1 Lm^-1); the sums
    over n of log beta, beta and beta |Y_n|^2 are all-reduced. Checked against the oracle (het_noise branches of
    var_dtc.py) on the whole data set; each rank checks ITS rows of dL_dR."""
    import torch
    import torch.distributed as dist
    from oracle import gpy_oracle as o
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, M, D, P = 301, 29, 2, 2
        rng = np.random.default_rng(6)
        X = rng.uniform(-3, 3, (N, D))
        Y = np.stack([np.sin(X).sum(1) + 0.1 * rng.standard_normal(N) for _ in range(P)], 1)
        Z = X[rng.permutation(N)[:M]] + 0.01 * rng.standard_normal((M, D))
        ls, var = np.array([1.2, 1.8]), 1.1
        nv = rng.uniform(0.01, 0.3, N)
        lml0, g0, Zg0, res0 = o.sparse_eval(X, Y, Z, "rbf", True, var, ls, nv)
        rows = gdist.shard_rows(N, rank, world)
        Xl, Yl, beta = X[rows], Y[rows], 1.0 / np.fmax(nv[rows], 1e-8)
        sb = np.sqrt(beta)
        kern = o.StationaryOracle("rbf", D, var, ls, True)

        def allsum(a):
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
            dist.all_reduce(t)
            return t.numpy()

        psi1 = kern.K(Xl, Z)
        ntot = allsum(np.array([float(Xl.shape[0])]))[0]
        Kmm = kern.K(Z) + 1e-8 * np.eye(M)
        Lm = np.linalg.cholesky(Kmm); Lmi = np.linalg.inv(Lm); Um = Lmi.T
        tmp = Lmi @ psi1.T                                   # M x n_local
        s1 = (tmp * tmp).sum(0)
        tmp_s = tmp * sb[None, :]
        Ar = allsum(tmp_s @ tmp_s.T)
        t = allsum(tmp_s @ (sb[:, None] * Yl))
        B = np.eye(M) + Ar
        LB = np.linalg.cholesky(B); LBi = np.linalg.inv(LB); UB = LBi.T
        v = LBi @ t; w = UB @ v; C = Um @ w
        DBi = P * (UB @ UB.T) + w @ w.T
        dKmm = Um @ (-0.5 * DBi - 0.5 * P * B + P * np.eye(M)) @ Um.T
        W2 = Um @ (P * np.eye(M) - DBi) @ Um.T
        r = (w.T @ tmp_s).T / sb[:, None]                    # n_local x P
        Q = LBi @ Lmi
        V2 = Q @ psi1.T
        s2 = (V2 * V2).sum(0)
        sums = allsum(np.array([np.log(beta).sum(), beta.sum(), (beta[:, None] * Yl * Yl).sum()]))
        data_fit, trA = float((v * v).sum()), np.trace(Ar)
        lml = (-0.5 * ntot * P * np.log(2 * np.pi) + 0.5 * P * sums[0] - 0.5 * sums[2]
               - 0.5 * P * (var * sums[1] - trA) - P * np.log(np.diag(LB)).sum() + 0.5 * data_fit)
        b2 = (beta ** 2)[:, None]
        dR = (-0.5 * beta[:, None] + 0.5 * (beta[:, None] * Yl) ** 2 + 0.5 * P * ((var - s1) * beta ** 2)[:, None]
              + 0.5 * (s2 * beta ** 2)[:, None] - r * Yl * b2 + 0.5 * r ** 2 * b2)
        dKnmT = (W2 @ psi1.T) * beta[None, :] + C @ (beta[:, None] * Yl).T
        dv1, dl1 = kern.update_gradients_full(dKnmT.T, Xl, Z)
        dZ1 = kern.gradients_X(dKnmT, Z, Xl)
        pieces = allsum(np.concatenate([[dv1], np.atleast_1d(dl1), dZ1.ravel()]))
        dv2, dl2 = kern.update_gradients_full(dKmm, Z, None)
        dvar = -0.5 * P * sums[1] + pieces[0] + dv2
        dlen = pieces[1:1 + D] + np.atleast_1d(dl2)
        Zg = pieces[1 + D:].reshape(M, D) + kern.gradients_X(dKmm, Z)
        assert abs(lml - lml0) < 1e-8 * max(1.0, abs(lml0))
        np.testing.assert_allclose(np.concatenate([[dvar], dlen]), g0[:1 + D], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(Zg, Zg0, rtol=1e-6, atol=1e-8)
        dR0 = g0[1 + D:].reshape(N, P)[rows]
        np.testing.assert_allclose(dR, dR0, rtol=1e-6, atol=1e-7 * np.abs(dR0).max())
        np.testing.assert_allclose(C, res0["woodbury_vector"], rtol=1e-6, atol=1e-8)
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sparse_heteroscedastic_row_shards():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sparse_het_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
