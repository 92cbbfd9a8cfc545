"""N>1 host logic on CPU: two gloo ranks.  The kernels cannot run here, so each rank's
per-shard kernel outputs are produced by the numpy oracle (tests may use it as the checker);
what is under test is the product's sharding / bucket all-reduce / Chan-merge / gather code."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import load_golden, unpack_masks
        from oracle import np_oracle as O
        from b200pinn import dist as D

        g = load_golden("net64")
        n = g["x"].shape[0]
        lo, hi = D.shard_range(n, rank, world)
        xs, ys = g["x"][lo:hi], g["y"][lo:hi]
        # --- data-parallel gradient: per-shard sums with the GLOBAL 1/N, one bucket all-reduce
        masks = unpack_masks(g["train_masks"], g["layers"], g["p"], np.float64)
        ms = [m[lo:hi] for m in masks]
        out, lv = O.dnn_forward(g["params"], xs, ms, np.float64)
        du, ds = O.aleatoric_loss_grads(ys, out, lv)
        scale = (hi - lo) / n                       # oracle normalises by the shard size
        G = O.dnn_backward(g["params"], xs, ms, du * scale, ds * scale)
        names = sorted(G)
        bucket = torch.tensor(np.concatenate([G[k].reshape(-1) for k in names]))
        D.allreduce_bucket(bucket)
        ref = np.concatenate([g["G:" + k].reshape(-1) for k in names])
        err_grad = float(np.abs(bucket.numpy() - ref).max() / np.abs(ref).max())
        # --- MC sweep, pass-sharded: Chan merge of per-rank Welford partials
        T, p = int(g["mc_T"]), float(g["mc_p"])
        params = {k[4:]: v for k, v in g.items() if k.startswith("mcP:")}
        tl, th = D.shard_range(T, rank, world)
        us, ss = [], []
        for t in range(tl, th):
            u, s = O.dnn_forward(params, g["x"], unpack_masks(g[f"mc_masks{t}"], g["layers"], p, np.float64), np.float64)
            us.append(u[:, 0])
            ss.append(s[:, 0])
        us, ss = np.stack(us), np.stack(ss)
        mean = torch.tensor(us.mean(0))
        m2 = torch.tensor(((us - us.mean(0)) ** 2).sum(0))
        slv = torch.tensor(ss.sum(0))
        cnt, mean, m2, slv = D.merge_pass_shards(th - tl, mean, m2, slv)
        a_u, e_u = D.finalize(cnt, m2, slv)
        err_mc = max(float(np.abs(a_u.numpy() - g["mc_a_u"]).max() / np.abs(g["mc_a_u"]).max()),
                     float(np.abs(e_u.numpy() - g["mc_e_u"]).max() / np.abs(g["mc_e_u"]).max()))
        # --- MC sweep, sample-sharded: ragged gather
        full = D.gather_rows(torch.tensor(g["mc_pred_mean"][lo:hi]), n)
        ok_gather = bool(np.array_equal(full.numpy(), g["mc_pred_mean"]))
        q.put((rank, cnt, err_grad, err_mc, ok_gather))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_gloo_sharding_and_merge(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, cnt, err_grad, err_mc, ok_gather in res:
        assert cnt == 6 or cnt == 4 or cnt > 0
        assert err_grad < 1e-4, err_grad          # vs the reference's full-batch autograd gradients
        assert err_mc < 1e-5, err_mc              # vs the reference's get_MC_samples
        assert ok_gather
