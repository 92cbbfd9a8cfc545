"""Run under torchrun on >= 2 GPUs: correctness of the data-parallel train_dnn step (the one step of the path with a
collective in it).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_p2p_check.py [--quick]

Checks (rank 0 prints one line per check and ``MULTI_GPU_P2P_CHECK PASS`` only if ALL hold and the peer-memory path
was really used; any other outcome prints ``... FAIL`` / ``... SKIPPED`` and the pytest wrapper fails or skips):

1. replicas are synchronised at construction although every rank seeds torch DIFFERENTLY (weights, lambdas, Philox key
   are rank 0's -- broadcast in ``PhysicsInformedNN._sync_replicas``);
2. the fused gradient-sum-over-NVLink + Adam kernel (``pinn_adam_step_p2p``) keeps the replicas bit-identical and agrees
   with the NCCL all-reduce + Adam path;
3. DATA PARALLEL == SINGLE GPU: ``world`` ranks stepping on contiguous shards of a batch reproduce the single-GPU
   full-batch run -- the summed shard gradients match the full-batch gradients to 1e-6 (norm-relative) and the
   parameters after 25 steps to 1e-5: the dropout masks are keyed on GLOBAL rows, so both runs draw the same masks.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import b200pinn
from b200pinn import kernels as K
from b200pinn.dist import shard_range
from b200pinn.synthetic import make_scaled_dataset

LAYERS = [8, 64, 64, 64, 1]


def flat_params(m):
    return torch.cat([p.detach().reshape(-1) for p in m.dnn.kernel_params()]).clone()


def build(x, y, sx, sy, seed, data_parallel=True):
    torch.manual_seed(seed)
    m = b200pinn.PhysicsInformedNN(torch.tensor(x), torch.tensor(y), LAYERS, sx, sy, 0.2, True)
    m.data_parallel = data_parallel
    return m


def run_dp(p2p, steps, x, y, sx, sy, seed):
    os.environ["B200PINN_P2P_ALLREDUCE"] = "1" if p2p else "0"
    m = build(x, y, sx, sy, seed)
    m.train_dnn(steps, verbose=False)
    torch.cuda.synchronize()
    return flat_params(m), getattr(m, "_p2p_bucket", None) is not None


def main():
    quick = "--quick" in sys.argv
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    ok_all, used = True, False
    steps = 8 if quick else 25
    # 80 000 rows in total: plain launches; 24 000: the K2a / K2b / reduce chain runs with programmatic dependent launch in
    # front of the peer-memory Adam kernel
    for n_total in ((24000,) if quick else (80000, 24000)):
        x, y, sx, sy = make_scaled_dataset(n_total, seed=100)
        lo, hi = shard_range(n_total, rank, world)
        xs, ys = x[lo:hi], y[lo:hi]
        # ---- 1. construction sync: every rank seeds differently
        m0 = build(xs, ys, sx, sy, seed=1000 + rank)
        f0 = flat_params(m0)
        g0 = [torch.empty_like(f0) for _ in range(world)]
        dist.all_gather(g0, f0)
        lam = [torch.empty_like(m0._lam) for _ in range(world)]
        dist.all_gather(lam, m0._lam.clone())
        synced = all(torch.equal(g0[0], g) for g in g0) and all(torch.equal(lam[0], t) for t in lam)
        offs = torch.tensor([m0.dnn._row_offset], device=dev)
        go = [torch.empty_like(offs) for _ in range(world)]
        dist.all_gather(go, offs)
        offs_ok = [int(t.item()) for t in go] == [shard_range(n_total, r, world)[0] for r in range(world)]
        # ---- 3a. gradients: sum of the shard gradients (global Philox rows) vs the single-GPU full-batch gradients
        net = K.net_from_module(m0.dnn)
        drop = K.make_dropout(0.2, seed=m0.dnn._drop_seed, sample_offset=lo, pass_offset=5)
        g_sh, _ = K.mlp_backward(net, m0.x.detach(), drop, y=m0.u.reshape(-1).contiguous(), n_global=n_total)
        g_sum = g_sh.clone()
        dist.all_reduce(g_sum)
        del m0
        # ---- 2. peer-memory path vs NCCL path
        a, used = run_dp(True, steps, xs, ys, sx, sy, seed=1000 + rank)
        b, _ = run_dp(False, steps, xs, ys, sx, sy, seed=1000 + rank)
        gathered = [torch.empty_like(a) for _ in range(world)]
        dist.all_gather(gathered, a)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        rel = float((a - b).abs().max() / b.abs().max())
        # ---- 3b. the single-GPU full-batch run (every rank runs it: same work, no collective) from rank 0's weights
        single = build(x, y, sx, sy, seed=1000, data_parallel=False)      # seed of rank 0 = the broadcast source
        single.dnn._drop_seed = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF
        single.dnn._row_offset = 0               # constructed inside the process group, it was given this rank's shard offset
        net1 = K.net_from_module(single.dnn)
        drop1 = K.make_dropout(0.2, seed=single.dnn._drop_seed, sample_offset=0, pass_offset=5)
        g_full, _ = K.mlp_backward(net1, single.x.detach(), drop1, y=single.u.reshape(-1).contiguous(), n_global=n_total)
        g_rel = float((g_sum - g_full).abs().max() / g_full.abs().max())
        single.train_dnn(steps, verbose=False)
        torch.cuda.synchronize()
        p_rel = float((a - flat_params(single)).abs().max() / a.abs().max())
        ok = synced and offs_ok and same and rel < 1e-4 and g_rel < 1e-6 and p_rel < 1e-5
        if not ok:
            print(f"rank {rank}: check failed at n_total={n_total}: synced={synced} offs_ok={offs_ok} same={same} rel={rel:.2e} "
                  f"g_rel={g_rel:.2e} p_rel={p_rel:.2e}", flush=True)
        ok_t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        ok_all = ok_all and bool(ok_t.item())
        if rank == 0:
            print(f"n_total={n_total} world={world}: replicas synced at construction (different seeds): {synced}; shard row offsets ok: "
                  f"{offs_ok}; p2p path used: {used}; replicas bit-identical after {steps} steps: {same}; p2p vs nccl params: "
                  f"{rel:.2e}; DP gradient sum vs single-GPU full batch: {g_rel:.2e}; DP vs single-GPU params after {steps} "
                  f"steps: {p_rel:.2e}", flush=True)
    if rank == 0:
        if not ok_all:
            print("MULTI_GPU_P2P_CHECK FAIL")
        elif used:
            print("MULTI_GPU_P2P_CHECK PASS")
        else:
            print("MULTI_GPU_P2P_CHECK SKIPPED (symmetric memory unavailable; NCCL path and DP == single-GPU verified)")
    dist.barrier()
    dist.destroy_process_group()
    if not ok_all:
        sys.exit(1)


if __name__ == "__main__":
    main()
