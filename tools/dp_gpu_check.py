"""2+ GPU check of the data-parallel step (run under torchrun, one rank per GPU):
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_gpu_check.py
Every rank trains on its shard of a global batch through DataParallelNeuMF (NCCL all-reduce of the flat
gradient buffer); rank 0 replays the same global batches on one GPU and compares the weights."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "movierecommender-tf-trt_b200"))
from movierec import _engine  # noqa: E402
from movierec._distributed import DataParallelNeuMF, shard_batch  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # grouped=False: the per-row launch sequence; grouped=True with nu <= groups / world: the projected step with the
    # fused per-tile kernel, whose user-table gradients are all-reduced early (MrGrads.user_tables_ready); l2 != 0:
    # the hidden kernels' regulariser must be counted once, not once per rank
    # (SGD for the projected sequence: the update is linear in the gradients, so the 1e-6 the two summation orders of
    # the per-user GEMMs differ by stays 1e-6; under Adam, m / (sqrt(v) + eps) turns it into up to 1e-4 of a weight
    # wherever a gradient entry nearly cancels -- see tests/test_gpu_parity.py: OraclePair)
    for grouped, l2, opt in ((False, [0, 0, 0], "adam"), (True, [0, 0, 0], "sgd"), (True, [0.01, 0.02, 0.005], "sgd"),
                             (True, [0, 0, 0], "adam")):
        check(rank, world, grouped, l2, opt)
    dist.barrier()
    dist.destroy_process_group()


def check(rank, world, grouped, l2, opt):
    nu, ni, L, f, negs = (1500 if grouped else 5000), 3000, [256, 128, 64], 64, 4
    eng = _engine.NeuMFEngine(nu, ni, L, l2, mf_dim=f, seed=11 + rank, optimizer=opt, lr=0.5 if opt == "sgd" else 1e-3)  # different seeds: broadcast must fix it
    dp = DataParallelNeuMF(eng)
    dp.broadcast_parameters(0)
    ref = None
    if rank == 0:
        ref = _engine.NeuMFEngine(nu, ni, L, l2, mf_dim=f, seed=11, optimizer=opt, lr=0.5 if opt == "sgd" else 1e-3)
    groups = 4003  # not divisible by the world size
    for step in range(3):
        rng = np.random.default_rng(step)
        users = np.repeat(rng.integers(0, nu, groups), negs + 1)
        items = rng.integers(0, ni, groups * (negs + 1))
        y = np.tile([0] * negs + [1], groups).astype(np.float32)
        u, i, l = shard_batch(users, items, y, negs + 1, world, rank)
        out = dp.train_step(u, i, l, global_rows=len(y), group=negs + 1, k=3, grouped=grouped)
        tot = dp.all_reduce_sums(out.clone())
        if rank == 0:
            want = ref.train_step(users, items, y, group=negs + 1, k=3, grouped=grouped)
            assert abs(float(tot[0]) - float(want[0])) <= 1e-5 * abs(float(want[0])), (tot, want)
            assert float(tot[1]) == float(want[1])
    torch.cuda.synchronize()
    if rank == 0:
        a, b = eng.get_weights(), ref.get_weights()
        worst = 0.0
        for k in a:
            err = float(np.max(np.abs(a[k] - b[k])) / max(np.max(np.abs(b[k])), 1e-30))
            worst = max(worst, err)
            assert err <= (3e-5 if (opt == "sgd" or not grouped) else 1e-3), (k, err)
        print("dp_gpu_check ok: world={} grouped={} l2={} {} early_user_all_reduce={} steps=3 worst relative weight "
              "difference {:.2e}".format(world, grouped, l2, opt, dp.early_user and grouped, worst))
    # every replica must hold identical weights
    h = torch.tensor([float(eng.dense.double().sum()) + float(eng.user_mlp.double().sum())], device="cuda", dtype=torch.float64)
    lo, hi = h.clone(), h.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert float(lo) == float(hi), "replicas diverged"


if __name__ == "__main__":
    main()
