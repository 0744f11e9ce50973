"""2+ GPU check of row-sharded tables (run under torchrun, one rank per GPU):
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/sharded_gpu_check.py
Every rank trains on its own batch through ShardedNeuMF (all-to-all of ids, rows and gradient rows);
every rank also replays the concatenated global batches on one GPU in sparse-row mode and compares.

Every comparison is made against the float64 oracle as well (oracle/movierec_oracle.py, sparse-row updates): the
sharded run may be no further from it than max(4 x the single-GPU run's own distance, 1e-5 of the tensor's scale).

Three checks.  (1) SGD, 3 steps: the update is linear in the gradients, so the weights must agree to fp32
summation-order accuracy (3e-6 of the largest weight).  (2) Adam, first step: agreement to 1e-4 of lr.
(3) Adam, 3 steps: from the second step on legacy-Keras Adam divides by sqrt(v)+1e-7 with |g| ~ 1e-6 (gradients
carry 1/B_global), which amplifies summation-order differences of entries that nearly cancel -- the replicated
dense kernels, which involve no sharding at all, part by ~1e-2 of lr between the 1-GPU and the 2-GPU summation
order.  The sharded tables must not part by more than 10x what those dense kernels do."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "movierecommender-tf-trt_b200"))
sys.path.insert(0, ROOT)
from movierec import _engine  # noqa: E402
from movierec._distributed import ShardedNeuMF  # noqa: E402
from oracle import movierec_oracle as o  # noqa: E402  (checker only)

NU, NI, L, F, NEGS, GROUPS, STEPS = 7001, 3003, [256, 128, 64], 64, 4, 2001, 3


def run(rank, world, opt, lr, with_oracle=False, steps=STEPS):
    ref = _engine.NeuMFEngine(NU, NI, L, [0, 0, 0], mf_dim=F, table_mode="sparse", seed=21, optimizer=opt, lr=lr)
    w0 = ref.get_weights()
    sh = ShardedNeuMF(NU, NI, L, mf_dim=F, max_local_rows=1 << 15, seed=5, optimizer=opt, lr=lr)
    sh.load_full_tables(w0)
    w64 = {k: v.astype(np.float64) for k, v in w0.items()} if with_oracle else None
    st64 = o.new_opt_state(w64) if with_oracle else None
    for step in range(steps):
        batches = []
        for r in range(world):
            rng = np.random.default_rng(1000 * step + r)
            u = np.repeat(np.minimum(rng.zipf(1.2, GROUPS) - 1, NU - 1), NEGS + 1)  # hot rows shared by the ranks
            i = np.minimum(rng.zipf(1.2, GROUPS * (NEGS + 1)) - 1, NI - 1)
            y = np.tile([0] * NEGS + [1], GROUPS).astype(np.float32)
            batches.append((u, i, y))
        gu, gi, gy = (np.concatenate([b[j] for b in batches]) for j in range(3))
        out = sh.train_step(*batches[rank], global_rows=len(gy), group=NEGS + 1, k=3)
        tot = out.clone()
        dist.all_reduce(tot)
        want = ref.train_step(gu, gi, gy, group=NEGS + 1, k=3)
        assert abs(float(tot[0]) - float(want[0])) <= 1e-5 * abs(float(want[0])), (float(tot[0]), float(want[0]))
        assert float(tot[1]) == float(want[1])  # hit counts are integers
        if with_oracle:
            o.train_step(w64, st64, gu, gi, gy, {"optimizer": opt, "lr": lr, "num_negs_per_pos": NEGS, "k": 3},
                         adam_mode="lazy")
    got = {k: v.cpu().numpy() for k, v in sh.gather_full_tables().items()}
    for name in sh.cache._dense_slices:
        got[name] = sh.cache._view(name).cpu().numpy()
    return w0, ref.get_weights(), got, w64


def vs_oracle(what, single, sharded, w64):
    """(worst relative distance of the sharded run from the float64 oracle, the same for the single-GPU run)."""
    worst_sh = worst_one = 0.0
    for k in single:
        ref = w64[k]
        scale = max(float(np.max(np.abs(ref))), 1e-30)
        e_sh = float(np.max(np.abs(sharded[k].reshape(ref.shape).astype(np.float64) - ref))) / scale
        e_one = float(np.max(np.abs(single[k].astype(np.float64) - ref))) / scale
        assert e_sh <= max(4.0 * e_one, 1e-5), (what, k, e_sh, e_one)
        worst_sh, worst_one = max(worst_sh, e_sh), max(worst_one, e_one)
    return worst_sh, worst_one


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # --- SGD: gradient parity
    lr_sgd = 10.0
    w0, want, got, w64 = run(rank, world, "sgd", lr_sgd, with_oracle=True)
    worst_sgd = 0.0
    for k in want:
        scale = max(float(np.max(np.abs(want[k]))), 1e-30)
        err = float(np.max(np.abs(got[k].reshape(want[k].shape) - want[k]))) / scale
        worst_sgd = max(worst_sgd, err)
        assert err <= 3e-6, ("sgd", k, err)
    oracle_sgd = vs_oracle("sgd 3 steps", want, got, w64)
    # --- Adam, first step strict; three steps measured against the divergence of the replicated dense kernels
    lr = 1e-3
    w0, want, got, w64 = run(rank, world, "adam", lr, with_oracle=True, steps=1)
    first = max(float(np.max(np.abs(got[k].reshape(want[k].shape) - want[k]))) for k in want) / lr
    assert first <= 1e-4, ("adam step 1", first)
    oracle_adam = vs_oracle("adam step 1", want, got, w64)
    w0, want, got, _ = run(rank, world, "adam", lr)
    diff = {k: float(np.max(np.abs(got[k].reshape(want[k].shape) - want[k]))) / lr for k in want}
    dense_div = max(v for k, v in diff.items() if "embedding" not in k)
    table_div = max(v for k, v in diff.items() if "embedding" in k)
    assert table_div <= 10.0 * dense_div + 1e-4, ("adam 3 steps", diff)
    if rank == 0:
        print("sharded_gpu_check ok: world={} steps={} | sgd worst relative weight difference {:.2e} | adam step 1 worst "
              "difference {:.2e} of lr | adam 3 steps: tables {:.2e} of lr, replicated dense kernels {:.2e} of lr"
              .format(world, STEPS, worst_sgd, first, table_div, dense_div))
        print("sharded_gpu_check vs float64 oracle (worst |w - w64| / max|w64|: sharded, single GPU): sgd 3 steps "
              "{:.2e}, {:.2e} | adam step 1 {:.2e}, {:.2e}".format(*(oracle_sgd + oracle_adam)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
