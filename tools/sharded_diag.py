"""Diagnostic for the row-sharded step (2 GPUs under torchrun): per Adam step, where do the sharded and the
single-GPU sparse-row weights part, and how do the rows that part look (touch counts per step and rank)."""
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

NU, NI, L, F, NEGS, GROUPS = 7001, 3003, [256, 128, 64], 64, 4, 2001


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lr = 1e-3
    ref = _engine.NeuMFEngine(NU, NI, L, [0, 0, 0], mf_dim=F, table_mode="sparse", seed=21, optimizer="adam", lr=lr)
    w0 = ref.get_weights()
    sh = ShardedNeuMF(NU, NI, L, mf_dim=F, max_local_rows=1 << 15, seed=5, optimizer="adam", lr=lr)
    sh.load_full_tables(w0)
    hist = []
    for step in range(3):
        batches = []
        for r in range(world):
            rng = np.random.default_rng(1000 * step + r)
            u = np.repeat(np.minimum(rng.zipf(1.2, GROUPS) - 1, NU - 1), NEGS + 1)
            i = np.minimum(rng.zipf(1.2, GROUPS * (NEGS + 1)) - 1, NI - 1)
            y = np.tile([0] * NEGS + [1], GROUPS).astype(np.float32)
            batches.append((u, i, y))
        hist.append(batches)
        gu, gi, gy = (np.concatenate([b[j] for b in batches]) for j in range(3))
        sh.train_step(*batches[rank], global_rows=len(gy), group=NEGS + 1, k=3)
        ref.train_step(gu, gi, gy, group=NEGS + 1, k=3)
        got = {k: v.cpu().numpy() for k, v in sh.gather_full_tables().items()}
        for name in sh.cache._dense_slices:
            got[name] = sh.cache._view(name).cpu().numpy()
        want = ref.get_weights()
        if rank == 0:
            for k in want:
                d = np.abs(got[k].reshape(want[k].shape) - want[k])
                line = "step {} {:30s} max diff {:.3e} of lr".format(step + 1, k, d.max() / lr)
                if "embedding" in k:
                    bad = np.flatnonzero(d.max(axis=1) > 1e-3 * lr)
                    line += "  rows off by > 1e-3 lr: {} {}".format(len(bad), bad[:8])
                    ids_idx = 0 if "user" in k else 1
                    for b in bad[:3]:
                        touches = [[int(np.sum(hist[s][r][ids_idx] == b)) for r in range(world)] for s in range(step + 1)]
                        line += "\n      row {} touched [step][rank] {}  worst col diff {:.3e} lr, row |delta| max {:.3e} lr".format(
                            b, touches, d[b].max() / lr, np.abs(want[k][b] - w0[k][b]).max() / lr)
                print(line, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
