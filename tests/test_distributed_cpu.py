"""Data-parallel step (movierec._distributed) on CPU: two gloo ranks, the oracle standing in for the
CUDA engine.  Checks the group-wise sharding and that local grads (scaled by 1/B_global) + ONE
all-reduce + identical update reproduce the single-process step on the global batch."""

import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from oracle import movierec_oracle as o  # noqa: E402


class OracleEngine(object):
    """Same surface as _engine.NeuMFEngine for the data-parallel wrapper (train_grads /
    gradient_tensors / apply / dense / _tables), computing with the NumPy oracle."""

    table_mode = "dense"

    def __init__(self, weights, params):
        self.params = params
        self.w = {k: v.copy() for k, v in weights.items()}
        self.state = o.new_opt_state(self.w)
        self.names = list(self.w)
        self.sizes = [self.w[k].size for k in self.names]
        self.g_flat = torch.zeros(sum(self.sizes), dtype=torch.float32)
        self.dense = torch.zeros(1)
        self._tables = {}

    def gradient_tensors(self):
        return [self.g_flat]

    def train_grads(self, users, items, labels, group=0, k=0, inv_global_batch=None, grouped=False, dense_l2=True):
        # as the CUDA engine: the hidden kernels' l2 term is part of the gradients only when dense_l2 is set; the
        # tables' l2 term is added in apply(), after the caller's reduction
        c = o.forward(self.w, users, items)
        l2 = list(self.params["layers_l2reg"])
        g = o.backward(self.w, c, labels, inv_global_batch, [0.0] + (l2[1:] if dense_l2 else [0.0] * (len(l2) - 1)))
        self.g_flat.copy_(torch.from_numpy(np.concatenate([g[k_].reshape(-1) for k_ in self.names]).astype(np.float32)))
        y = np.asarray(labels, np.float32)
        return torch.tensor([float(np.sum(o.bce_from_logits(c["z"], y)))])

    def apply(self):
        flat = self.g_flat.numpy()
        g, off = {}, 0
        for k_, n in zip(self.names, self.sizes):
            g[k_] = flat[off:off + n].reshape(self.w[k_].shape).copy()
            if "embedding" in k_ and self.params["layers_l2reg"][0]:
                g[k_] = g[k_] + np.float32(2.0 * self.params["layers_l2reg"][0]) * self.w[k_]
            off += n
        p = self.params
        o.adam_step(self.w, self.state, g, p["lr"], p["beta_1"], p["beta_2"])


class RegionOracleEngine(OracleEngine):
    """Adds the region-wise surface (gradient_regions / apply_region / finish_apply) the wrapper uses with
    MR_DP_OVERLAP=1: one region per parameter tensor, largest first."""

    def __init__(self, weights, params):
        OracleEngine.__init__(self, weights, params)
        self.applied = []

    def gradient_regions(self):
        regs, off = [], 0
        for k_, n in zip(self.names, self.sizes):
            regs.append((k_, self.g_flat[off:off + n]))
            off += n
        return sorted(regs, key=lambda r: -r[1].numel())

    def apply_region(self, name):
        self.applied.append(name)

    def finish_apply(self):
        assert sorted(self.applied) == sorted(self.names), self.applied  # every region exactly once
        self.applied = []
        self.apply()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


PARAMS = {"layers_sizes": [8, 6, 4], "layers_l2reg": [0.0, 0.0, 0.0], "optimizer": "adam", "lr": 0.01, "beta_1": 0.9,
          "beta_2": 0.999, "num_negs_per_pos": 3, "k": 2}


def _global_batch(step):
    rng = np.random.default_rng(100 + step)
    groups, negs = 11, 3  # 11 groups over 2 ranks: uneven split 6 + 5
    users = np.repeat(rng.integers(0, 9, groups), negs + 1)
    items = rng.integers(0, 13, groups * (negs + 1))
    y = np.tile([0] * negs + [1], groups).astype(np.float32)
    return users, items, y


PARAMS_L2 = dict(PARAMS, layers_l2reg=[0.01, 0.02, 0.005])  # the model's own DEFAULT_PARAMS use 0.01


def _worker(rank, world, port, out, overlap=False, l2=False):
    global PARAMS
    if l2:
        PARAMS = PARAMS_L2
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if overlap:
        os.environ["MR_DP_OVERLAP"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from movierec._distributed import DataParallelNeuMF, shard_batch
    w0 = o.init_weights(9, 13, PARAMS["layers_sizes"], 2, np.random.default_rng(1))
    dp = DataParallelNeuMF((RegionOracleEngine if overlap else OracleEngine)(w0, PARAMS))
    assert dp.overlap == overlap
    loss_local = []
    for step in range(3):
        users, items, y = _global_batch(step)
        u, i, l = shard_batch(users, items, y, 4, world, rank)
        assert len(l) % 4 == 0 and len(l) in (24, 20)  # whole groups only
        loss = dp.train_step(u, i, l, global_rows=len(y), group=4, k=2)
        loss_local.append(float(dp.all_reduce_sums(loss.clone())[0]))
    if rank == 0:
        np.savez(out, losses=np.array(loss_local), **dp.engine.w)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap,l2", [(False, False), (True, False), (False, True)],
                         ids=["one-all-reduce", "region-wise", "l2-counted-once"])
def test_two_rank_data_parallel_matches_single_process(tmp_path, overlap, l2):
    """l2 != 0: the regulariser gradients must enter the summed gradients ONCE, not once per rank."""
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out, overlap, l2), nprocs=2, join=True)
    got = np.load(out)
    params = PARAMS_L2 if l2 else PARAMS
    w = o.init_weights(9, 13, params["layers_sizes"], 2, np.random.default_rng(1))
    st = o.new_opt_state(w)
    for step in range(3):
        users, items, y = _global_batch(step)
        pen = o.l2_penalty(w, params["layers_l2reg"])
        loss, _, _ = o.train_step(w, st, users, items, y, params)
        assert got["losses"][step] / len(y) == pytest.approx(loss - pen, rel=1e-5)  # (the workers sum the BCE part)
    for k in w:
        np.testing.assert_allclose(got[k], w[k], rtol=2e-5, atol=1e-7, err_msg=k)


def test_split_groups_covers_everything_once():
    from movierec._distributed import split_groups
    for n in (0, 1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            spans = [split_groups(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


# ---- row-sharded tables: routing of ids, rows and gradient rows (BASELINE config 5) --------------------
def _sharded_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from movierec._distributed import ShardedNeuMF
    sh = object.__new__(ShardedNeuMF)  # the routing needs no CUDA engine
    sh.world, sh.rank, sh.group = world, rank, None
    total, dim = 23, 3
    full = torch.arange(total * dim, dtype=torch.float32).reshape(total, dim)
    shard = full[rank::world].clone()  # owner(row) = row % world, local index = row // world
    rng = np.random.default_rng(40 + rank)
    ids = torch.from_numpy(rng.integers(0, total, 37)).long()
    if rank == 1:
        ids[:5] = 22  # a hot row and a row both ranks ask for
    slots, uniq_s, sc, wanted, rc = sh._route(ids)
    assert sum(sc) == uniq_s.numel() and sum(rc) == wanted.numel()
    assert torch.all(wanted % world == rank)          # only rows this rank owns were requested from it
    local = torch.div(wanted, world, rounding_mode="floor").long()
    rows = sh._a2a(shard[local], rc, sc, width=dim)
    assert torch.equal(rows[slots.long()], full[ids])  # every batch row finds its table row in the cache
    # way back: one gradient row per distinct id (sum over the batch rows that share it) goes to its owner
    g_rows = torch.ones(ids.numel(), dim) * (rank + 1)
    g_unique = torch.zeros(uniq_s.numel(), dim).index_add_(0, slots.long(), g_rows)
    g_recv = sh._a2a(g_unique, sc, rc, width=dim)
    acc = torch.zeros_like(shard).index_add_(0, local, g_recv)
    gathered = [torch.zeros(((total - r + world - 1) // world, dim)) for r in range(world)]
    all_ids = [torch.zeros(37, dtype=torch.long) for _ in range(world)]
    dist.all_gather(all_ids, ids)
    for r in range(world):
        buf = acc.clone() if r == rank else gathered[r]
        dist.broadcast(buf, r)
        gathered[r] = buf
    if rank == 0:
        total_g = torch.zeros(total, dim)
        for r in range(world):
            total_g[r::world] = gathered[r]
        want = torch.zeros(total, dim)
        for r in range(world):
            want.index_add_(0, all_ids[r], torch.ones(37, dim) * (r + 1))
        np.savez(out, got=total_g.numpy(), want=want.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_routing_round_trip(tmp_path, world):
    out = str(tmp_path / "sharded.npz")
    mp.spawn(_sharded_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r = np.load(out)
    assert np.array_equal(r["got"], r["want"])


def test_owner_slices_tile_every_region():
    """The peer exchange's ownership map: per region the ranks' slices are disjoint, in rank order, 16-byte aligned and
    cover the region exactly -- for region sizes that do and do not divide by the world size."""
    from movierec._distributed import owner_slices
    regions = [("user", 0, 64 * 1000, 0.01), ("gmf_user", 64000, 64 * 7, 0.01), ("dense", 64448, 64, 0.0),
               ("item", 64512, 64 * 333, 0.01)]
    for world in (1, 2, 3, 4, 7, 8, 16):
        per_rank = [owner_slices(regions, world, r, lambda n: {"user": 1, "gmf_user": 0}.get(n, 2)) for r in range(world)]
        for i, (name, off, count, l2) in enumerate(regions):
            cur = off
            for r in range(world):
                lo, hi, got_l2, stage = per_rank[r][i]
                assert lo == cur and hi >= lo and lo % 4 == 0 and hi % 4 == 0
                assert got_l2 == l2 and stage == {"user": 1, "gmf_user": 0}.get(name, 2)
                cur = hi
            assert cur == off + count
