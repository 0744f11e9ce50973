"""Parity of the CUDA path (through the C ABI) against the CPU oracle, on a B200.

Tolerances (BASELINE.json north_star): ids / ranks / positions bit-exact given equal scores;
fp32 logits, loss and updated weights within 1e-5 relative; HR@k / NDCG@k within 1e-3 absolute.
"""

import json
import os

import numpy as np
import pytest

from oracle import movierec_oracle as o

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

RTOL = 1e-5  # north_star: fp32 logits, loss, updated weights within 1e-5 relative


@pytest.fixture(scope="module")
def eng_mod():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from movierec import _engine
    return _engine


def make_batch(rng, nu, ni, groups, negs):
    users = np.repeat(rng.integers(0, nu, groups), negs + 1)
    items = rng.integers(0, ni, groups * (negs + 1))
    y = np.tile([0] * negs + [1], groups).astype(np.float32)
    return users, items, y


def rel_close(got, want, rtol=RTOL, what="", atol=0.0):
    """Relative to the magnitude of the reference tensor (a per-tensor scale keeps entries that
    cancel to ~0 from demanding absolute precision fp32 cannot give)."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = max(float(np.max(np.abs(want))), 1e-30)
    err = float(np.max(np.abs(got - want)))
    assert err <= rtol * scale + atol, "{}: max err {:.3e} > {:.1e} * max|ref| {:.3e} + {:.1e}".format(
        what, err, rtol, scale, atol)


def close_to_truth(got, ref32, ref64, what, rtol=RTOL, k=4.0, slack=None, elementwise=True):
    """The kernel against the float64 oracle, calibrated by the float32 oracle's own distance from it:
      per tensor:  max |got - ref64| <= max(k * e32, rtol * max|ref64|) + max(slack),  e32 = max |ref32 - ref64|
                   (within the north star's tolerance of the tensor's scale, or as close to the truth as a plain
                   fp32 evaluation of the same formulas gets -- what decides for sums that nearly cancel);
      per element, where |ref64| > 1e-3 * max|ref64|:  |got - ref64| <= rtol * |ref64| + k * e32 + slack
                   (the north-star quantities: logits, probabilities, updated weights.  Not applied to gradient
                   tensors: a gradient entry is a dot product over the batch whose rounding error scales with
                   sum |a_k b_k|, not with the entry, so only the per-tensor scale is meaningful for it).
    slack (per element, optional): what an ALLOWED error of the inputs does to this value to first order -- for
    updated weights |d update / d gradient| x (rtol x the gradient tensor's scale), see OraclePair.step: Adam's
    m / (sqrt(v) + eps) has a slope of up to lr / eps where a gradient entry nearly cancels, so two correct fp32
    summation orders of the gradient legitimately move such a weight by more than rtol of it."""
    got = np.asarray(got, np.float64).reshape(-1)
    ref32 = np.asarray(ref32, np.float64).reshape(-1)
    ref64 = np.asarray(ref64, np.float64).reshape(-1)
    slack = np.zeros(1) if slack is None else np.asarray(slack, np.float64).reshape(-1)
    scale = max(float(np.max(np.abs(ref64))), 1e-30)
    e32 = float(np.max(np.abs(ref32 - ref64)))
    err = np.abs(got - ref64)
    bound = max(k * e32, rtol * scale) + float(slack.max())
    assert float(err.max()) <= bound, "{}: max err {:.3e} > max({} x fp32-oracle err {:.3e}, {:.0e} x scale {:.3e}) + {:.3e}".format(
        what, float(err.max()), k, e32, rtol, scale, float(slack.max()))
    big = np.abs(ref64) > 1e-3 * scale
    if elementwise and big.any():
        sl = slack[big] if slack.size == got.size else float(slack.max())
        excess = err[big] - (rtol * np.abs(ref64[big]) + k * e32 + sl)
        i = int(np.argmax(excess))
        assert excess[i] <= 0, "{}: element {} got {:.9e} want {:.9e} (fp32-oracle err {:.3e})".format(
            what, i, got[big][i], ref64[big][i], e32)


class OraclePair:
    """The fp32 oracle and a float64 copy of it stepping through the same batches (the truth `close_to_truth` wants),
    plus the first-order slack of the updated weights: the gradient parity tolerance (rtol of each gradient tensor's
    scale) pushed through the optimizer's update, accumulated over the steps taken."""

    def __init__(self, w):
        self.w = w
        self.st = o.new_opt_state(w)
        self.w64 = {k: v.astype(np.float64) for k, v in w.items()}
        self.st64 = o.new_opt_state(self.w64)
        self.slack = {k: np.zeros(v.shape) for k, v in w.items()}

    def grads(self, users, items, y, l2):
        g = o.backward(self.w, o.forward(self.w, users, items), y, None, l2)
        g64 = o.backward(self.w64, o.forward(self.w64, users, items), y.astype(np.float64), None, l2)
        return g, g64

    def step(self, users, items, y, params, adam_mode="dense", g64=None):
        out = o.train_step(self.w, self.st, users, items, y, params, adam_mode=adam_mode)
        out64 = o.train_step(self.w64, self.st64, users, items, y.astype(np.float64), params, adam_mode=adam_mode)
        if g64 is not None:
            lr, b1, b2 = params["lr"], params.get("beta_1", 0.9), params.get("beta_2", 0.999)
            for name, g in g64.items():
                g = np.asarray(g, np.float64).reshape(self.w64[name].shape)
                tol_g = RTOL * float(np.max(np.abs(g))) if g.size else 0.0
                if params["optimizer"] == "adam":
                    m, v = self.st64["m"][name], self.st64["v"][name]
                    lr_t = o.adam_lr_t(lr, b1, b2, self.st64["iterations"])
                    sv = np.sqrt(v)
                    jac = lr_t * ((1 - b1) / (sv + o.ADAM_EPS) - m * (1 - b2) * g / (np.maximum(sv, 1e-300) * (sv + o.ADAM_EPS) ** 2))
                    if adam_mode == "lazy" and name.endswith("embeddings"):
                        jac = jac * (np.abs(g).sum(axis=1, keepdims=True) > 0)  # untouched rows do not move
                else:
                    jac = np.full(g.shape, lr)
                self.slack[name] = self.slack[name] + np.abs(jac) * tol_g
        return out, out64


def check_step_against_oracles(eng, pair, users, items, y, params, l2, mode, step, grouped, k):
    """One train step of the engine against both oracles: dense gradients, gradient tables (dense mode), loss,
    HR / DCG, updated weights."""
    negs = params["num_negs_per_pos"]
    B = len(y)
    g, g64 = pair.grads(users, items, y, l2)
    w_before, w64_before = pair.w, pair.w64
    out = eng.train_step(users, items, y, group=negs + 1, k=k, grouped=grouped).cpu().numpy().astype(np.float64)
    assert out[4] == 0
    for name, (off, shape) in eng._dense_slices.items():
        got = eng.g_dense[off:off + int(np.prod(shape))].cpu().numpy()
        close_to_truth(got, g[name], g64[name], "grad {} step {}".format(name, step), elementwise=False)
    if mode == "dense":
        for name, t in eng.g_tables.items():  # (the kernel adds the table l2 term in the update, not here)
            want = g[name] - (2.0 * l2[0]) * w_before[name] if l2[0] else g[name]
            want64 = g64[name] - (2.0 * l2[0]) * w64_before[name] if l2[0] else g64[name]
            close_to_truth(t.cpu().numpy(), want, want64, "table grad {} step {}".format(name, step), elementwise=False)
    (loss, hr, dcg), (loss64, _, _) = pair.step(users, items, y, params, adam_mode="lazy" if mode == "sparse" else "dense",
                                                g64=g64)
    got_loss = out[0] / B + out[3]
    assert abs(got_loss - loss64) <= max(4 * abs(loss - loss64), RTOL * abs(loss64)), (got_loss, loss, loss64)
    G = B // (negs + 1)
    assert abs(out[1] / G - hr) <= 1e-3 and abs(out[2] / G - dcg) <= 1e-3
    got = eng.get_weights()
    for name in pair.w:
        close_to_truth(got[name], pair.w[name], pair.w64[name], "weight {} after step {}".format(name, step + 1),
                       slack=pair.slack[name])


CONFIGS = [
    # (num_users, num_items, layers, mf_dim, negs, groups)
    (5, 10, [6, 4], 0, 3, 2),             # reference test params (test/test_model.py:8-26)
    (5, 10, [5, 4], 0, 2, 7),             # reference toy defaults, odd L0 (model.py:15-34)
    (37, 53, [7], 0, 1, 9),               # no hidden layer (model.py:175 loop empty)
    (37, 53, [9, 5, 3], 3, 4, 13),        # odd widths + GMF
    (200, 300, [64, 32, 16, 8], 0, 4, 50),   # reference trainer defaults (trainer.py:12)
    (200, 300, [64, 32, 16, 8], 8, 4, 77),   # NeuMF on the ML-1M config
    (300, 200, [256, 128, 64], 64, 4, 41),   # ML-20M tower (tensor-core path)
    (300, 200, [128, 128, 32], 0, 4, 301),   # tensor-core path, no GMF, several tiles + a ragged tail
    (100, 100, [256, 256], 16, 2, 90),       # tensor-core path, one hidden layer, N = 256
    (300, 200, [256, 128, 64], 128, 4, 37),  # BASELINE config 5 widths (embed dim 128): wide head
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "L{}f{}".format("x".join(map(str, c[2])), c[3]))
def test_forward_matches_oracle(eng_mod, cfg):
    nu, ni, L, f, negs, groups = cfg
    rng = np.random.default_rng(11)
    eng = eng_mod.NeuMFEngine(nu, ni, L, [0] * len(L), mf_dim=f, seed=3)
    w = eng.get_weights()
    for k in w:  # non-zero biases
        if k.endswith("bias"):
            w[k] = rng.normal(0, 0.1, w[k].shape).astype(np.float32)
    eng.set_weights(w)
    users, items, y = make_batch(rng, nu, ni, groups, negs)
    logits, probs, loss = eng.forward(users, items, labels=y)
    c = o.forward(w, users, items)
    rel_close(logits.cpu().numpy(), c["z"], what="logits")
    rel_close(probs.cpu().numpy(), c["p"], what="probs")
    want_loss = float(np.sum(o.bce_from_logits(c["z"].astype(np.float64), y.astype(np.float64))))
    assert abs(float(loss) - want_loss) <= RTOL * abs(want_loss)
    # float64 oracle: the kernel must be as close to the truth as the fp32 oracle is
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    z64 = o.forward(w64, users, items)["z"]
    err_gpu = np.max(np.abs(logits.cpu().numpy() - z64))
    err_cpu = np.max(np.abs(c["z"] - z64))
    assert err_gpu <= max(4 * err_cpu, 1e-6)


def test_forward_user_div_and_tail_tiles(eng_mod):
    rng = np.random.default_rng(5)
    eng = eng_mod.NeuMFEngine(30, 40, [16, 8], [0, 0], mf_dim=4, seed=2)
    w = eng.get_weights()
    for G, group in ((1, 2), (3, 100), (33, 5), (0, 4)):
        u = rng.integers(0, 30, G)
        it = rng.integers(0, 40, G * group)
        logits, _, _ = eng.forward(u, it, user_div=group)
        want = o.forward(w, np.repeat(u, group), it)["z"] if G else np.zeros(0)
        assert logits.numel() == G * group
        if G:
            rel_close(logits.cpu().numpy(), want, what="logits G={} group={}".format(G, group))


def test_gather_rows_bit_exact(eng_mod):
    rng = np.random.default_rng(0)
    for rows, dim in ((943, 32), (1682, 33), (100, 128), (17, 1)):
        table = torch.from_numpy(rng.normal(size=(rows, dim)).astype(np.float32)).cuda()
        idx = rng.integers(0, rows, 1000)
        out = eng_mod.gather_rows(table, idx)
        assert torch.equal(out, table[torch.from_numpy(idx).cuda()])
    assert eng_mod.gather_rows(table, np.zeros(0, np.int64)).shape == (0, 1)


@pytest.mark.parametrize("n,bits", [(1, 3), (31, 5), (2048, 8), (2049, 9), (100000, 17), (1310720, 18), (70000, 24)])
def test_sort_pairs_is_a_stable_sort(eng_mod, n, bits):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << bits, n).astype(np.int32)
    if n > 1000:
        keys[: n // 3] = keys[0]  # long run of one key: stability is visible
    k, i = eng_mod.sort_pairs(keys, bits)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(i.cpu().numpy(), order.astype(np.int32))
    np.testing.assert_array_equal(k.cpu().numpy(), keys[order])


TRAIN_CASES = [
    # cfg index, optimizer, table mode, l2
    (0, "adam", "dense", [0.01, 0.01]),   # the reference's test params incl. its l2
    (1, "sgd", "dense", [0.0, 0.0]),
    (2, "adam", "dense", [0.0]),
    (3, "adam", "sparse", [0.0, 0.0, 0.0]),
    (3, "sgd", "sparse", [0.0, 0.02, 0.0]),
    (4, "adam", "dense", [0, 0, 0, 0]),
    (5, "adam", "dense", [0, 0, 0, 0]),
    (5, "adam", "sparse", [0, 0, 0, 0]),
    (6, "adam", "dense", [0, 0, 0]),
    (6, "adam", "sparse", [0, 0, 0]),
    (7, "adam", "dense", [0, 0.01, 0]),
    (8, "sgd", "dense", [0, 0]),
    (9, "adam", "sparse", [0, 0, 0]),
    (9, "sgd", "dense", [0, 0, 0]),
]


@pytest.mark.parametrize("case", TRAIN_CASES, ids=lambda c: "cfg{}-{}-{}".format(c[0], c[1], c[2]))
def test_train_steps_match_oracle(eng_mod, case):
    _train_steps_vs_oracle(eng_mod, case, grouped=False)


# the grouped-batch path (user-only work once per group) on the tensor-core configs, plus configs that are not
# eligible for it (the flag must then be harmless): same oracle, same tolerances
GROUPED_CASES = [c for c in TRAIN_CASES if c[0] in (4, 6, 7, 8, 9)]


@pytest.mark.parametrize("case", GROUPED_CASES, ids=lambda c: "cfg{}-{}-{}".format(c[0], c[1], c[2]))
def test_grouped_train_steps_match_oracle(eng_mod, case):
    _train_steps_vs_oracle(eng_mod, case, grouped=True)


def _train_steps_vs_oracle(eng_mod, case, grouped):
    ci, opt, mode, l2 = case
    nu, ni, L, f, negs, groups = CONFIGS[ci]
    rng = np.random.default_rng(100 + ci)
    params = {"layers_sizes": L, "layers_l2reg": l2, "optimizer": opt, "lr": 0.001, "beta_1": 0.9, "beta_2": 0.999,
              "num_negs_per_pos": negs, "k": min(2, negs + 1)}
    eng = eng_mod.NeuMFEngine(nu, ni, L, l2, mf_dim=f, optimizer=opt, lr=0.001, table_mode=mode, seed=7)
    pair = OraclePair(eng.get_weights())
    for step in range(3):
        users, items, y = make_batch(rng, nu, ni, groups, negs)
        check_step_against_oracles(eng, pair, users, items, y, params, l2, mode, step, grouped, params["k"])
    assert eng.iterations == 3


@pytest.mark.parametrize("mode", ["dense", "sparse"])
def test_hot_rows_span_many_chunks(eng_mod, mode):
    """Few users/items and a large batch: every segment spans hundreds of 32-entry chunks, so the
    recursive segmented reduction runs all its levels (incl. whole-chunk runs and empty slots)."""
    nu, ni, L, f, negs = 50, 30, [64, 32, 16, 8], 8, 4
    rng = np.random.default_rng(21)
    groups = 8000
    users = np.repeat(np.minimum(rng.zipf(1.3, groups) - 1, nu - 1), negs + 1)
    items = np.minimum(rng.zipf(1.2, groups * (negs + 1)) - 1, ni - 1)
    y = np.tile([0] * negs + [1], groups).astype(np.float32)
    params = {"layers_sizes": L, "layers_l2reg": [0] * 4, "optimizer": "adam", "lr": 0.001, "num_negs_per_pos": negs, "k": 3}
    eng = eng_mod.NeuMFEngine(nu, ni, L, [0] * 4, mf_dim=f, table_mode=mode, seed=4)
    w = eng.get_weights()
    st = o.new_opt_state(w)
    # segments hold ~10^4 samples: the fp32 oracle's own sequential sum is only good to ~1e-4 there,
    # so the gradient tables are checked against the float64 oracle (the kernel's tree sum is tighter)
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    g64 = o.backward(w64, o.forward(w64, users, items), y.astype(np.float64))
    eng.train_step(users, items, y, group=negs + 1, k=3)
    if mode == "dense":
        for name, t in eng.g_tables.items():
            rel_close(t.cpu().numpy(), g64[name], rtol=1e-5, what="table grad " + name)
    o.train_step(w, st, users, items, y, params, adam_mode="lazy" if mode == "sparse" else "dense")
    got = eng.get_weights()
    for k in w:
        rel_close(got[k], w[k], rtol=2e-5, what="weight " + k)


def test_tensor_core_path_is_selected_and_matches_simt(eng_mod):
    """The tcgen05 path and the fp32 SIMT kernel are two implementations of the same step."""
    nu, ni, L, f, negs = 500, 400, [256, 128, 64], 64, 4
    rng = np.random.default_rng(12)
    users, items, y = make_batch(rng, nu, ni, 1000, negs)
    out, weights, logits = {}, {}, {}
    for path in ("tc", "simt"):
        eng_mod.set_compute_path(path)
        try:
            eng = eng_mod.NeuMFEngine(nu, ni, L, [0, 0, 0], mf_dim=f, seed=5)
            assert eng.uses_tensor_cores() == (path == "tc")
            logits[path] = eng.forward(users, items)[0].cpu().numpy()
            out[path] = eng.train_step(users, items, y, group=negs + 1, k=3).cpu().numpy()
            weights[path] = eng.get_weights()
        finally:
            eng_mod.set_compute_path("auto")
    rel_close(logits["tc"], logits["simt"], what="logits tc vs simt")
    assert abs(out["tc"][0] - out["simt"][0]) <= 1e-5 * abs(out["simt"][0])
    assert out["tc"][1] == out["simt"][1]
    for k in weights["tc"]:
        rel_close(weights["tc"][k], weights["simt"][k], rtol=2e-5, what="weights tc vs simt " + k)
    small = eng_mod.NeuMFEngine(5, 10, [6, 4], [0, 0], seed=1)
    assert not small.uses_tensor_cores()


def test_train_step_is_deterministic(eng_mod):
    nu, ni, L, f, negs, groups = CONFIGS[5]
    rng = np.random.default_rng(1)
    batches = [make_batch(rng, nu, ni, 400, negs) for _ in range(3)]
    results = []
    for _ in range(2):
        eng = eng_mod.NeuMFEngine(nu, ni, L, [0] * 4, mf_dim=f, seed=9)
        outs = [eng.train_step(u, i, y, group=negs + 1, k=3).cpu().numpy() for u, i, y in batches]
        results.append((eng.get_weights(), outs))
    for k in results[0][0]:
        assert np.array_equal(results[0][0][k], results[1][0][k]), k
    for a, b in zip(results[0][1], results[1][1]):
        assert np.array_equal(a, b)


def test_grouped_promise_is_verified(eng_mod):
    """grouped=True with a batch whose groups hold several users: bit 1 of the flag word is raised."""
    nu, ni, L, f, negs = 300, 200, [256, 128, 64], 64, 4
    rng = np.random.default_rng(3)
    users, items, y = make_batch(rng, nu, ni, 64, negs)
    eng = eng_mod.NeuMFEngine(nu, ni, L, [0, 0, 0], mf_dim=f, seed=2)
    assert eng.users_grouped(users, negs + 1)
    assert int(eng.train_step(users, items, y, group=negs + 1, k=3, grouped=True).cpu().numpy()[4]) == 0
    broken = users.copy()
    broken[7] = (broken[7] + 1) % nu
    assert not eng.users_grouped(broken, negs + 1)
    assert not eng.users_grouped(users[:-1], negs + 1)
    assert int(eng.train_step(broken, items, y, group=negs + 1, k=3, grouped=True).cpu().numpy()[4]) & 2
    assert int(eng.train_step(broken, items, y, group=negs + 1, k=3).cpu().numpy()[4]) == 0  # no promise, no assumption


def test_grouped_and_ungrouped_steps_agree(eng_mod):
    """Same batches through both launch sequences of the tensor-core path: Zipf-hot users and items, three
    steps, dense and sparse tables."""
    nu, ni, L, f, negs, groups = 5000, 3000, [256, 128, 64], 64, 4, 6000
    for mode in ("dense", "sparse"):
        weights = {}
        for grouped in (False, True):
            rng = np.random.default_rng(17)
            eng = eng_mod.NeuMFEngine(nu, ni, L, [0, 0, 0], mf_dim=f, table_mode=mode, optimizer="sgd", lr=0.5, seed=6)
            assert eng.uses_tensor_cores()
            outs = []
            for _ in range(3):
                users = np.repeat(np.minimum(rng.zipf(1.3, groups) - 1, nu - 1), negs + 1)
                items = np.minimum(rng.zipf(1.2, groups * (negs + 1)) - 1, ni - 1)
                y = np.tile([0] * negs + [1], groups).astype(np.float32)
                outs.append(eng.train_step(users, items, y, group=negs + 1, k=3, grouped=grouped).cpu().numpy())
            weights[grouped] = (eng.get_weights(), outs)
        for a, b in zip(weights[False][1], weights[True][1]):
            assert abs(a[0] - b[0]) <= 1e-5 * abs(a[0]) and a[1] == b[1] and a[4] == 0 and b[4] == 0
        # SGD: the update is linear in the gradients; three steps at lr 0.5 let a 1e-6 difference of the first
        # step's gradients feed back through the weights, hence 1e-4 after the third step
        for k in weights[False][0]:
            rel_close(weights[True][0][k], weights[False][0][k], rtol=1e-4, what="{} grouped vs ungrouped {}".format(mode, k))


def test_fused_and_unfused_train_steps_agree(eng_mod):
    """The fused per-tile kernel (tc_fused.cu) against the kernel-per-layer launch sequence it replaces: same batches,
    three dense-Adam steps, Zipf-hot users and items, a ragged last tile."""
    nu, ni, L, f, negs, groups = 700, 300, [256, 128, 64], 64, 4, 2077
    runs = {}
    for mode in ("auto", "off"):
        rng = np.random.default_rng(29)
        eng = eng_mod.NeuMFEngine(nu, ni, L, [0, 0, 0], mf_dim=f, table_mode="dense", optimizer="sgd", lr=0.5, seed=6,
                                  fused_train=mode)
        assert eng.uses_user_projection(groups * (negs + 1), negs + 1)
        outs = []
        for _ in range(3):
            users = np.repeat(np.minimum(rng.zipf(1.3, groups) - 1, nu - 1), negs + 1)
            items = np.minimum(rng.zipf(1.2, groups * (negs + 1)) - 1, ni - 1)
            y = np.tile([0] * negs + [1], groups).astype(np.float32)
            outs.append(eng.train_step(users, items, y, group=negs + 1, k=3, grouped=True).cpu().numpy())
        runs[mode] = (eng.get_weights(), outs)
    for a, b in zip(runs["auto"][1], runs["off"][1]):
        assert abs(a[0] - b[0]) <= 1e-5 * abs(a[0]) and a[1] == b[1] and a[4] == 0 and b[4] == 0
    for k in runs["auto"][0]:  # SGD at lr 0.5 over three steps: see test_grouped_and_ungrouped_steps_agree
        rel_close(runs["auto"][0][k], runs["off"][0][k], rtol=1e-4, what="fused vs unfused " + k)


def test_out_of_range_ids_are_flagged(eng_mod):
    eng = eng_mod.NeuMFEngine(5, 10, [6, 4], [0, 0], seed=1)
    out = eng.train_step([0, 9], [1, 2], [0.0, 1.0], group=2, k=1).cpu().numpy()
    assert out[4] != 0


# ---- item-projected first layer: E_item . W1[item rows] once per item, per-item sums of dZ1 in the backward pass ----

PROJECTED_CASES = [
    # (num_users, num_items, layers, mf_dim, negs, groups), optimizer, l2, selector
    ((300, 200, [256, 128, 64], 64, 4, 41), "adam", [0, 0, 0], "on"),        # fewer rows than 2 x items: forced
    ((300, 60, [256, 128, 64], 64, 4, 301), "adam", [0, 0.01, 0], "auto"),   # several tiles + ragged tail, Pi < 1 tile
    ((300, 200, [256, 128, 64], 128, 4, 37), "sgd", [0.001, 0, 0], "on"),    # BASELINE config 5 widths, table l2
    ((500, 130, [256, 128, 128, 32], 0, 1, 700), "adam", [0, 0, 0, 0], "auto"),  # no GMF, 4 layers, groups of 2
    ((2000, 90, [256, 128, 64], 64, 4, 301), "adam", [0, 0, 0], "auto"),     # more users than groups: items only
]


@pytest.mark.parametrize("case", PROJECTED_CASES, ids=lambda c: "ni{}-f{}-{}-{}".format(c[0][1], c[0][3], c[1], c[3]))
def test_item_projected_train_steps_match_oracle(eng_mod, case):
    (nu, ni, L, f, negs, groups), opt, l2, selector = case
    rng = np.random.default_rng(300 + ni)
    params = {"layers_sizes": L, "layers_l2reg": l2, "optimizer": opt, "lr": 0.001, "beta_1": 0.9, "beta_2": 0.999,
              "num_negs_per_pos": negs, "k": 2}
    eng_mod.set_item_projection(selector)
    try:
        eng = eng_mod.NeuMFEngine(nu, ni, L, l2, mf_dim=f, optimizer=opt, lr=0.001, table_mode="dense", seed=11)
        assert eng.uses_tensor_cores() and eng.uses_item_projection(groups * (negs + 1))
        assert eng.uses_user_projection(groups * (negs + 1), negs + 1) == (selector == "on" or nu <= groups)
        pair = OraclePair(eng.get_weights())
        for step in range(3):
            users, items, y = make_batch(rng, nu, ni, groups, negs)
            if step == 1:
                items[: len(items) // 2] = items[0]  # one hot item: a segment spanning many chunks of the reduction
            check_step_against_oracles(eng, pair, users, items, y, params, l2, "dense", step, True, 2)
    finally:
        eng_mod.set_item_projection("auto")


def test_item_projection_on_and_off_agree_and_are_deterministic(eng_mod, monkeypatch):
    """Zipf-hot users and items, 30,000 rows over 3,000 items (the automatic choice projects): the per-item launch
    sequence against the per-row one, and against itself."""
    nu, ni, L, f, negs, groups = 5000, 3000, [256, 128, 64], 64, 4, 6000
    runs = {}
    try:
        for tag, selector in (("off", "off"), ("on", "auto"), ("again", "auto")):
            eng_mod.set_item_projection(selector)
            rng = np.random.default_rng(23)
            eng = eng_mod.NeuMFEngine(nu, ni, L, [0, 0, 0], mf_dim=f, table_mode="dense", optimizer="sgd", lr=0.5, seed=6)
            assert eng.uses_item_projection(groups * (negs + 1)) == (selector != "off")
            outs = []
            for _ in range(3):
                users = np.repeat(np.minimum(rng.zipf(1.3, groups) - 1, nu - 1), negs + 1)
                items = np.minimum(rng.zipf(1.2, groups * (negs + 1)) - 1, ni - 1)
                y = np.tile([0] * negs + [1], groups).astype(np.float32)
                outs.append(eng.train_step(users, items, y, group=negs + 1, k=3, grouped=True).cpu().numpy())
            runs[tag] = (eng.get_weights(), outs)
    finally:
        eng_mod.set_item_projection("auto")
    for a, b in zip(runs["off"][1], runs["on"][1]):
        assert abs(a[0] - b[0]) <= 1e-5 * abs(a[0]) and a[1] == b[1] and a[4] == 0 and b[4] == 0
    for k in runs["off"][0]:  # SGD at lr 0.5 over three steps: see test_grouped_and_ungrouped_steps_agree
        rel_close(runs["on"][0][k], runs["off"][0][k], rtol=1e-4, what="projected vs per-row " + k)
        assert np.array_equal(runs["on"][0][k], runs["again"][0][k]), k
    for a, b in zip(runs["on"][1], runs["again"][1]):
        assert np.array_equal(a, b)


def test_item_projected_rank_eval_matches_oracle(eng_mod, monkeypatch):
    nu, ni, L, f = 300, 500, [256, 128, 64], 64
    rng = np.random.default_rng(14)
    eng = eng_mod.NeuMFEngine(nu, ni, L, [0] * 3, mf_dim=f, seed=5)
    w = eng.get_weights()
    G, group, k = 257, 100, 10
    users = rng.integers(0, nu, G)
    items = rng.integers(0, ni, G * group)
    items[5 * group:6 * group] = items[5 * group]
    assert eng.uses_item_projection(G * group)
    hr_o, dcg_o, pos_o, p_o = o.evaluate_groups(w, users, items, group, k)
    got = {}
    try:
        for selector in ("auto", "off"):
            eng_mod.set_item_projection(selector)
            eng = eng_mod.NeuMFEngine(nu, ni, L, [0] * 3, mf_dim=f, seed=5)
            assert eng.uses_item_projection(G * group) == (selector == "auto")
            # positions only: the fused sequence (a full permutation request takes the forward + rank kernels)
            pos, sums, _, probs = eng.rank_eval(users, items, group, k, want_probs=True)
            p = probs.cpu().numpy()
            rel_close(p, p_o, what="eval probs, item projection " + selector)
            np.testing.assert_array_equal(pos.cpu().numpy(), o.positive_positions(p, group))
            assert pos.cpu().numpy()[5] == group - 1
            s = sums.cpu().numpy()
            assert abs(s[0] / G - hr_o) <= 1e-3 and abs(s[1] / G - dcg_o) <= 1e-3
            got[selector] = p
    finally:
        eng_mod.set_item_projection("auto")
    rel_close(got["auto"], got["off"], what="eval probs projected vs per-row")


@pytest.mark.parametrize("f,group,L", [(32, 7, [256, 128, 32]), (128, 100, [256, 128, 64]), (64, 256, [256, 128, 128, 64]),
                                       (64, 2, [256, 128, 64])])
def test_fused_rank_eval_widths_and_group_sizes(eng_mod, f, group, L):
    """The warp-per-group score + position kernel (last layer folded into the dot): GMF widths 32 / 64 / 128, groups
    that are not multiples of four, the largest group, bad ids, a fully tied group."""
    nu, ni = 400, 90
    rng = np.random.default_rng(f + group)
    eng = eng_mod.NeuMFEngine(nu, ni, L, [0] * len(L), mf_dim=f, seed=8)
    w = eng.get_weights()
    G, k = 301, min(10, group)
    users = rng.integers(0, nu, G)
    items = rng.integers(0, ni, G * group)
    items[3 * group:4 * group] = items[3 * group]
    pos, sums, _, probs = eng.rank_eval(users, items, group, k, want_probs=True)
    p = probs.cpu().numpy()
    hr_o, dcg_o, pos_o, p_o = o.evaluate_groups(w, users, items, group, k)
    rel_close(p, p_o, what="eval probs f={} group={}".format(f, group))
    np.testing.assert_array_equal(pos.cpu().numpy(), o.positive_positions(p, group))
    assert pos.cpu().numpy()[3] == group - 1
    s = sums.cpu().numpy()
    assert abs(s[0] / G - hr_o) <= 1e-3 and abs(s[1] / G - dcg_o) <= 1e-3
    bad = items.copy()
    bad[5 * group + 1] = ni + 3
    pos_b, _, _, probs_b = eng.rank_eval(users, bad, group, k, want_probs=True)
    assert np.isnan(probs_b.cpu().numpy()[5 * group + 1])
    keep = np.arange(G) != 5
    np.testing.assert_array_equal(pos_b.cpu().numpy()[keep], pos.cpu().numpy()[keep])


def test_item_projection_treats_bad_item_ids_like_the_per_row_path(eng_mod):
    nu, ni, L, f, negs = 300, 40, [256, 128, 64], 64, 4
    rng = np.random.default_rng(3)
    users, items, y = make_batch(rng, nu, ni, 64, negs)
    items[17] = ni + 5
    eng = eng_mod.NeuMFEngine(nu, ni, L, [0, 0, 0], mf_dim=f, seed=2)
    assert eng.uses_item_projection(len(y))
    assert int(eng.train_step(users, items, y, group=negs + 1, k=3, grouped=True).cpu().numpy()[4]) & 1


# ---- the reference's default tower 64-32-16-8 (+ GMF 8): projected thread-per-group kernel (small_tower.cu) --------

SMALL_TOWER_CASES = [
    # (num_users, num_items, layers, mf_dim, negs, groups), optimizer, l2, selector
    ((60, 40, [64, 32, 16, 8], 8, 4, 301), "adam", [0, 0, 0, 0], "auto"),            # ragged last warp
    ((200, 300, [64, 32, 16, 8], 8, 4, 77), "sgd", [0.001, 0.01, 0, 0.02], "on"),   # forced; table and kernel l2
    ((700, 90, [64, 32, 16, 8], 8, 4, 2000), "adam", [0, 0, 0, 0], "auto"),          # several CTAs
]


@pytest.mark.parametrize("case", SMALL_TOWER_CASES, ids=lambda c: "nu{}-{}-{}".format(c[0][0], c[1], c[3]))
def test_small_tower_train_steps_match_oracle(eng_mod, case):
    (nu, ni, L, f, negs, groups), opt, l2, selector = case
    rng = np.random.default_rng(500 + nu)
    params = {"layers_sizes": L, "layers_l2reg": l2, "optimizer": opt, "lr": 0.001, "beta_1": 0.9, "beta_2": 0.999,
              "num_negs_per_pos": negs, "k": 2}
    eng = eng_mod.NeuMFEngine(nu, ni, L, l2, mf_dim=f, optimizer=opt, lr=0.001, table_mode="dense", seed=13,
                              item_projection=selector)
    assert not eng.uses_tensor_cores() and eng.uses_small_tower(groups * (negs + 1), negs + 1)
    pair = OraclePair(eng.get_weights())
    for step in range(3):
        users, items, y = make_batch(rng, nu, ni, groups, negs)
        if step == 1:
            items[: len(items) // 2] = items[0]  # one hot item: a segment spanning many chunks of the reduction
        check_step_against_oracles(eng, pair, users, items, y, params, l2, "dense", step, True, 2)
    assert eng.iterations == 3


def test_small_tower_and_tile_kernel_agree_and_are_deterministic(eng_mod):
    """The default tower at a size where the automatic choice projects (20,000 rows, 1,500 items, 3,000 users):
    thread-per-group kernel against the generic tile kernel, and against itself; a bad id is flagged by both."""
    nu, ni, L, f, negs, groups = 3000, 1500, [64, 32, 16, 8], 8, 4, 4000
    runs = {}
    for tag, fused in (("tile", "off"), ("small", "auto"), ("again", "auto")):
        rng = np.random.default_rng(29)
        eng = eng_mod.NeuMFEngine(nu, ni, L, [0, 0, 0, 0], mf_dim=f, table_mode="dense", optimizer="sgd", lr=0.5, seed=6,
                                  fused_train=fused)
        assert eng.uses_small_tower(groups * (negs + 1), negs + 1) == (fused != "off")
        outs = []
        for step in range(4):
            users = np.repeat(np.minimum(rng.zipf(1.3, groups) - 1, nu - 1), negs + 1)
            items = np.minimum(rng.zipf(1.2, groups * (negs + 1)) - 1, ni - 1)
            y = np.tile([0] * negs + [1], groups).astype(np.float32)
            if step == 3:
                items[4321] = ni + 7  # the last step only tests the flag (its weights are not compared)
            outs.append(eng.train_step(users, items, y, group=negs + 1, k=3, grouped=True).cpu().numpy())
            if step == 2:
                weights = eng.get_weights()
        runs[tag] = (weights, outs)
    for a, b in zip(runs["tile"][1][:3], runs["small"][1][:3]):
        assert abs(a[0] - b[0]) <= 1e-5 * abs(a[0]) and a[1] == b[1] and a[4] == 0 and b[4] == 0
    assert int(runs["tile"][1][3][4]) & 1 and int(runs["small"][1][3][4]) & 1
    for k in runs["tile"][0]:  # SGD at lr 0.5 over three steps: see test_grouped_and_ungrouped_steps_agree
        rel_close(runs["small"][0][k], runs["tile"][0][k], rtol=1e-4, what="small tower vs tile kernel " + k)
        assert np.array_equal(runs["small"][0][k], runs["again"][0][k]), k
    for a, b in zip(runs["small"][1][:3], runs["again"][1][:3]):
        assert np.array_equal(a, b)


# ---- ranking ------------------------------------------------------------------------------------

def test_rank_scores_reference_vectors(eng_mod, golden_dir):
    with open(os.path.join(golden_dir, "reference_tests.json")) as f:
        ref = json.load(f)
    for phase in ("train", "eval"):
        v = ref["rank_layer"][phase]
        rank, _, _ = eng_mod.rank_scores(np.array(v["input"], np.float32), v["negs"] + 1, 1)
        assert rank.cpu().numpy().tolist() == v["expected"]
    v = ref["ties"]
    for k in v["zero_for_k"] + [v["hit_k"]]:
        _, pos, sums = eng_mod.rank_scores(np.array(v["y_pred"], np.float32), 4, k, want_rank=False)
        assert pos.cpu().numpy().tolist() == [v["hit_position"]]
        s = sums.cpu().numpy()
        if k == v["hit_k"]:
            assert s[0] == 1.0 and abs(s[1] - np.log(2) / np.log(v["hit_position"] + 2)) < 1e-6
        else:
            assert s[0] == 0.0 and s[1] == 0.0


def test_label_column_is_argmax_of_labels(eng_mod):
    """The reference ranks the column argmax(y_true) of every group (model.py:447-451), wherever the batch puts its
    positive: the train step's HR / DCG and rank_scores(labels=...) must follow the labels, not the last column."""
    rng = np.random.default_rng(31)
    G, group, k = 200, 5, 2
    scores = rng.random(G * group).astype(np.float32)
    col = rng.integers(0, group, G)
    y = np.zeros((G, group), np.float32)
    y[np.arange(G), col] = 1.0
    _, pos, sums = eng_mod.rank_scores(scores, group, k, want_rank=False, labels=y.reshape(-1))
    rank = o.rank_groups(scores, group)
    assert abs(float(sums[0]) / G - o.hit_rate(y, rank, k)) < 1e-6
    assert abs(float(sums[1]) / G - o.discounted_cumulative_gain(y, rank, k)) < 1e-6
    _, pos_lc, _ = eng_mod.rank_scores(scores, group, k, want_rank=False, label_col=col)
    assert np.array_equal(pos.cpu().numpy(), pos_lc.cpu().numpy())
    # the train step: same metrics whether the positive is last or anywhere else
    nu, ni, L, f = 40, 60, [64, 32, 16, 8], 8
    params = {"layers_sizes": L, "layers_l2reg": [0] * 4, "optimizer": "sgd", "lr": 0.01, "num_negs_per_pos": group - 1, "k": k}
    eng = eng_mod.NeuMFEngine(nu, ni, L, [0] * 4, mf_dim=f, optimizer="sgd", lr=0.01, seed=3)
    w = eng.get_weights()
    users = np.repeat(rng.integers(0, nu, G), group)
    items = rng.integers(0, ni, G * group)
    out = eng.train_step(users, items, y.reshape(-1), group=group, k=k).cpu().numpy()
    loss, hr, dcg = o.train_step(w, o.new_opt_state(w), users, items, y.reshape(-1), params)
    assert abs(out[1] / G - hr) < 1e-6 and abs(out[2] / G - dcg) < 1e-5, (out[1] / G, hr, out[2] / G, dcg)


def test_rank_scores_match_reference_execution(eng_mod, golden_dir):
    g = np.load(os.path.join(golden_dir, "rank_metrics.npz"))
    for c in range(int(g["num_cases"])):
        pre = "c{}_".format(c)
        s = g[pre + "scores"]
        G, group = s.shape
        for k, hr, dcg in zip(g[pre + "ks"], g[pre + "hr"], g[pre + "dcg"]):
            rank, pos, sums = eng_mod.rank_scores(s, group, int(k))
            np.testing.assert_array_equal(rank.cpu().numpy(), g[pre + "rank"])
            np.testing.assert_array_equal(pos.cpu().numpy(), g[pre + "pos"])
            sm = sums.cpu().numpy()
            assert abs(sm[0] / G - hr) <= 1e-6 and abs(sm[1] / G - dcg) <= 1e-5


def test_rank_scores_nan_and_label_col(eng_mod):
    s = np.array([[0.3, np.nan, 0.5, 0.4], [0.3, 0.1, 0.2, np.nan]], np.float32)
    rank, pos, _ = eng_mod.rank_scores(s, 4, 2)
    assert rank.cpu().numpy().tolist() == o.rank_groups(s, 4).tolist()
    assert pos.cpu().numpy().tolist() == o.positive_positions(s, 4).tolist()
    _, pos, _ = eng_mod.rank_scores(s, 4, 2, label_col=np.array([0, 2], np.int32))
    assert pos.cpu().numpy().tolist() == [2, 1]


@pytest.mark.parametrize("fused", ["auto", "off"], ids=["projected-thread-per-row", "tile-kernel"])
def test_rank_eval_matches_oracle(eng_mod, fused):
    """The default tower's eval: 25,700 rows over 300 + 500 table rows, so the automatic choice projects the first layer
    over the tables and runs the thread-per-row forward (small_tower.cu); fused_train='off' keeps the tile kernel."""
    nu, ni, L, f = 300, 500, [64, 32, 16, 8], 8
    rng = np.random.default_rng(4)
    eng = eng_mod.NeuMFEngine(nu, ni, L, [0] * 4, mf_dim=f, seed=5, fused_train=fused)
    w = eng.get_weights()
    G, group, k = 257, 100, 10
    users = rng.integers(0, nu, G)
    items = rng.integers(0, ni, G * group)
    items[5 * group:6 * group] = items[5 * group]  # a fully tied group: positive must lose
    pos, sums, rank, probs = eng.rank_eval(users, items, group, k, want_rank=True, want_probs=True)
    p = probs.cpu().numpy()
    hr_o, dcg_o, pos_o, p_o = o.evaluate_groups(w, users, items, group, k)
    rel_close(p, p_o, what="eval probs")
    # bit-exact given equal scores: rank the kernel's own scores with the oracle's rule
    np.testing.assert_array_equal(pos.cpu().numpy(), o.positive_positions(p, group))
    np.testing.assert_array_equal(rank.cpu().numpy(), o.rank_groups(p, group))
    assert pos.cpu().numpy()[5] == group - 1
    s = sums.cpu().numpy()
    assert abs(s[0] / G - hr_o) <= 1e-3 and abs(s[1] / G - dcg_o) <= 1e-3
    hs, ds = o.metrics_from_positions(pos.cpu().numpy(), k)
    assert s[0] == hs and abs(s[1] - ds) <= 1e-4 * max(ds, 1)


@pytest.mark.parametrize("L,f", [([256, 128, 64], 64), ([64, 32, 16, 8], 8)], ids=["ml20m-tower", "default-tower"])
def test_rank_eval_from_host_arrays_is_pipelined_and_identical(eng_mod, monkeypatch, L, f):
    """Host inputs above EVAL_PIPELINE_MIN_ROWS are uploaded in chunks under the sweep: same positions, same sums (to
    the order of the chunks' partial sums) as one call on device tensors; the id check sees a bad id in any chunk."""
    nu, ni, G, group, k = 500, 300, 1003, 100, 10
    rng = np.random.default_rng(12)
    eng = eng_mod.NeuMFEngine(nu, ni, L, [0] * len(L), mf_dim=f, seed=9)
    users = rng.integers(0, nu, G).astype(np.int32)
    items = rng.integers(0, ni, G * group).astype(np.int32)
    pos_d, sums_d, _, _ = eng.rank_eval(torch.from_numpy(users).cuda(), torch.from_numpy(items).cuda(), group, k)
    monkeypatch.setattr(eng_mod.NeuMFEngine, "EVAL_PIPELINE_MIN_ROWS", 1000)
    for host_u, host_i in ((users, items), (torch.from_numpy(users).pin_memory(), torch.from_numpy(items).pin_memory()),
                           (users.astype(np.int64), items.astype(np.int64))):
        pos_h, sums_h, _, _ = eng.rank_eval(host_u, host_i, group, k, check_ids=True)
        np.testing.assert_array_equal(pos_h.cpu().numpy(), pos_d.cpu().numpy())
        assert sums_h.cpu().numpy()[0] == sums_d.cpu().numpy()[0]
        assert abs(sums_h.cpu().numpy()[1] - sums_d.cpu().numpy()[1]) <= 1e-5 * abs(sums_d.cpu().numpy()[1])
        assert float(eng.last_eval_bad.item()) == 0.0
    bad = items.copy()
    bad[-7] = ni + 1  # in the last chunk
    eng.rank_eval(users, bad, group, k, check_ids=True)
    assert float(eng.last_eval_bad.item()) != 0.0


# ---- sampler --------------------------------------------------------------------------------------

def test_sampler_bit_exact_vs_oracle(eng_mod):
    rng = np.random.default_rng(8)
    nu, ni = 40, 60
    users = np.repeat(np.arange(nu), 12)
    items = np.concatenate([rng.choice(ni, 12, replace=False) for _ in range(nu)])
    users = np.concatenate([users, np.full(58, 3)])  # user 3 has 2 candidates left -> with replacement
    items = np.concatenate([items, np.setdiff1d(np.arange(ni), [7, 11])[:58]])
    rowptr, csr = o.build_csr(nu, users, items)
    pu = rng.integers(0, nu, 64).astype(np.int32)
    pu[:4] = 3
    pi = rng.integers(0, ni, 64).astype(np.int32)
    d_rowptr, d_csr = torch.from_numpy(rowptr).cuda(), torch.from_numpy(csr).cuda()
    for negs, seed, epoch, first in ((4, 1, 0, 0), (9, 2 ** 40 + 5, 3, 640), (20, 7, 2 ** 33, 2 ** 32 + 1)):
        xu, xi, y = eng_mod.sample_negatives(d_rowptr, d_csr, ni, pu, pi, first, negs, seed, epoch)
        want = o.device_sample_batch(rowptr, csr, ni, pu, pi, first, negs, seed, epoch)
        np.testing.assert_array_equal(xi.cpu().numpy(), want)
        np.testing.assert_array_equal(xu.cpu().numpy(), np.repeat(pu, negs + 1))
        np.testing.assert_array_equal(y.cpu().numpy(), np.tile([0] * negs + [1], 64))
        got = xi.cpu().numpy().reshape(64, negs + 1)
        for p in range(64):
            seen = set(csr[rowptr[pu[p]]:rowptr[pu[p] + 1]].tolist())
            assert not seen & set(got[p, :negs].tolist())
            if ni - len(seen) >= negs:
                assert len(set(got[p, :negs].tolist())) == negs


def test_sampler_ignores_seen_items_beyond_num_items(eng_mod):
    """Candidates are arange(num_items) minus the seen items (np.setdiff1d, data_pipeline.py:104-108): seen ids at or
    beyond num_items -- raw id spaces, a num_items smaller than the lists' range -- take no candidate away, so every
    valid unseen item can still be drawn and the draws equal the oracle's bit for bit."""
    rng = np.random.default_rng(9)
    nu, ni_lists, ni = 30, 90, 50  # the lists hold ids up to 89, the sampler is asked for items below 50
    users = np.repeat(np.arange(nu), 20)
    items = np.concatenate([rng.choice(ni_lists, 20, replace=False) for _ in range(nu)])
    rowptr, csr = o.build_csr(nu, users, items)
    pu = rng.integers(0, nu, 256).astype(np.int32)
    pi = rng.integers(0, ni, 256).astype(np.int32)
    d_rowptr, d_csr = torch.from_numpy(rowptr).cuda(), torch.from_numpy(csr).cuda()
    drawn = {u: set() for u in range(nu)}
    for epoch in range(40):
        xu, xi, _ = eng_mod.sample_negatives(d_rowptr, d_csr, ni, pu, pi, 0, 4, 5, epoch)
        if epoch < 3:
            np.testing.assert_array_equal(xi.cpu().numpy(), o.device_sample_batch(rowptr, csr, ni, pu, pi, 0, 4, 5, epoch))
        got = xi.cpu().numpy().reshape(256, 5)
        assert got[:, :4].max() < ni
        for p in range(256):
            drawn[int(pu[p])].update(got[p, :4].tolist())
    for u in set(pu.tolist()):
        seen = set(csr[rowptr[u]:rowptr[u + 1]].tolist())
        cand = set(range(ni)) - seen
        assert not drawn[u] & seen
        if sum(pu == u) * 40 * 4 >= 40 * len(cand):  # enough draws that every candidate is all but certain to appear
            assert drawn[u] == cand, (u, sorted(cand - drawn[u]))


# ---- dataset preparation on the device: split and per-user item lists ---------------------------

def _ratings(rng, nu, ni, n, min_per_user=2):
    users = np.concatenate([np.repeat(np.arange(nu), min_per_user), rng.integers(0, nu, n - nu * min_per_user)])
    rng.shuffle(users)
    return users.astype(np.int32), rng.integers(0, ni, len(users)).astype(np.int32)


@pytest.mark.parametrize("nu,ni,n", [(3, 5, 9), (40, 60, 1000), (943, 1682, 100000), (6040, 3706, 1000209)])
def test_device_split_matches_oracle(eng_mod, nu, ni, n):
    rng = np.random.default_rng(n)
    users, _ = _ratings(rng, nu, ni, n)
    order, part = eng_mod.split_last_two(users, nu)
    order, part = order.cpu().numpy(), part.cpu().numpy()
    np.testing.assert_array_equal(order, np.argsort(users, kind="stable"))  # bit-exact: a stable sort by user
    tr, va, te = o.leave_last_two_out(users)  # pinned by the reference's own known-answer test (test_oracle.py)
    np.testing.assert_array_equal(order[part == 0], tr)
    np.testing.assert_array_equal(order[part == 1], va)
    np.testing.assert_array_equal(order[part == 2], te)


def test_device_split_edge_cases(eng_mod):
    # a user with one rating has a test row only; empty input; out-of-range ids are refused
    order, part = eng_mod.split_last_two(np.array([2, 0, 2, 2, 1, 0], np.int32), 3)
    assert order.cpu().numpy().tolist() == [1, 5, 4, 0, 2, 3] and part.cpu().numpy().tolist() == [1, 2, 2, 0, 1, 2]
    order, part = eng_mod.split_last_two(np.zeros(0, np.int32), 3)
    assert order.numel() == 0 and part.numel() == 0
    with pytest.raises(IndexError):
        eng_mod.split_last_two(np.array([0, 3], np.int32), 3)


@pytest.mark.parametrize("nu,ni,n", [(3, 5, 9), (40, 60, 5000), (943, 1682, 100000), (6040, 3706, 1000209)])
def test_device_user_csr_matches_oracle(eng_mod, nu, ni, n):
    rng = np.random.default_rng(n + 1)
    users, items = _ratings(rng, nu, ni, n)  # dense enough for many duplicate pairs at the small sizes
    rowptr, csr = eng_mod.build_user_csr(users, items, nu, ni)
    want_rowptr, want_csr = o.build_csr(nu, users, items)
    np.testing.assert_array_equal(rowptr.cpu().numpy(), want_rowptr)
    np.testing.assert_array_equal(csr.cpu().numpy(), want_csr)


def test_device_user_csr_edge_cases(eng_mod):
    # users without ratings (empty rows at both ends and in the middle), duplicates, empty input, bad ids
    users = np.array([5, 2, 5, 5, 2, 7], np.int32)
    items = np.array([9, 1, 3, 9, 1, 0], np.int32)
    rowptr, csr = eng_mod.build_user_csr(users, items, 10, 10)
    assert rowptr.cpu().numpy().tolist() == [0, 0, 0, 1, 1, 1, 3, 3, 4, 4, 4]
    assert csr.cpu().numpy().tolist() == [1, 3, 9, 0]
    rowptr, csr = eng_mod.build_user_csr(np.zeros(0, np.int32), np.zeros(0, np.int32), 4, 4)
    assert rowptr.cpu().numpy().tolist() == [0] * 5 and csr.numel() == 0
    with pytest.raises(IndexError):
        eng_mod.build_user_csr(np.array([0, 1], np.int32), np.array([0, 4], np.int32), 2, 4)


@pytest.mark.parametrize("n,limit", [(1, 5), (9, 4), (5000, 70000), (1000209, 3953), (100000, 2 ** 30)])
def test_device_remap_ids_matches_numpy_unique(eng_mod, n, limit):
    rng = np.random.default_rng(n)
    ids = (rng.integers(0, min(limit, 200000), n) * max(1, limit // 200000)).astype(np.int32)  # sparse, with repeats
    ids = np.minimum(ids, limit - 1)
    dense, unique = eng_mod.remap_ids(ids, limit)
    want_unique, want_dense = np.unique(ids, return_inverse=True)
    np.testing.assert_array_equal(unique.cpu().numpy(), want_unique)
    np.testing.assert_array_equal(dense.cpu().numpy(), want_dense.reshape(-1))
    d2, u2 = eng_mod.remap_ids(ids)  # limit from the data
    assert torch.equal(d2, dense) and torch.equal(u2, unique)
    with pytest.raises(IndexError):
        eng_mod.remap_ids(np.array([0, limit], np.int64).astype(np.int32) if limit < 2 ** 30 else np.array([-1], np.int32), limit)
    d0, u0 = eng_mod.remap_ids(np.zeros(0, np.int32), 7)
    assert d0.numel() == 0 and u0.numel() == 0


def test_generator_uses_the_device_csr(eng_mod):
    """The generator's table is the device-built one and equals the NumPy statement of the same table."""
    import pandas as pd
    from movierec import data_pipeline
    rng = np.random.default_rng(5)
    users, items = _ratings(rng, 50, 80, 2000, min_per_user=3)
    df = pd.DataFrame({"userId": users, "itemId": items, "rating": np.ones(len(users), np.float32)})
    gen = data_pipeline.MovieLensDataGenerator("ml-100k", df, batch_size=50, negatives_per_positive=4, shuffle=False)
    gen.device_batch(0)
    rowptr, csr = data_pipeline.build_user_csr(users, items)
    np.testing.assert_array_equal(gen._device["rowptr"].cpu().numpy(), rowptr)
    np.testing.assert_array_equal(gen._device["csr"].cpu().numpy(), csr)
    train, validation, test = data_pipeline.split_leave_last_two_out(df)             # on the device
    train_h, validation_h, test_h = data_pipeline.split_leave_last_two_out(df, on_device=False)
    for a, b in ((train, train_h), (validation, validation_h), (test, test_h)):
        pd.testing.assert_frame_equal(a, b)


# ---- BASELINE sizes: size-independent properties --------------------------------------------------

def test_full_size_ml20m_properties(eng_mod):
    """ML-20M shape, 1.3M-row batch: determinism, loss = sum of row losses, dense grads are the
    sum of two half-batch grads (linearity), untouched rows do not move on step 1."""
    nu, ni, L, f, negs = 138493, 26744, [256, 128, 64], 64, 4
    B = 5 * 2 ** 16  # a quarter of the bench batch keeps the test short; tiles, sort and reduce all scale
    rng = np.random.default_rng(0)
    users, items, y = make_batch(rng, nu, ni, B // (negs + 1), negs)
    eng = eng_mod.NeuMFEngine(nu, ni, L, [0, 0, 0], mf_dim=f, seed=1)
    w0_user = eng.user_mlp.clone()
    out_full = eng.train_grads(users, items, y, group=negs + 1, k=3, inv_global_batch=1.0 / B)
    g_full = eng.g_dense.clone()
    gt_full = eng.g_tables["item_embedding/embeddings"].clone()
    h = B // 2
    eng.train_grads(users[:h], items[:h], y[:h], inv_global_batch=1.0 / B)
    g_a, gt_a = eng.g_dense.clone(), eng.g_tables["item_embedding/embeddings"].clone()
    eng.train_grads(users[h:], items[h:], y[h:], inv_global_batch=1.0 / B)
    g_b, gt_b = eng.g_dense.clone(), eng.g_tables["item_embedding/embeddings"].clone()
    rel_close((g_a + g_b).cpu().numpy(), g_full.cpu().numpy(), rtol=1e-4, what="dense grad linearity")
    rel_close((gt_a + gt_b).cpu().numpy(), gt_full.cpu().numpy(), rtol=1e-4, what="item grad linearity")
    # sampled rows against the oracle forward on the same weights
    w = eng.get_weights()
    sel = rng.choice(B, 4096, replace=False)
    logits, _, loss = eng.forward(users, items, labels=y)
    c = o.forward(w, users[sel], items[sel])
    rel_close(logits.cpu().numpy()[sel], c["z"], what="logits sample")
    assert abs(float(loss) - float(out_full[0])) <= 1e-6 * abs(float(loss))
    # determinism + untouched rows
    eng.train_grads(users, items, y, group=negs + 1, k=3, inv_global_batch=1.0 / B)
    assert torch.equal(eng.g_dense, g_full) and torch.equal(eng.g_tables["item_embedding/embeddings"], gt_full)
    eng.apply()
    touched = torch.zeros(nu, dtype=torch.bool, device="cuda")
    touched[torch.from_numpy(users).cuda()] = True
    assert torch.equal(eng.user_mlp[~touched], w0_user[~touched])
    assert not torch.equal(eng.user_mlp[touched], w0_user[touched])


def test_bench_sequence_ml20m_matches_oracle(eng_mod):
    """The exact launch sequence bench.py times -- ML-20M tables (BASELINE configs[2]), 1,310,720 rows, grouped batch,
    item- and user-projected first layer, the fused per-tile kernel, dense Adam -- one step from injected weights
    against the fp32 and the float64 oracle: loss, every dense gradient, all four gradient tables, updated weights."""
    nu, ni, L, f, negs = 138493, 26744, [256, 128, 64], 64, 4
    B = 5 * 2 ** 18
    groups = B // (negs + 1)
    rng = np.random.default_rng(42)
    # bench.py's batch shape: lognormal user activity, Zipf positives over a fixed permutation, uniform negatives
    act = 20.0 + rng.lognormal(3.0, 1.0, nu)
    u = np.minimum(np.searchsorted(np.cumsum(act / act.sum()), rng.random(groups)), nu - 1).astype(np.int32)
    zipf = 1.0 / (np.arange(ni) + 1.0)
    perm = rng.permutation(ni)
    pos = perm[np.minimum(np.searchsorted(np.cumsum(zipf / zipf.sum()), rng.random(groups)), ni - 1)].astype(np.int32)
    items = rng.integers(0, ni, (groups, negs + 1), dtype=np.int32)
    items[:, -1] = pos
    users, items = np.repeat(u, negs + 1), items.reshape(-1)
    y = np.tile(np.array([0] * negs + [1], np.float32), groups)
    params = {"layers_sizes": L, "layers_l2reg": [0, 0, 0], "optimizer": "adam", "lr": 0.001, "beta_1": 0.9,
              "beta_2": 0.999, "num_negs_per_pos": negs, "k": negs + 1}
    eng = eng_mod.NeuMFEngine(nu, ni, L, [0, 0, 0], mf_dim=f, optimizer="adam", lr=0.001, table_mode="dense", seed=1)
    assert eng.uses_tensor_cores() and eng.uses_item_projection(B) and eng.uses_user_projection(B, negs + 1)
    w = eng.get_weights()
    for name in w:  # injected weights: non-zero biases, so that every term of the step is exercised
        if name.endswith("bias"):
            w[name] = rng.normal(0, 0.05, w[name].shape).astype(np.float32)
    eng.set_weights(w)
    pair = OraclePair(w)
    check_step_against_oracles(eng, pair, users, items, y, params, [0, 0, 0], "dense", 0, True, negs + 1)


def test_full_size_eval_sweep_properties(eng_mod):
    """Config 4 shape: every ML-20M-shaped user, 1 positive + 99 negatives."""
    nu, ni, L, f = 138493, 26744, [256, 128, 64], 64
    group, k = 100, 10
    rng = np.random.default_rng(2)
    eng = eng_mod.NeuMFEngine(nu, ni, L, [0, 0, 0], mf_dim=f, seed=3)
    users = np.arange(nu, dtype=np.int32)
    items = rng.integers(0, ni, nu * group).astype(np.int32)
    pos, sums, _, probs = eng.rank_eval(users, items, group, k, want_probs=True)
    p = probs.cpu().numpy()
    np.testing.assert_array_equal(pos.cpu().numpy(), o.positive_positions(p, group))
    hs, ds = o.metrics_from_positions(pos.cpu().numpy(), k)
    s = sums.cpu().numpy()
    assert s[0] == hs and abs(s[1] - ds) <= 1e-5 * ds
    # a permutation of the negatives inside each group leaves the position unchanged when scores are distinct
    sel = rng.choice(nu, 2000, replace=False)
    it = items.reshape(nu, group)[sel].copy()
    it[:, :-1] = it[:, :-1][:, ::-1]
    pos2, _, _, probs2 = eng.rank_eval(users[sel], it.reshape(-1), group, k, want_probs=True)
    p2 = probs2.cpu().numpy().reshape(-1, group)
    distinct = np.array([len(np.unique(r)) == group for r in p2])
    assert np.array_equal(pos2.cpu().numpy()[distinct], pos.cpu().numpy()[sel][distinct])
    w = eng.get_weights()
    c = o.forward(w, np.repeat(users[sel[:64]], group), items.reshape(nu, group)[sel[:64]].reshape(-1))
    rel_close(p.reshape(nu, group)[sel[:64]].reshape(-1), c["p"], what="eval probs sample")
