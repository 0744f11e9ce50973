"""The oracle against the reference's known-answer vectors and against independent checks
(float64 finite differences, torch autograd).  CPU only."""

import json
import math
import os

import numpy as np
import pytest

from oracle import movierec_oracle as o


@pytest.fixture(scope="module")
def ref(golden_dir):
    with open(os.path.join(golden_dir, "reference_tests.json")) as f:
        return json.load(f)


# ---- rank layer + metrics: reference test vectors (test/test_model.py) -----------------------

def test_hit_rate_known_answers(ref):
    v = ref["hit_rate"]
    for k, want in v["k_to_hr"].items():
        assert o.hit_rate(np.array(v["y_true"]), np.array(v["rank"], np.int32), int(k)) == pytest.approx(want)


def test_dcg_known_answers(ref):
    v = ref["dcg"]
    rank = o.rank_groups(np.array(v["y_pred"], np.float32), 4)
    _, pos = o.hits_per_user(np.array(v["y_true"]), rank, 4)
    assert pos.tolist() == v["positions"]
    assert o.positive_positions(np.array(v["y_pred"], np.float32), 4).tolist() == v["positions"]
    for k, hits in v["k_to_hits"].items():
        want = sum(h * math.log(2) / math.log(p + 2) for h, p in zip(hits, v["positions"])) / 2.0
        got = o.discounted_cumulative_gain(np.array(v["y_true"]), rank, int(k))
        assert got == pytest.approx(want, abs=1e-6)


def test_ties_positive_loses(ref):
    v = ref["ties"]
    s = np.array(v["y_pred"], np.float32)
    rank = o.rank_groups(s, 4)
    for k in v["zero_for_k"]:
        assert o.hit_rate(np.array(v["y_true"]), rank, k) == 0.0
        assert o.discounted_cumulative_gain(np.array(v["y_true"]), rank, k) == 0.0
    assert o.hit_rate(np.array(v["y_true"]), rank, v["hit_k"]) == pytest.approx(1.0)
    want = math.log(2) / math.log(v["hit_position"] + 2)
    assert o.discounted_cumulative_gain(np.array(v["y_true"]), rank, v["hit_k"]) == pytest.approx(want, abs=1e-6)
    assert o.positive_positions(s, 4).tolist() == [v["hit_position"]]


def test_rank_layer_permutations(ref):
    for phase in ("train", "eval"):
        v = ref["rank_layer"][phase]
        got = o.rank_groups(np.array(v["input"], np.float32), v["negs"] + 1)
        assert got.tolist() == v["expected"]


def test_rank_metrics_match_reference_execution(golden_dir):
    """Vectors produced by running the reference's own metric source (make_golden.py)."""
    g = np.load(os.path.join(golden_dir, "rank_metrics.npz"))
    for c in range(int(g["num_cases"])):
        pre = "c{}_".format(c)
        s, rank = g[pre + "scores"], g[pre + "rank"]
        group = s.shape[1]
        np.testing.assert_array_equal(o.rank_groups(s, group), rank)
        np.testing.assert_array_equal(o.positive_positions(s, group), g[pre + "pos"])
        y = np.zeros(s.shape, np.int64)
        y[:, -1] = 1
        for k, hr, dcg in zip(g[pre + "ks"], g[pre + "hr"], g[pre + "dcg"]):
            assert o.hit_rate(y, rank, int(k)) == pytest.approx(hr, abs=1e-7)
            assert o.discounted_cumulative_gain(y, rank, int(k)) == pytest.approx(dcg, abs=1e-6)
            hs, ds = o.metrics_from_positions(g[pre + "pos"], int(k))
            assert hs / len(s) == pytest.approx(hr, abs=1e-7)
            assert ds / len(s) == pytest.approx(dcg, abs=1e-6)


def test_nan_scores_rank_last():
    s = np.array([[0.3, np.nan, 0.5, 0.4]], np.float32)
    assert o.rank_groups(s, 4).tolist() == [[2, 3, 0, 1]]
    assert o.positive_positions(s, 4).tolist() == [1]
    assert o.positive_positions(np.array([[0.3, 0.1, np.nan]], np.float32), 3).tolist() == [2]


# ---- generator / sampler / split: reference vectors (test/test_data_pipeline.py) -------------

def test_split_known_answer(ref):
    v = ref["split"]
    tr, va, te = o.leave_last_two_out(np.array(v["userId"]))
    items, rating = np.array(v["itemId"]), np.array(v["rating"])
    for idx, name in ((tr, "train"), (va, "validation"), (te, "test")):
        assert np.array(v["userId"])[idx].tolist() == v[name]["userId"]
        assert items[idx].tolist() == v[name]["itemId"]
        assert rating[idx].tolist() == v[name]["rating"]


def test_generator_forced_outcomes(ref):
    v = ref["generator_get_item"]
    du, di = np.array(v["data"]["userId"]), np.array(v["data"]["itemId"])
    eu, ei = np.array(v["extra"]["userId"]), np.array(v["extra"]["itemId"])
    rng = np.random.RandomState(0)
    for _ in range(10):
        (xu, xi), y = o.reference_batch(du, di, np.arange(4), 0, v["batch_size"], v["negs"], v["num_items"],
                                        rng, eu, ei)
        assert xu.tolist() == v["batch0"]["users"]
        assert xi.tolist() == v["batch0"]["items"]
        assert y.tolist() == v["batch0"]["y"]
        (xu, xi), y = o.reference_batch(du, di, np.arange(4), 1, v["batch_size"], v["negs"], v["num_items"],
                                        rng, eu, ei)
        assert xu.tolist() == v["batch1"]["users"]
        assert xi[:3].tolist() == v["batch1"]["items_first3"]
        assert len(np.setdiff1d(v["batch1"]["user1_candidates"], xi[3:5])) == 1
        assert xi[5] == v["batch1"]["last_item"]


def test_generator_matches_reference_execution(golden_dir):
    """Bit-exact replay of batches drawn from the reference's own generator under np.random.seed."""
    g = np.load(os.path.join(golden_dir, "generator_batches.npz"))
    for tag, use_extra in (("noextra", False), ("extra", True)):
        rng = np.random.RandomState(int(g["seed"]))
        idx = np.arange(len(g["data_users"]))
        rng.shuffle(idx)  # on_epoch_end in the constructor (data_pipeline.py:71,152-154)
        np.testing.assert_array_equal(idx, g[tag + "_indexes"])
        negs, bs = int(g[tag + "_negs"]), int(g[tag + "_bs"])
        assert o.generator_len(len(idx), bs) == int(g[tag + "_len"])
        for b in range(3):
            (xu, xi), y = o.reference_batch(
                g["data_users"], g["data_items"], idx, b, bs, negs, int(g["num_items"]), rng,
                g["extra_users"] if use_extra else None, g["extra_items"] if use_extra else None)
            np.testing.assert_array_equal(xu, g["{}_b{}_users".format(tag, b)])
            np.testing.assert_array_equal(xi, g["{}_b{}_items".format(tag, b)])
            np.testing.assert_array_equal(y, g["{}_b{}_y".format(tag, b)])


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = o.philox4x32_10(np.array(c, np.uint64), np.array(k, np.uint64))
        assert tuple(int(x) for x in got) == want


def test_device_sampler_constraints(ref):
    v = ref["generator_duplicated_user"]
    rowptr, csr = o.build_csr(2, v["data"]["userId"], v["data"]["itemId"])
    seen_sets = set()
    for p in range(50):
        g = o.device_sample_group(rowptr, csr, v["num_items"], 0, p, v["negs"], seed=3, epoch=0)
        assert len(np.setdiff1d(v["candidates"], g)) == 1  # 2 distinct picks out of 3 candidates
        seen_sets.add(tuple(g.tolist()))
    assert len(seen_sets) > 1
    # forced outcome: a single candidate -> sampling with replacement returns it every time
    v = ref["generator_get_item"]
    users = v["data"]["userId"] + v["extra"]["userId"]
    items = v["data"]["itemId"] + v["extra"]["itemId"]
    rowptr, csr = o.build_csr(3, users, items)
    for u in (0, 2):
        assert o.device_sample_group(rowptr, csr, v["num_items"], u, 5, 2, seed=1, epoch=2).tolist() == [3, 3]


def test_device_sampler_uniform_without_replacement():
    rowptr, csr = o.build_csr(1, [0] * 4, [1, 3, 4, 8])
    counts = np.zeros(10)
    for p in range(4000):
        g = o.device_sample_group(rowptr, csr, 10, 0, p, 3, seed=11, epoch=1)
        assert len(set(g.tolist())) == 3 and not set(g.tolist()) & {1, 3, 4, 8}
        counts[g] += 1
    cand = counts[[0, 2, 5, 6, 7, 9]]
    assert counts[[1, 3, 4, 8]].sum() == 0
    assert np.all(np.abs(cand / cand.sum() - 1 / 6) < 0.02)


# ---- float path (parity unpinned by the reference): independent cross-checks -----------------

PARAMS = {"num_users": 7, "num_items": 11, "layers_sizes": [6, 5, 4], "layers_l2reg": [0.01, 0.02, 0.0],
          "optimizer": "adam", "lr": 0.01, "num_negs_per_pos": 3, "k": 2}


def _batch(rng, B, nu, ni, negs):
    users = np.repeat(rng.integers(0, nu, B // (negs + 1)), negs + 1)
    items = rng.integers(0, ni, B)
    y = np.tile([0] * negs + [1], B // (negs + 1))
    return users, items, y


@pytest.mark.parametrize("mf_dim", [0, 3])
def test_backward_matches_finite_differences(mf_dim):
    rng = np.random.default_rng(0)
    w = o.init_weights(7, 11, [6, 5, 4], mf_dim, rng, dtype=np.float64)
    for k in w:  # non-zero biases so every path is exercised
        if k.endswith("bias"):
            w[k] = rng.normal(0, 0.1, w[k].shape)
    users, items, y = _batch(rng, 16, 7, 11, 3)
    l2 = PARAMS["layers_l2reg"]
    g = o.backward(w, o.forward(w, users, items), y, None, l2)
    for k in w:
        flat = w[k].reshape(-1)
        for j in rng.choice(flat.size, min(flat.size, 6), replace=False):
            old = flat[j]
            flat[j] = old + 1e-6
            lp = o.loss_value(w, users, items, y, l2)
            flat[j] = old - 1e-6
            lm = o.loss_value(w, users, items, y, l2)
            flat[j] = old
            assert g[k].reshape(-1)[j] == pytest.approx((lp - lm) / 2e-6, rel=1e-5, abs=1e-8), k


def test_forward_backward_match_torch_autograd():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(3)
    w = o.init_weights(9, 13, [8, 6, 4], 2, rng)
    users, items, y = _batch(rng, 24, 9, 13, 5)
    c = o.forward(w, users, items)
    g = o.backward(w, c, y)
    tw = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in w.items()}
    tu, ti = torch.tensor(users), torch.tensor(items)
    x = torch.cat([tw[o.K_USER][tu], tw[o.K_ITEM][ti]], 1)
    for i in (1, 2):
        x = torch.relu(x @ tw["hidden_%d/kernel" % i] + tw["hidden_%d/bias" % i])
    h = torch.cat([tw[o.K_GMF_USER][tu] * tw[o.K_GMF_ITEM][ti], x], 1)
    z = (h @ tw[o.K_OUT_W] + tw[o.K_OUT_B]).reshape(-1)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(z, torch.tensor(y, dtype=torch.float64))
    loss.backward()
    np.testing.assert_allclose(c["z"], z.detach().numpy(), rtol=2e-5, atol=1e-6)
    for k in w:
        np.testing.assert_allclose(g[k].reshape(w[k].shape), tw[k].grad.numpy(), rtol=2e-4, atol=1e-7, err_msg=k)


def test_bce_forms_agree_away_from_saturation():
    z = np.linspace(-8, 8, 101).astype(np.float32)
    for y in (0.0, 1.0):
        a = o.bce_from_logits(z, np.float32(y))
        b = o.bce_from_probs(o.sigmoid(z), np.float32(y))
        np.testing.assert_allclose(a, b, rtol=5e-4, atol=1e-6)


def test_adam_is_keras_legacy_not_torch():
    """eps sits outside the bias-corrected sqrt: p -= lr_t * m / (sqrt(v) + eps)."""
    w = {"x": np.array([1.0, -2.0], np.float64)}
    st = o.new_opt_state(w)
    g = {"x": np.array([0.5, 1e-9])}
    o.adam_step(w, st, g, lr=0.1)
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    m, v = 0.1 * g["x"], 0.001 * g["x"] ** 2
    np.testing.assert_allclose(w["x"], np.array([1.0, -2.0]) - lr_t * m / (np.sqrt(v) + 1e-7), rtol=1e-12)
    assert st["iterations"] == 1


def test_dense_adam_moves_untouched_rows_lazy_does_not():
    rng = np.random.default_rng(5)
    p = dict(PARAMS, layers_l2reg=[0, 0, 0])
    w_d = o.init_weights(7, 11, p["layers_sizes"], 0, rng)
    w_l = {k: v.copy() for k, v in w_d.items()}
    s_d, s_l = o.new_opt_state(w_d), o.new_opt_state(w_l)
    u1, i1, y = np.array([0, 0, 0, 0]), np.array([1, 2, 3, 4]), np.array([0, 0, 0, 1])
    u2, i2 = np.array([1, 1, 1, 1]), np.array([5, 6, 7, 8])
    for (u, i) in ((u1, i1), (u2, i2)):
        o.train_step(w_d, s_d, u, i, y, p, adam_mode="dense")
        o.train_step(w_l, s_l, u, i, y, p, adam_mode="lazy")
    # after step 1 both agree; step 2 does not touch user 0, dense mode still moves it
    assert not np.allclose(w_d[o.K_USER][0], w_l[o.K_USER][0])
    np.testing.assert_allclose(w_d[o.K_USER][1], w_l[o.K_USER][1], rtol=1e-6)
    np.testing.assert_allclose(w_d["hidden_1/kernel"], w_l["hidden_1/kernel"], rtol=1e-5, atol=1e-8)


def test_mf_dim_zero_is_the_reference_model():
    rng = np.random.default_rng(9)
    w = o.init_weights(5, 10, [6, 4], 0, rng)
    assert o.weight_names([6, 4]) == list(w.keys()) and len(w) == 6
    c = o.forward(w, [1, 2], [3, 4])
    x = np.concatenate([w[o.K_USER][[1, 2]], w[o.K_ITEM][[3, 4]]], 1)
    h = np.maximum(x @ w["hidden_1/kernel"] + w["hidden_1/bias"], 0)
    np.testing.assert_allclose(c["z"], (h @ w[o.K_OUT_W] + w[o.K_OUT_B]).ravel(), rtol=1e-6)
    assert o.embedding_dims([5, 4]) == (2, 3)  # odd L0: item side gets the extra unit


def test_early_stopping_and_epoch_mean():
    assert o.early_stopping_trace([0.1, 0.3, 0.2, 0.2, 0.2, 0.2, 0.2, 0.9], patience=5) == (7, 1, True)
    assert o.early_stopping_trace([0.1, 0.2, 0.3], patience=5) == (3, 2, False)
    assert o.weighted_epoch_mean([1.0, 0.0], [3, 1]) == pytest.approx(0.75)
