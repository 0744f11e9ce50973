"""The drop-in Python face on a GPU: the reference's own unit tests (test/test_model.py,
test/test_data_pipeline.py) restated against the new package, plus an end-to-end training run."""

import copy
import json
import math
import os

import numpy as np
import pandas as pd
import pytest

from oracle import movierec_oracle as o

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TEST_PARAMS = {  # reference test/test_model.py:8-26
    "num_users": 5, "num_items": 10, "layers_sizes": [6, 4], "layers_l2reg": [0.01, 0.01],
    "optimizer": "adam", "lr": 0.001, "beta_1": 0.9, "beta_2": 0.999,
    "batch_size": 8, "num_negs_per_pos": 3, "batch_size_eval": 10, "num_negs_per_pos_eval": 4, "k": 4,
}


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from movierec import data_pipeline, model, trainer
    return model, data_pipeline, trainer


@pytest.fixture(scope="module")
def ref(golden_dir):
    with open(os.path.join(golden_dir, "reference_tests.json")) as f:
        return json.load(f)


# ---- test/test_model.py ---------------------------------------------------------------------------

def test_build_mlp_model(mods, tmp_path):
    model = mods[0]
    m = model.MovierecModel(copy.deepcopy(TEST_PARAMS), output_dir=str(tmp_path))
    w = m.model.get_weights()
    assert len(w) == 6  # 2 tables, hidden kernel + bias, output kernel + bias (test_model.py:52)
    assert [x.shape for x in w] == [(5, 3), (10, 3), (6, 4), (4,), (4, 1), (1,)]
    assert m.model.weight_names == ["user_embedding/embeddings", "item_embedding/embeddings", "hidden_1/kernel",
                                    "hidden_1/bias", "output/kernel", "output/bias"]
    lim = math.sqrt(6.0 / (10 + 3))
    assert np.all(np.abs(w[1]) <= lim) and np.abs(w[1]).max() > 0.5 * lim  # glorot-uniform range
    assert np.all(w[3] == 0) and np.all(w[5] == 0)


def test_outputs(mods, ref, tmp_path):
    model = mods[0]
    v = ref["outputs_shape"]
    m = model.MovierecModel(copy.deepcopy(v["params"]), output_dir=str(tmp_path))
    m.log_summary()
    output, rank = m.model.predict_on_batch([np.array(v["x_users"]), np.array(v["x_items"])])
    assert output.shape == tuple(v["output_shape"]) and rank.shape == tuple(v["rank_shape"])
    assert output.dtype == np.float32 and rank.dtype == np.int32
    w = dict(zip(m.model.weight_names, m.model.get_weights()))
    c = o.forward(w, v["x_users"], v["x_items"])
    np.testing.assert_allclose(output.ravel(), c["p"], rtol=1e-5)
    np.testing.assert_array_equal(rank, o.rank_groups(output.ravel(), 5))


def test_hit_rate_and_dcg(mods, ref):
    model = mods[0]
    v = ref["hit_rate"]
    for k, want in v["k_to_hr"].items():
        got = model.hit_rate(np.array(v["y_true"]), None, k=int(k), pred_rank_idx=np.array(v["rank"], np.int32))
        assert got == pytest.approx(want)
    v = ref["dcg"]
    rank = model.RankLayer(3, 3, "rank").call(np.array(v["y_pred"], np.float32))
    for k, hits in v["k_to_hits"].items():
        want = sum(h * math.log(2) / math.log(p + 2) for h, p in zip(hits, v["positions"])) / 2.0
        got = model.discounted_cumulative_gain(np.array(v["y_true"]), np.array(v["y_pred"]), int(k), rank)
        assert got == pytest.approx(want, abs=1e-6)


def test_dcg_hr_on_ties(mods, ref):
    model = mods[0]
    v = ref["ties"]
    y_true, y_pred = np.array(v["y_true"]), np.array(v["y_pred"], np.float32)
    rank = model.RankLayer(3, 3, "rank").call(y_pred)
    for k in v["zero_for_k"]:
        assert model.hit_rate(y_true, y_pred, k, rank) == 0.0
        assert model.discounted_cumulative_gain(y_true, y_pred, k, rank) == 0.0
    assert model.hit_rate(y_true, y_pred, v["hit_k"], rank) == pytest.approx(1.0)
    assert model.discounted_cumulative_gain(y_true, y_pred, v["hit_k"], rank) == pytest.approx(
        math.log(2) / math.log(v["hit_position"] + 2), abs=1e-6)


def test_rank_layer(mods, ref):
    model = mods[0]
    layer = model.RankLayer(num_negs_per_pos_train=2, num_negs_per_pos_eval=3, name="rank")
    model.set_learning_phase(1)
    try:
        np.testing.assert_equal(layer.call(np.array(ref["rank_layer"]["train"]["input"])),
                                np.array(ref["rank_layer"]["train"]["expected"]))
    finally:
        model.set_learning_phase(0)
    np.testing.assert_equal(layer.call(np.array(ref["rank_layer"]["eval"]["input"])),
                            np.array(ref["rank_layer"]["eval"]["expected"]))


# ---- test/test_data_pipeline.py -------------------------------------------------------------------

def test_generator_get_item(mods, ref, monkeypatch):
    dp = mods[1]
    v = ref["generator_get_item"]
    monkeypatch.setattr(dp.MovieLensDataGenerator, "num_items", property(lambda self: v["num_items"]))
    gen = dp.MovieLensDataGenerator("ml-100k", pd.DataFrame(v["data"]), batch_size=v["batch_size"],
                                    negatives_per_positive=v["negs"], extra_data_df=pd.DataFrame(v["extra"]), shuffle=False)
    for _ in range(10):
        (xu, xi), y = gen[0]
        np.testing.assert_equal(xu, np.array(v["batch0"]["users"]))
        np.testing.assert_equal(xi, np.array(v["batch0"]["items"]))
        np.testing.assert_equal(y, np.array(v["batch0"]["y"]))
    for _ in range(10):
        (xu, xi), y = gen[1]
        np.testing.assert_equal(xu, np.array(v["batch1"]["users"]))
        np.testing.assert_equal(xi[:3], np.array(v["batch1"]["items_first3"]))
        assert len(np.setdiff1d(np.array(v["batch1"]["user1_candidates"]), xi[3:5])) == 1
        assert xi[5] == v["batch1"]["last_item"]
        np.testing.assert_equal(y, np.array(v["batch1"]["y"]))


def test_generator_get_item_duplicated_user_batch(mods, ref, monkeypatch):
    dp = mods[1]
    v = ref["generator_duplicated_user"]
    monkeypatch.setattr(dp.MovieLensDataGenerator, "num_items", property(lambda self: v["num_items"]))
    gen = dp.MovieLensDataGenerator("ml-100k", pd.DataFrame(v["data"]), batch_size=v["batch_size"],
                                    negatives_per_positive=v["negs"], extra_data_df=None, shuffle=False)
    differ_in_batch = differ_between_runs = False
    last = None
    for _ in range(50):
        (xu, xi), y = gen[0]
        np.testing.assert_equal(xu, np.array(v["users"]))
        assert len(np.setdiff1d(np.array(v["candidates"]), xi[:2])) == 1
        assert len(np.setdiff1d(np.array(v["candidates"]), xi[3:5])) == 1
        np.testing.assert_equal(y, np.array(v["y"]))
        differ_in_batch |= not np.array_equal(xi[:2], xi[3:5])
        differ_between_runs |= last is not None and not np.array_equal(last, xi[:2])
        last = xi[:2]
    assert differ_between_runs and differ_in_batch


def test_generator_matches_device_sampler_oracle(mods):
    dp = mods[1]
    rng = np.random.default_rng(3)
    users = np.repeat(np.arange(30), 8)
    items = np.concatenate([rng.choice(100, 8, replace=False) for _ in range(30)])
    df = pd.DataFrame({"userId": users.astype(np.int32), "itemId": items.astype(np.int32)})
    gen = dp.MovieLensDataGenerator("ml-100k", df, batch_size=50, negatives_per_positive=4, shuffle=False, seed=77)
    rowptr, csr = o.build_csr(30, users, items)
    (xu, xi), y = gen[2]
    want = o.device_sample_batch(rowptr, csr, 1682, users[20:30], items[20:30], 20, 4, 77, 1)
    np.testing.assert_array_equal(xi, want)
    np.testing.assert_array_equal(xu, np.repeat(users[20:30], 5))


# ---- end to end -----------------------------------------------------------------------------------

def synthetic_ml100k(tmp_path, n_users=60, n_items=120, per_user=24, seed=0):
    rng = np.random.default_rng(seed)
    rows = []
    for u in range(n_users):
        liked = rng.choice(n_items // 2, per_user, replace=False) + (n_items // 2) * (u % 2)  # two taste clusters
        rows.extend("{}\t{}\t{}\t{}".format(u + 1, i + 1, 5, 0) for i in liked)
    d = tmp_path / "ml-100k"
    d.mkdir()
    (d / "u.data").write_text("\n".join(rows) + "\n")
    return str(tmp_path)


def test_trainer_end_to_end_learns_and_round_trips(mods, tmp_path):
    model, dp, trainer = mods
    data_dir = synthetic_ml100k(tmp_path)
    params = dict(trainer.DEFAULT_PARAMS)
    params.update(layers_sizes=[32, 16, 8], layers_l2reg=[0, 0, 0], batch_size=60, num_negs_per_pos=5, k=5,
                  batch_size_eval=200, num_negs_per_pos_eval=99, epochs=12, lr=0.01, seed=3, mf_dim=4)
    np.random.seed(0)
    out_dir = str(tmp_path / "models")
    m = trainer.train("toy", "ml-100k", data_dir, out_dir, params=params, verbose=0)
    assert params["num_users"] == 943 and params["num_items"] == 1682  # taken from the generator (trainer.py:72-73)
    assert os.path.exists(os.path.join(out_dir, "toy_weights.h5")) and os.path.exists(os.path.join(out_dir, "toy_params.json"))
    loaded = model.MovierecModel.load_from_dir(out_dir, "toy", verbose=0)
    for a, b in zip(m.model.get_weights(), loaded.model.get_weights()):
        np.testing.assert_array_equal(a, b)
    x = [np.arange(100) % 60, np.arange(100) % 120]
    np.testing.assert_array_equal(m.model.predict_on_batch(x)[0], loaded.model.predict_on_batch(x)[0])


def test_fit_generator_history_and_learning(mods, tmp_path):
    model, dp, trainer = mods
    data_dir = synthetic_ml100k(tmp_path, n_users=240, seed=1)
    train_df, val_df, _ = dp.load_ratings_train_test_sets("ml-100k", data_dir, download=False)
    params = dict(trainer.DEFAULT_PARAMS)
    params.update(layers_sizes=[32, 16, 8], layers_l2reg=[0, 0, 0], batch_size=120, num_negs_per_pos=5, k=5,
                  batch_size_eval=40, num_negs_per_pos_eval=19, lr=0.01, seed=5, num_users=943, num_items=1682,
                  mf_dim=8)
    np.random.seed(1)
    gen, val = trainer.build_generators("ml-100k", train_df, val_df, params)
    # the reference's epoch covers floor(N / batch_size) batches, i.e. 1/(negs+1) of the positives, and
    # validation scores floor(N_users / batch_size_eval) batches of batch_size_eval/(negs+1) users (App. B-1)
    assert len(gen) == len(train_df) // 120 and len(val) == len(val_df) // 40 == 6
    m = model.MovierecModel(params, "fit", str(tmp_path / "fit"), verbose=0)
    h = m.fit_generator(gen, val, epochs=8)
    keys = {"loss", "output_loss", "output_hr", "output_dcg", "val_loss", "val_output_loss", "val_output_hr", "val_output_dcg"}
    assert keys <= set(h.history)
    assert h.history["loss"][-1] < h.history["loss"][0]
    assert all(0.0 <= v <= 1.0 for v in h.history["val_output_hr"])  # 12 users only: the reference's quirk
    ckpts = [f for f in os.listdir(str(tmp_path / "fit")) if "checkpoint" in f]
    assert ckpts and all(f.startswith("fit-checkpoint-") and f.endswith(".h5") for f in ckpts)
    m.model.fit_generator(gen, None, epochs=80, verbose=0)  # keep training without the 12-user early stopping
    # full-sweep evaluate (BASELINE config 4 entry point) agrees with the oracle on the trained weights
    w = dict(zip(m.model.weight_names, m.model.get_weights()))
    rng = np.random.default_rng(2)
    m100 = model.MovierecModel(dict(params, num_negs_per_pos_eval=99, batch_size_eval=200), "fit", str(tmp_path / "fit"), verbose=0)
    m100.model.set_weights(m.model.get_weights())
    users = val_df.userId.values
    other = ((users % 2) ^ 1) * 60  # 99 negatives from the cluster the user never rated, held-out positive last
    items = np.concatenate([np.append(rng.choice(60, 99, replace=True) + off, it) for off, it in zip(other, val_df.itemId.values)])
    hr, dcg = m100.evaluate(users, items, k=10)
    hr_o, dcg_o, _, _ = o.evaluate_groups(w, users, items, 100, 10)
    assert abs(hr - hr_o) <= 1e-3 and abs(dcg - dcg_o) <= 1e-3
    assert hr > 0.5  # chance level is 0.1: the model has learned the two taste clusters


# ---- serving-shaped scoring (reference client: one user x N candidates -> top K, trt_client.py:47-57) -------------

@pytest.mark.parametrize("layers,mf_dim,n", [([64, 32, 16, 8], 8, 1000), ([256, 128, 64], 64, 1000), ([6, 4], 0, 7)])
def test_recommend_top_k_matches_oracle(mods, tmp_path, layers, mf_dim, n):
    model = mods[0]
    params = copy.deepcopy(TEST_PARAMS)
    params.update(num_users=50, num_items=1200, layers_sizes=layers, layers_l2reg=[0.0] * len(layers), mf_dim=mf_dim, seed=3)
    m = model.MovierecModel(params, output_dir=str(tmp_path))
    w = dict(zip(m.model.weight_names, m.model.get_weights()))
    rng = np.random.default_rng(5)
    cand = rng.permutation(1200)[:n].astype(np.int32)
    items, scores = m.model.recommend(17, cand, top_k=10)
    k = min(10, n)
    assert items.shape == (k,) and scores.shape == (k,)
    p = o.forward(w, np.full(n, 17), cand)["p"].reshape(-1)
    order = np.argsort(-p, kind="stable")[:k]
    np.testing.assert_allclose(scores, p[order], rtol=1e-5, atol=1e-7)
    # the same items wherever the oracle's scores are separated by more than the tolerance
    sep = np.abs(np.diff(np.sort(p)[::-1][:k + 1])) > 1e-6
    if np.all(sep):
        assert np.array_equal(items, cand[order])
    assert np.all(np.diff(scores) <= 0)


def test_evaluation_paths_raise_on_out_of_range_ids(mods, tmp_path):
    """An id beyond the tables is an IndexError in every evaluation path (the kernels mark such rows with NaN and
    rank them last; without the check the metrics of a partly invalid batch would be returned): predict_on_batch,
    test_on_batch, evaluate_generator's batches, and the full-sweep evaluate()."""
    model_mod, _, _ = mods
    m = model_mod.MovierecModel(copy.deepcopy(TEST_PARAMS), "bad_ids", str(tmp_path), verbose=0)
    group = TEST_PARAMS["num_negs_per_pos_eval"] + 1
    users = np.repeat(np.array([0, 4], dtype=np.int32), group)
    items = np.arange(2 * group, dtype=np.int32) % TEST_PARAMS["num_items"]
    y = np.tile([0.0] * (group - 1) + [1.0], 2).astype(np.float32)
    m.model.predict_on_batch([users, items])           # in range: fine
    m.model.test_on_batch([users, items], y)
    m.evaluate(users[::group], items)
    bad_items = items.copy()
    bad_items[3] = TEST_PARAMS["num_items"]
    bad_users = users.copy()
    bad_users[group:] = TEST_PARAMS["num_users"] + 2
    for u, i in ((users, bad_items), (bad_users, items)):
        with pytest.raises(IndexError):
            m.model.predict_on_batch([u, i])
        with pytest.raises(IndexError):
            m.model.test_on_batch([u, i], y)
        with pytest.raises(IndexError):
            m.evaluate(u[::group], i)
