"""Recommender model: the `movierec.model` surface of carlamb/MovieRecommender-TF-TRT
(reference movierec/model.py) on top of the B200 kernels.

Same constructor contract, parameter dict, exceptions, method names and constants as the
reference; the Keras graph (model.py:135-215) is replaced by `_engine.NeuMFEngine`, which owns
the device tensors and calls libmovierec_b200.so.  Optional new parameters keep the reference
behaviour by default: `mf_dim` (0 = the reference's MLP-only model), `adam_mode` ("dense" =
legacy-Keras Adam over every table row, "sparse" = touched rows only), `seed`.
Deviations, each deliberate: weights are saved under the reference's `.h5` file names in the Keras
HDF5 layout when h5py is installed (util/keras_h5.py) and as NumPy `.npz` payloads under the same
names otherwise (this image has no h5py; `load_weights` tells the two apart by the file signature),
and `load_from_files` passes name and directory in the declared order (the reference swaps them,
model.py:301).
"""

import json
import logging
import os
import random

import numpy as np

DEFAULT_PARAMS = {  # the reference's toy defaults (model.py:15-34)
    "num_users": 5,
    "num_items": 10,
    "layers_sizes": [5, 4],
    "layers_l2reg": [0.01, 0.01],
    "optimizer": "adam",
    "lr": 0.001,
    "beta_1": 0.9,
    "beta_2": 0.999,
    "batch_size": 6,
    "num_negs_per_pos": 2,
    "batch_size_eval": 12,
    "num_negs_per_pos_eval": 5,
    "k": 3,
}

ADAM_NAME = "adam"
SGD_NAME = "sgd"
OPTIMIZERS = [ADAM_NAME, SGD_NAME]

HIT_RATE = "hr"
DCG = "dcg"

OUTPUT_PRED = "output"
OUTPUT_RANK = "rank"

METRIC_VAL_DCG = "val_{}_{}".format(OUTPUT_PRED, DCG)

EARLY_STOPPING_PATIENCE = 5  # model.py:324


def _engine_module():
    # imported lazily so that parameter validation (pure host logic) works without a GPU;
    # anything that computes goes through the CUDA library and fails loudly without it
    from . import _engine
    return _engine


class History(object):
    """What Keras' fit_generator returns: `.history` maps metric name -> list of epoch values."""

    def __init__(self):
        self.history = {}
        self.epoch = []

    def _append(self, epoch, logs):
        self.epoch.append(epoch)
        for k, v in logs.items():
            self.history.setdefault(k, []).append(v)


class NeuMFModel(object):
    """Stands where the Keras `Model` stood (`MovierecModel.model`): batch-level entry points with
    Keras' names, backed by the CUDA engine."""

    def __init__(self, owner, engine):
        self._owner = owner
        self.engine = engine
        self.name = owner.name
        self.stop_training = False
        self.output_names = [OUTPUT_PRED, OUTPUT_RANK]
        self.metrics_names = ["loss", OUTPUT_PRED + "_loss", OUTPUT_PRED + "_" + HIT_RATE, OUTPUT_PRED + "_" + DCG]

    # ---- weights ---------------------------------------------------------------------------------
    @property
    def weight_names(self):
        return self.engine.weight_names()

    def get_weights(self):
        w = self.engine.get_weights()
        return [w[k] for k in self.engine.weight_names()]

    def set_weights(self, weights):
        self.engine.set_weights(weights)

    def save_weights(self, path):
        """Keras-layout HDF5 when h5py is installed (the reference's format, model.py:245), else an .npz payload
        under the same file name."""
        from .util import keras_h5
        if keras_h5.have_h5py():
            keras_h5.write(path, self.engine.get_weights(), self.engine.weight_names(),
                           self.engine.get_optimizer_state())
            return
        arrays = dict(self.engine.get_weights())
        arrays.update({"optimizer/" + k: np.asarray(v) for k, v in self.engine.get_optimizer_state().items()})
        with open(path, "wb") as f:  # keep the caller's file name (np.savez would append .npz)
            np.savez(f, **arrays)

    def load_weights(self, path):
        from .util import keras_h5
        if keras_h5.is_hdf5(path):  # written by Keras (the reference) or by save_weights above
            weights, opt = keras_h5.read(path)
            missing = [k for k in self.engine.weight_names() if k not in weights]
            if missing:
                raise ValueError("weight file {} lacks {}".format(path, ", ".join(missing)))
            self.engine.set_weights({k: weights[k] for k in self.engine.weight_names()})
            if opt:
                self.engine.set_optimizer_state(opt)
            return
        with np.load(path) as z:
            self.engine.set_weights({k: z[k] for k in self.engine.weight_names()})
            opt = {k[len("optimizer/"):]: z[k] for k in z.files if k.startswith("optimizer/")}
        if opt:
            self.engine.set_optimizer_state(opt)

    def summary(self, print_fn=print):
        e = self.engine
        print_fn("Model: {}".format(self.name))
        total = 0
        for k in e.weight_names():
            shape = tuple(e._view(k).shape)
            total += int(np.prod(shape))
            print_fn("  {:<36s} {}".format(k, shape))
        print_fn("Total params: {}".format(total))

    # ---- batch entry points ------------------------------------------------------------------------
    def predict_on_batch(self, x):
        """[x_users, x_items] -> [output (B,1) float32, rank (G, negs_eval+1) int32]
        (reference test/test_model.py:161-166; the rank layer is in eval phase, model.py:347)."""
        x_users, x_items = x
        o = self._owner
        group = o._num_negs_per_pos_eval + 1
        _, probs, _ = self.engine.forward(x_users, x_items, want_logits=False)
        if probs.numel() % group:
            raise ValueError("batch of {} rows is not divisible by (num_negs_per_pos_eval + 1) = {}".format(
                probs.numel(), group))
        rank, _, _ = _engine_module().rank_scores(probs, group, o._k, want_rank=True, device=self.engine.device)
        out = probs.cpu().numpy().reshape(-1, 1)
        self._raise_on_bad_ids(bool(np.isnan(out).any()))  # the kernels mark rows with an out-of-range id with NaN
        return [out, rank.cpu().numpy()]

    def recommend(self, user_id, candidate_items, top_k=10):
        """Serving-shaped scoring (reference client, trt_client.py:47-57: one user, N candidate items, the K best):
        returns (items (K,), scores (K,)) ordered as the RankLayer orders them (descending score, the earlier
        candidate first among ties).  One forward of N rows that share the user (the user row is read once per
        launch tile) and one rank pass, both on the device."""
        items = np.asarray(candidate_items, dtype=np.int32).reshape(-1)
        n = int(items.size)
        if n == 0:
            return items, np.zeros(0, np.float32)
        eng = self.engine
        _, probs, _ = eng.forward(np.asarray([user_id], dtype=np.int32), items, user_div=n, want_logits=False)
        rank, _, _ = _engine_module().rank_scores(probs, n, min(int(top_k), n), want_rank=True, device=eng.device)
        order = rank.cpu().numpy().reshape(-1)[:min(int(top_k), n)]
        p = probs.cpu().numpy()
        return items[order], p[order]

    def _step_logs(self, out, rows, group):
        loss = out[0] / rows + out[3]
        return {"loss": loss, OUTPUT_PRED + "_loss": out[0] / rows,
                OUTPUT_PRED + "_" + HIT_RATE: out[1] / (rows // group), OUTPUT_PRED + "_" + DCG: out[2] / (rows // group)}

    def train_on_batch(self, x, y, prefetch=None):
        """One optimisation step; returns [loss, output_loss, output_hr, output_dcg] like Keras.

        prefetch=([x_users, x_items], y) (optional, not in Keras): the NEXT batch.  Its host-to-device copies are
        issued on a copy stream before this step's loss is read back, so they run under this step; the next call
        finds them by the identity of the host arrays and skips its own upload.  Pinned host tensors make the
        copies asynchronous."""
        x_users, x_items = x
        o = self._owner
        group = o._num_negs_per_pos + 1
        # the reference accepts any users per row; the once-per-group fast path is taken only when the batch
        # really has the generator's layout (checked on the device before the step: one tiny kernel)
        eng = self.engine
        mod = _engine_module()
        rows = int(np.asarray(y).size if not hasattr(y, "numel") else y.numel())
        staged = getattr(self, "_prefetched", None)
        if staged is not None and staged[0] == (id(x_users), id(x_items), id(y)):
            _, d_users, d_items, d_y, ready = staged
            import torch
            torch.cuda.current_stream(eng.device).wait_event(ready)
        else:
            d_users = mod.as_device_i32(x_users, eng.device)
            d_items, d_y = x_items, y
        self._prefetched = None
        grouped = eng.users_grouped(d_users, group)
        out_dev = eng.train_step(d_users, d_items, d_y, group=group, k=o._k, grouped=grouped)
        if prefetch is not None:
            import torch
            (n_users, n_items), n_y = prefetch
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=eng.device)
            with torch.cuda.stream(self._copy_stream):
                staged = (mod.as_device_i32(n_users, eng.device), mod.as_device_i32(n_items, eng.device),
                          mod.as_device_f32(n_y, eng.device))
                ready = torch.cuda.Event()
                ready.record(self._copy_stream)
            self._prefetched = ((id(n_users), id(n_items), id(n_y)),) + staged + (ready,)
        out = out_dev.cpu().numpy().astype(np.float64)
        if int(out[4]) & 1:
            raise IndexError("user/item id out of range in batch (num_users={}, num_items={})".format(
                o._num_users, o._num_items))
        logs = self._step_logs(out, rows, group)
        return [logs[n] for n in self.metrics_names]

    def test_on_batch(self, x, y):
        """Loss and ranking metrics of one validation batch (eval phase: groups of negs_eval+1)."""
        sums, rows = self._eval_batch_sums(x, y)
        o = self._owner
        out = sums.cpu().numpy().astype(np.float64)
        self._raise_on_bad_ids(out[3])
        group = o._num_negs_per_pos_eval + 1
        logs = self._step_logs([out[0], out[1], out[2], 0.0], rows, group)
        return [logs[n] for n in self.metrics_names]

    def _eval_batch_sums(self, x, y):
        import torch
        x_users, x_items = x
        o = self._owner
        eng = self.engine
        group = o._num_negs_per_pos_eval + 1
        _, probs, loss = eng.forward(x_users, x_items, labels=y, want_logits=False)
        rows = probs.numel()
        # the label column is the argmax of y_true per group (model.py:447-448), wherever the batch puts its positive
        _, _, sums = _engine_module().rank_scores(probs, group, o._k, want_rank=False, device=eng.device, labels=y)
        # an out-of-range id makes its row's probability NaN (and is left out of the loss): reported with the sums so
        # that callers raise instead of returning metrics of a partly invalid batch
        bad = torch.isnan(probs).any().to(loss.dtype).reshape(1)
        return torch.cat([loss, sums, bad]), rows

    def evaluate_generator(self, generator, steps=None):
        """Keras `evaluate_generator`: batch-size-weighted mean of the per-batch values over
        `len(generator)` batches."""
        import torch
        steps = len(generator) if steps is None else steps
        o = self._owner
        group = o._num_negs_per_pos_eval + 1
        acc = torch.zeros(4, dtype=torch.float64, device=self.engine.device)
        rows = 0
        for b in range(steps):
            x, y = _device_batch(generator, b)
            s, r = self._eval_batch_sums(x, y)
            acc += s.double()
            rows += r
        if rows == 0:
            return [float("nan")] * 4
        out = acc.cpu().numpy()
        self._raise_on_bad_ids(out[3])
        logs = self._step_logs([out[0], out[1], out[2], self._l2_penalty()], rows, group)
        return [logs[n] for n in self.metrics_names]

    def _raise_on_bad_ids(self, flag):
        if flag:
            o = self._owner
            raise IndexError("user/item id out of range in an evaluation batch (num_users={}, num_items={})".format(
                o._num_users, o._num_items))

    def _l2_penalty(self):
        o = self._owner
        if not any(o._layers_l2reg):
            return 0.0
        # the regulariser term of the reported loss (model.py:163,168,178); off the hot path and only
        # for non-default l2: computed from a host copy of the weights
        w = self.engine.get_weights()
        tot = 0.0
        for k, v in w.items():
            if k.endswith("embeddings"):
                tot += o._layers_l2reg[0] * float(np.sum(np.square(v, dtype=np.float64)))
            elif k.startswith("hidden_") and k.endswith("kernel"):
                tot += o._layers_l2reg[int(k.split("/")[0].split("_")[1])] * float(np.sum(np.square(v, dtype=np.float64)))
        return tot

    def fit_generator(self, generator, validation_data=None, epochs=1, callbacks=None, verbose=1, shuffle=True):
        """The Keras training loop the reference relies on (model.py:329-333): `epochs` x
        `len(generator)` train_on_batch steps (batch order shuffled per epoch as Keras does for a
        Sequence), validation after every epoch, callbacks, `generator.on_epoch_end()`.
        Step outputs stay on the device and are read back once per epoch."""
        import torch
        o = self._owner
        eng = self.engine
        history = History()
        callbacks = callbacks or []
        for cb in callbacks:
            cb.set_model(self)
            cb.on_train_begin()
        self.stop_training = False
        group = o._num_negs_per_pos + 1
        for epoch in range(epochs):
            steps = len(generator)
            order = list(range(steps))
            if shuffle:
                random.shuffle(order)
            acc = torch.zeros(5, dtype=torch.float64, device=eng.device)
            rows = 0
            for b in order:
                x, y = _device_batch(generator, b)
                out = eng.train_step(x[0], x[1], y, group=group, k=o._k, grouped=_is_grouped(generator, eng, x[0], group))
                acc[:3] += out[:3].double()
                acc[3] += out[3].double() * y.numel()
                acc[4] += out[4].double()
                rows += int(y.numel())
            logs = {}
            if rows:
                a = acc.cpu().numpy()
                if a[4] != 0:  # sum over steps of the per-step flag word (bit 0 bad id, bit 1 group layout broken)
                    raise IndexError("user/item id out of range (or a generator that declares grouped batches "
                                     "produced a group with more than one user) during training")
                logs = self._step_logs([a[0], a[1], a[2], a[3] / rows], rows, group)
            if validation_data is not None:
                vals = self.evaluate_generator(validation_data)
                for n, v in zip(self.metrics_names, vals):
                    logs["val_" + n] = v
            history._append(epoch, logs)
            if verbose:
                logging.info("Epoch %d/%d - %s", epoch + 1, epochs,
                             " - ".join("{}: {:.4f}".format(k, v) for k, v in logs.items()))
            for cb in callbacks:
                cb.on_epoch_end(epoch, logs)
            if hasattr(generator, "on_epoch_end"):
                generator.on_epoch_end()
            if self.stop_training:
                break
        for cb in callbacks:
            cb.on_train_end()
        return history


def _is_grouped(generator, engine, x_users, group):
    """Whether a training batch has one user per group.  Generators of this package say so themselves
    (`grouped_batches`: MovieLensDataGenerator always builds groups of one user, data_pipeline.py:99-150, and
    the device verifies it during the step); batches of any other Sequence are checked on the device first."""
    if getattr(generator, "grouped_batches", False):
        return True
    return engine.users_grouped(x_users, group)


def _device_batch(generator, b):
    """Batch `b` of a generator as device-ready arrays.  Our generator hands out device tensors
    (no host round trip); any other Sequence gives NumPy arrays that the engine uploads."""
    if hasattr(generator, "device_batch"):
        return generator.device_batch(b)
    import torch
    x, y = generator[b]
    return [x[0], x[1]], torch.as_tensor(np.asarray(y, dtype=np.float32))


class EarlyStopping(object):
    """Keras EarlyStopping as the reference configures it (model.py:324-325): monitor a metric to
    maximise, stop after `patience` epochs without improvement, restore the best weights when
    stopping."""

    def __init__(self, monitor=METRIC_VAL_DCG, mode="max", restore_best_weights=True, patience=5, verbose=0):
        if mode != "max":
            raise NotImplementedError("only mode='max' is used by the reference")
        self.monitor, self.patience, self.restore_best_weights, self.verbose = monitor, patience, restore_best_weights, verbose
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self):
        self.wait, self.stopped_epoch, self.best, self.best_weights = 0, 0, -np.inf, None

    def on_epoch_end(self, epoch, logs):
        current = logs.get(self.monitor)
        if current is None:
            return
        if current > self.best:
            self.best, self.wait = current, 0
            if self.restore_best_weights:
                self.best_weights = self.model.get_weights()
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.stopped_epoch = epoch
                self.model.stop_training = True
                if self.restore_best_weights and self.best_weights is not None:
                    self.model.set_weights(self.best_weights)

    def on_train_end(self):
        if self.stopped_epoch > 0 and self.verbose:
            logging.info("Epoch %05d: early stopping", self.stopped_epoch + 1)


class ModelCheckpoint(object):
    """Keras ModelCheckpoint(save_best_only=True, mode='max') (model.py:326-327); the file name
    pattern is the reference's, the payload is `.npz` weights + optimizer state."""

    def __init__(self, filepath, monitor=METRIC_VAL_DCG, save_best_only=True, mode="max", verbose=0):
        self.filepath, self.monitor, self.save_best_only, self.verbose = filepath, monitor, save_best_only, verbose
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self):
        self.best = -np.inf

    def on_epoch_end(self, epoch, logs):
        current = logs.get(self.monitor)
        if current is None:
            return
        if self.save_best_only and not current > self.best:
            return
        self.best = max(self.best, current)
        path = self.filepath.format(epoch=epoch + 1, **logs)
        self.model.save_weights(path)
        if self.verbose:
            logging.info("Epoch %05d: %s improved to %.5f, saving model to %s", epoch + 1, self.monitor, current, path)

    def on_train_end(self):
        pass


class MovierecModel(object):
    """
    Movie Recommendation Model (reference movierec/model.py:49-333).
    """

    def __init__(self, params=DEFAULT_PARAMS, model_name='movierec', output_dir="models/", verbose=1):
        # Same order of reads and checks as the reference (model.py:73-112) so that a missing key or
        # a bad value raises the same exception type at the same point.
        self._num_users = params["num_users"]
        self._num_items = params["num_items"]
        self._layers_sizes = params["layers_sizes"]
        self._layers_l2reg = params["layers_l2reg"]
        if len(self._layers_sizes) != len(self._layers_l2reg):
            raise ValueError("'layers_sizes' length = {}, 'layers_l2reg' length = {}, but must be equal."
                             .format(len(self._layers_sizes), len(self._layers_l2reg)))
        self._num_layers = len(self._layers_sizes)

        self._optimizer = params["optimizer"]
        if self._optimizer not in OPTIMIZERS:
            raise NotImplementedError("Optimizer {} is not implemented.".format(params["optimizer"]))
        self._lr = params["lr"]
        self._beta_1 = params.get("beta_1", 0.9)
        self._beta_2 = params.get("beta_2", 0.999)
        self._batch_size = params["batch_size"]
        self._num_negs_per_pos = params["num_negs_per_pos"]
        if self._num_negs_per_pos <= 0:
            raise ValueError("num_negs_per_pos must be > 0, found {}".format(self._num_negs_per_pos))
        if self._batch_size % (self._num_negs_per_pos + 1):
            raise ValueError("Batch size must be divisible by (num_negs_per_pos + 1). Found: batch_size={}, "
                             "num_negs_per_pos={}".format(self._batch_size, self._num_negs_per_pos))

        self._batch_size_eval = params["batch_size_eval"]
        self._num_negs_per_pos_eval = params["num_negs_per_pos_eval"]
        if self._num_negs_per_pos_eval <= 0:
            raise ValueError("num_negs_per_pos_eval must be > 0, found {}".format(self._num_negs_per_pos_eval))
        if self._batch_size_eval % (self._num_negs_per_pos_eval + 1):
            raise ValueError("Batch size (eval) must be divisible by (num_negs_per_pos_eval + 1). Found: "
                             "batch_size_eval={}, num_negs_per_pos_eval={}".format(self._batch_size_eval,
                                                                                   self._num_negs_per_pos_eval))

        self._k = params.get("k", self._num_negs_per_pos + 1)
        if self._k > (self._num_negs_per_pos + 1):  # the reference checks the TRAIN negatives only (model.py:108-112)
            raise ValueError("'k' must be lower than (num_negs_per_pos + 1) and lower than (num_negs_per_pos_eval + 1)."
                             "Found: k={}, num_negs_per_pos={}, num_negs_per_pos_eval={}"
                             .format(self._k, self._num_negs_per_pos, self._num_negs_per_pos_eval))

        # optional extensions; defaults reproduce the reference model exactly
        self._mf_dim = params.get("mf_dim", 0)
        self._adam_mode = params.get("adam_mode", "dense")
        self._seed = params.get("seed", None)

        os.makedirs(output_dir, exist_ok=True)
        self.name = model_name
        self._model_weights_path = self.get_model_weights_path(output_dir, model_name)
        self._params_path = self.get_params_json_path(output_dir, model_name)
        self._serialized_params = json.dumps(params)
        self._output_model_checkpoints = os.path.join(
            output_dir, "{}-checkpoint-{{epoch:02d}}-{{val_loss:.2f}}.h5".format(model_name))
        self.verbose = verbose

        self.model = self.build_mlp_model()
        self.compile_model()

    def build_mlp_model(self):
        """
        [ User Embedding ][ Item Embedding ] -> N x [ Hidden Dense ReLU ] -> [ Prediction (sigmoid) ]
        (+ GMF branch when mf_dim > 0) -- reference model.py:135-195, built as device tensors.
        """
        eng = _engine_module().NeuMFEngine(
            self._num_users, self._num_items, self._layers_sizes, self._layers_l2reg, mf_dim=self._mf_dim,
            optimizer=self._optimizer, lr=self._lr, beta_1=self._beta_1, beta_2=self._beta_2,
            table_mode=self._adam_mode, seed=self._seed)
        return NeuMFModel(self, eng)

    def compile_model(self):
        # optimizer, BCE loss and the hr/dcg metrics are part of the fused train step
        # (reference model.py:197-215); nothing to build here beyond a consistency check.
        if self._optimizer not in OPTIMIZERS:
            raise NotImplementedError("Optimizer {} is not implemented.".format(self._optimizer))

    @staticmethod
    def get_model_weights_path(output_dir, model_name):
        return os.path.join(output_dir, "{}_weights.h5".format(model_name))

    @staticmethod
    def get_params_json_path(output_dir, model_name):
        return os.path.join(output_dir, "{}_params.json".format(model_name))

    def get_pred_rank(self):
        """Callable standing in for the rank layer's output tensor (model.py:225-234)."""
        return RankLayer(self._num_negs_per_pos, self._num_negs_per_pos_eval, name=OUTPUT_RANK)

    def log_summary(self):
        self.model.summary(print_fn=logging.info)

    def save(self):
        """Save params and weights to files (model.py:239-249)."""
        self.model.save_weights(self._model_weights_path)
        logging.info('Model weights saved to: {}'.format(self._model_weights_path))
        with open(self._params_path, 'w') as f_out:
            f_out.write(self._serialized_params)
        logging.info('Model params saved to: {}'.format(self._params_path))

    @staticmethod
    def load_from_dir(model_dir, model_name, verbose=1):
        params_path = MovierecModel.get_params_json_path(model_dir, model_name)
        weights_path = MovierecModel.get_model_weights_path(model_dir, model_name)
        return MovierecModel.load_from_files(params_path, weights_path, model_dir, model_name, verbose)

    @staticmethod
    def load_from_files(params_path, weights_path, output_model_dir, output_model_name, verbose=1):
        with open(params_path, 'r') as f_in:
            params = json.load(f_in)
        movierec = MovierecModel(params, output_model_name, output_model_dir, verbose)
        movierec.model.load_weights(weights_path)
        return movierec

    def fit_generator(self, train_data_generator, validation_data_generator, epochs):
        """Training loop with the reference's callbacks (model.py:305-333): early stopping on
        `val_output_dcg` (patience 5, best weights restored) and best-only checkpoints."""
        callbacks = [
            EarlyStopping(monitor=METRIC_VAL_DCG, mode='max', restore_best_weights=True,
                          patience=EARLY_STOPPING_PATIENCE, verbose=self.verbose),
            ModelCheckpoint(self._output_model_checkpoints, monitor=METRIC_VAL_DCG, save_best_only=True,
                            mode='max', verbose=self.verbose),
        ]
        return self.model.fit_generator(generator=train_data_generator,
                                        validation_data=validation_data_generator,
                                        epochs=epochs, callbacks=callbacks, verbose=self.verbose)

    def evaluate(self, users, items, k=None):
        """Full-sweep ranking evaluation (BASELINE config 4): one user id per group and
        (negs_eval+1) candidate items per user with the positive last.  Returns (hr, dcg)."""
        group = self._num_negs_per_pos_eval + 1
        k = self._k if k is None else k
        import torch
        eng = self.model.engine
        # (the fused eval kernels rank a candidate with an out-of-range id last and say nothing: the engine checks the
        # ids on the device -- four reductions next to a sweep of millions of rows -- and the verdict is read back
        # with the sums; host arrays are uploaded in chunks under the sweep, see NeuMFEngine.rank_eval)
        pos, sums, _, _ = eng.rank_eval(users, items, group, k, check_ids=True)
        bad = eng.last_eval_bad.to(sums.dtype)
        s = torch.cat([sums.reshape(-1), bad]).cpu().numpy().astype(np.float64)
        self.model._raise_on_bad_ids(s[2])
        G = max(int(pos.numel()), 1)
        return s[0] / G, s[1] / G


class RankLayer(object):
    """Rank of every item inside its (negs+1)-wide group, descending score, lower index first among
    equal scores (reference model.py:336-358).  `call` accepts array-likes or device tensors and
    returns an int32 NumPy array of shape (groups, negs+1)."""

    learning_phase = 0  # 1 = training phase: group width uses the train negatives (model.py:347)

    def __init__(self, num_negs_per_pos_train, num_negs_per_pos_eval, name, **kwargs):
        self.name = name
        self.num_negs_per_pos_train = num_negs_per_pos_train
        self.num_negs_per_pos_eval = num_negs_per_pos_eval

    def call(self, inputs, training=None, **kwargs):
        training = bool(RankLayer.learning_phase) if training is None else training
        negs = self.num_negs_per_pos_train if training else self.num_negs_per_pos_eval
        rank, _, _ = _engine_module().rank_scores(_flat_scores(inputs), negs + 1, negs + 1, want_rank=True)
        return rank.cpu().numpy()

    __call__ = call

    def get_config(self):
        return {'name': self.name, 'num_negs_per_pos_train': self.num_negs_per_pos_train,
                'num_negs_per_pos_eval': self.num_negs_per_pos_eval}


def set_learning_phase(value):
    """Stand-in for K.set_learning_phase used by the reference's rank-layer test."""
    RankLayer.learning_phase = int(value)


def _flat_scores(x):
    if hasattr(x, "detach"):
        return x.detach().reshape(-1)
    return np.asarray(x, dtype=np.float32).reshape(-1)


def _positions_from_rank(y_true, pred_rank_idx, k):
    """(hit_sum, dcg_sum, groups) for a given rank permutation, on device: item j of a group gets the
    pseudo-score -(its place in the rank), so the kernel's position of the label column is exactly
    `where(rank == argmax(y_true))` of model.py:447-451."""
    import torch
    eng = _engine_module()
    eng.require_cuda()
    dev = torch.device("cuda:{}".format(torch.cuda.current_device()))
    rank = torch.as_tensor(np.asarray(pred_rank_idx), device=dev).long()
    G, group = rank.shape
    y = torch.as_tensor(np.asarray(y_true), device=dev).reshape(G, group)
    label_col = torch.argmax(y, dim=1).to(torch.int32)
    place = torch.arange(group, device=dev, dtype=torch.float32).expand(G, group)
    scores = torch.empty((G, group), dtype=torch.float32, device=dev)
    scores.scatter_(1, rank, -place)
    _, _, sums = eng.rank_scores(scores, group, k, label_col=label_col, want_rank=False, device=dev)
    s = sums.cpu().numpy().astype(np.float64)
    return s[0], s[1], G


def hit_rate(y_true, _, k, pred_rank_idx):
    """HR@k of a batch (reference model.py:361-385)."""
    hits, _, G = _positions_from_rank(y_true, pred_rank_idx, k)
    return np.float32(hits / G)


def discounted_cumulative_gain(y_true, _, k, pred_rank_idx):
    """DCG@k of a batch, ln2/ln(pos+2) on hits (reference model.py:388-417)."""
    _, dcg, G = _positions_from_rank(y_true, pred_rank_idx, k)
    return np.float32(dcg / G)
