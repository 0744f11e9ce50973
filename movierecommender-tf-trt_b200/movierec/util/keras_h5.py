"""Keras-layout HDF5 weight files: what `keras.Model.save_weights` / `load_weights` of the
reference write and read (movierec/model.py:245, :302).

Layout (Keras 2.x / tf.keras 1.x `save_weights_to_hdf5_group`): root attributes `layer_names` (bytes), `backend`,
`keras_version`; one group per layer with attribute `weight_names` (bytes, e.g. b"hidden_1/kernel:0") and one
dataset per weight at <layer group>/<weight name>.  Weight names here are the engine's keys
("user_embedding/embeddings", "hidden_1/kernel", "output/bias", ...) plus the ":0" suffix, so a file written by the
reference loads into this package and the other way round; layers without weights (inputs, Flatten, Concatenate,
the rank layer) may be listed with an empty `weight_names` and are skipped.

h5py is optional: without it the package keeps its `.npz` payload and `read` raises ImportError with
that explanation.  Untested against a real TensorFlow-written file in this environment (no TensorFlow, no h5py).
"""

import numpy as np

HDF5_MAGIC = b"\x89HDF\r\n\x1a\n"
OPTIMIZER_GROUP = "movierec_b200_optimizer"  # not a Keras group: Keras walks `layer_names` only


def have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except ImportError:
        return False


def is_hdf5(path):
    with open(path, "rb") as f:
        return f.read(8) == HDF5_MAGIC


def layer_of(weight_name):
    return weight_name.split("/")[0]


def write(path, weights, weight_order, optimizer_state=None):
    """weights: {engine key: array}; weight_order: keys in Keras creation order (layer order of the file)."""
    import h5py
    layers = []
    for k in weight_order:
        if layer_of(k) not in layers:
            layers.append(layer_of(k))
    with h5py.File(path, "w") as f:
        f.attrs["layer_names"] = np.array([n.encode("utf8") for n in layers])
        f.attrs["backend"] = b"movierec_b200"
        f.attrs["keras_version"] = b"2.2.4-tf"
        for layer in layers:
            g = f.create_group(layer)
            names = [k for k in weight_order if layer_of(k) == layer]
            g.attrs["weight_names"] = np.array([(k + ":0").encode("utf8") for k in names])
            for k in names:
                g.create_dataset(k + ":0", data=np.asarray(weights[k]))
        if optimizer_state:
            g = f.create_group(OPTIMIZER_GROUP)
            for k, v in optimizer_state.items():
                g.create_dataset(k, data=np.asarray(v))


def read(path):
    """-> ({engine key: array}, {optimizer key: array})."""
    try:
        import h5py
    except ImportError:
        raise ImportError("{} is an HDF5 (Keras) weight file and h5py is not installed; this package writes "
                          ".npz payloads when h5py is missing".format(path))
    weights, opt = {}, {}
    with h5py.File(path, "r") as f:
        # `save_weights` files carry layer_names at the root; full-model files (ModelCheckpoint with
        # save_weights_only=False, the reference's callback at model.py:317-320) keep them under 'model_weights'
        root = f if "layer_names" in f.attrs else (f["model_weights"] if "model_weights" in f else None)
        if root is None:
            raise ValueError("{} is not a Keras weight file: no layer_names attribute at the root or under "
                             "'model_weights'".format(path))
        for layer in root.attrs["layer_names"]:
            layer = layer.decode("utf8") if isinstance(layer, bytes) else str(layer)
            g = root[layer]
            for wn in g.attrs.get("weight_names", []):
                wn = wn.decode("utf8") if isinstance(wn, bytes) else str(wn)
                key = wn[:-2] if wn.endswith(":0") else wn
                weights[key] = np.asarray(g[wn])
        if OPTIMIZER_GROUP in f:
            def visit(name, obj):
                if isinstance(obj, h5py.Dataset):
                    opt[name] = np.asarray(obj)
            f[OPTIMIZER_GROUP].visititems(visit)
    return weights, opt
