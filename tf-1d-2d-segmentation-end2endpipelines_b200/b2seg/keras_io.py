"""Keras weight files -> the `{layer/weight: array}` dictionary Model.set_weight_dict takes (SURVEY §8(f) rank 1;
the reference calls model.load_weights / tf.keras.saving.load_model on them: 2DCNN/Train.py:363,375, Test.py:114).

Reading HDF5 needs `h5py`, which is NOT in the build image (and cannot be installed there): it is imported lazily, the product works
without it through `.npz` files, and `tools/keras_weights_to_npz.py` converts on any machine that has it (the reference's own
environment does: its notebooks import h5py).  The layout logic below works on anything that quacks like an h5py group (mapping
with `.attrs`), which is how it is tested here without HDF5 (tests/test_callbacks_cpu.py).  †: the layouts are the documented
Keras ones; no file written by a real Keras has been read by this code yet.

Layouts understood:
  * Keras-2 `save_weights('x.h5')`: root attribute `layer_names`; one group per layer with attribute `weight_names`
    (e.g. b'conv2d/kernel:0') naming datasets below it.  `model.save('x.h5')` nests the same under `model_weights`.
  * `.keras` archives (TF >= 2.13 `model.save('x.keras')`): a zip whose `model.weights.h5` holds `layers/<layer>/vars/<i>` (and
    `_layer_checkpoint_dependencies/<layer>/vars/<i>` in some versions) with variables in creation order and no names; they are
    matched to the product's per-layer weight order (kernel, bias / gamma, beta, moving_mean, moving_variance /
    kernel, recurrent_kernel, bias).
"""
from __future__ import annotations

import io
import zipfile
from typing import Dict, List, Tuple

import numpy as np


def _s(v) -> str:
    return v.decode("utf-8") if isinstance(v, (bytes, np.bytes_)) else str(v)


def _legacy_tree(root) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    for lname in [_s(n) for n in root.attrs["layer_names"]]:
        grp = root[lname]
        for wname in [_s(n) for n in grp.attrs.get("weight_names", [])]:
            key = wname.split(":")[0]                       # 'conv2d/kernel:0' -> 'conv2d/kernel'
            if not key.startswith(lname + "/"):             # nested models list their variables without the outer name
                key = f"{lname}/{key}"
            out[key] = np.asarray(grp[wname])
    return out


def _v3_tree(layers, order: Dict[str, List[str]]) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    for lname in layers.keys():
        if "vars" not in layers[lname].keys():
            continue
        vs = layers[lname]["vars"]
        idx = sorted(vs.keys(), key=lambda k: int(k))
        if not idx:
            continue
        names = order.get(lname)
        if names is None or len(names) != len(idx):
            raise ValueError(f"layer '{lname}': the file holds {len(idx)} variables, the model expects {names}")
        for i, wn in zip(idx, names):
            out[f"{lname}/{wn}"] = np.asarray(vs[i])
    return out


def weights_from_tree(root, order: Dict[str, List[str]]) -> Dict[str, np.ndarray]:
    """root: an h5py.File-like mapping; order: {layer name: [weight names in creation order]} of the receiving model"""
    if "layer_names" in root.attrs:
        return _legacy_tree(root)
    if "model_weights" in root.keys() and "layer_names" in root["model_weights"].attrs:
        return _legacy_tree(root["model_weights"])
    for top in ("layers", "_layer_checkpoint_dependencies"):
        if top in root.keys():
            return _v3_tree(root[top], order)
    raise ValueError("not a Keras weight file: no `layer_names` attribute and no `layers` group")


def weight_order(param_specs) -> Dict[str, List[str]]:
    order: Dict[str, List[str]] = {}
    for (layer, wname, _shape, _init, _tr) in param_specs:
        order.setdefault(layer, []).append(wname)
    return order


def read_keras_weights(path: str, param_specs) -> Dict[str, np.ndarray]:
    try:
        import h5py
    except ImportError as e:
        raise NotImplementedError(
            f"reading '{path}' needs the h5py package, which this environment does not have; convert the file where the reference runs "
            f"with `python tools/keras_weights_to_npz.py {path} out.npz` and load the .npz") from e
    order = weight_order(param_specs)
    if zipfile.is_zipfile(path):
        with zipfile.ZipFile(path) as z:
            with h5py.File(io.BytesIO(z.read("model.weights.h5")), "r") as f:
                return weights_from_tree(f, order)
    with h5py.File(path, "r") as f:
        return weights_from_tree(f, order)


def select_for_model(found: Dict[str, np.ndarray], param_specs) -> Tuple[Dict[str, np.ndarray], List[str]]:
    """(weights the model has, keys of the file it does not) — Keras' by_name=False loading is strict about the former"""
    want = {f"{l}/{w}": tuple(s) for (l, w, s, _i, _t) in param_specs}
    missing = [k for k in want if k not in found]
    if missing:
        raise ValueError(f"the weight file lacks {len(missing)} of the model's {len(want)} weights, e.g. {missing[:4]}")
    for k, shp in want.items():
        if tuple(found[k].shape) != shp:
            raise ValueError(f"weight {k}: file has shape {tuple(found[k].shape)}, the model {shp}")
    return {k: found[k].astype(np.float32) for k in want}, [k for k in found if k not in want]
