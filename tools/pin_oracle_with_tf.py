"""Pin the oracle against the REAL reference, wherever TensorFlow can run (plan item 3).

The build image has no TensorFlow, so the oracle (oracle/keras_ref.py) is an unpinned reading of Keras-2 semantics (SURVEY §9, items
marked †).  This script closes that gap on any machine with TensorFlow 2.13-2.15 (or `tf_keras`) and a checkout of the reference:

    python tools/pin_oracle_with_tf.py --reference /path/to/TF-1D-2D-Segmentation-End2EndPipelines [--case unet2d ...]

For each case it builds the reference model with the reference's own builder, checks that the layer names / weight shapes equal
the product graph's (the Keras auto-name replay), copies one set of weights into both, and compares: every layer output of a
training-mode forward pass, the loss, every parameter gradient (tf.GradientTape), the weights after one Adam step, and the
BatchNorm moving statistics — reference (float32) against oracle (float64).  Exit code 0 = everything within 1e-4 relative.
NOT YET RUN: written without a TensorFlow to run it against; tests/test_oracle_vs_tf.py calls it and skips where TF is missing.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tf-1d-2d-segmentation-end2endpipelines_b200"))

CASES = {
    # name: (ndim, reference builder call as (module path, factory), product builder, oracle, input shape, loss)
    "unet2d": dict(ndim=2, variant="UNet", args=(32, 32, 8, 2), kw=dict(num_channels=3), loss="bce"),
    "unetpp2d_ds_ag": dict(ndim=2, variant="UNetPP", args=(32, 32, 8, 2), kw=dict(num_channels=2, ds=1, ag=1, output_nums=3, final_activation="softmax"), loss="cce"),
    "unet2d_lstm": dict(ndim=2, variant="UNet", args=(32, 32, 16, 2), kw=dict(num_channels=1, lstm=1, dense_loop=2), loss="bce"),
    "unet2d_bilinear": dict(ndim=2, variant="UNet", args=(32, 32, 8, 2), kw=dict(num_channels=1, is_transconv=False), loss="bce"),
    "multires2d": dict(ndim=2, variant="MultiResUNet", args=(32, 32, 16, 2), kw=dict(num_channels=1), loss="bce"),
    "unet1d": dict(ndim=1, variant="UNet", args=(64, 2, 1, 8, 3), kw=dict(problem_type="Regression", output_nums=1, ds=1, ag=1), loss="mse"),
}


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def run_case(name, spec, reference_root, tol=1e-4, verbose=True):
    import tensorflow as tf
    try:                                     # Keras 3 installs need the Keras-2 shim the reference was written for
        import tf_keras as keras
    except ImportError:
        keras = tf.keras
    import torch
    from b2seg.graph import init_params
    from b2seg.models1d import UNet as ProductUNet
    from b2seg.models2d import unet_model_builder as product_builder
    from oracle.keras_ref import KerasRef, keras_adam_step, keras_loss
    from oracle.ref_models import Ref1D, Ref2D
    ndim, variant, args, kw = spec["ndim"], spec["variant"], spec["args"], spec["kw"]
    sub = "2DCNN" if ndim == 2 else "1DCNN"
    sys.path.insert(0, os.path.join(reference_root, "TensorFlow", sub))
    keras.backend.clear_session()            # fresh auto-name counters, as the product assumes
    if ndim == 2:
        from models.unet_variants import unet_model_builder as ref_builder
        ref_model = ref_builder(variant, *args, train_mode="from_scratch", **kw).ResNet50()
        graph = product_builder(variant, *args, train_mode="from_scratch", **kw).build_graph()
        oracle = Ref2D(variant, *args, **kw)
        x = np.random.default_rng(1).random((2, args[0], args[1], kw.get("num_channels", 3)), dtype=np.float32)
    else:
        from Models.unet_variants import UNet as RefUNet
        ref_model = getattr(RefUNet(*args, **kw), variant)()
        graph = getattr(ProductUNet(*args, **kw), variant)().graph
        oracle = Ref1D(variant, *args, **kw)
        x = np.random.default_rng(1).standard_normal((2, args[0], args[2])).astype(np.float32)
    report = {"case": name, "failures": []}

    def check(what, err):
        if verbose:
            print(f"  {what:<58s} {err:.2e}")
        if not err < tol:
            report["failures"].append((what, err))

    # ---- 1. names and shapes: the Keras auto-name replay
    theirs = {}
    for layer in ref_model.layers:
        for v in layer.weights:
            theirs[f"{layer.name}/{v.name.split('/')[-1].split(':')[0]}"] = tuple(v.shape)
    ours = {f"{l}/{w}": tuple(s) for (l, w, s, _i, _t) in graph.param_specs()}
    if theirs != ours:
        only_t, only_o = sorted(set(theirs) - set(ours))[:5], sorted(set(ours) - set(theirs))[:5]
        shape = [(k, theirs[k], ours[k]) for k in theirs if k in ours and theirs[k] != ours[k]][:5]
        report["failures"].append(("weight names / shapes", f"reference only {only_t}, product only {only_o}, shapes {shape}"))
        return report
    # ---- 2. one set of weights in both (BN affine and biases perturbed so that they matter)
    rng = np.random.default_rng(2)
    params = init_params(graph, seed=7)
    for k in params:
        if k.endswith("/gamma"):
            params[k] = (1 + 0.2 * rng.standard_normal(params[k].shape)).astype(np.float32)
        elif k.endswith(("/beta", "/bias")):
            params[k] = (0.1 * rng.standard_normal(params[k].shape)).astype(np.float32)
    for layer in ref_model.layers:
        if layer.weights:
            layer.set_weights([params[f"{layer.name}/{v.name.split('/')[-1].split(':')[0]}"] for v in layer.weights])
    # ---- 3. training-mode forward: every layer output
    taps = [l for l in ref_model.layers if not isinstance(l, keras.layers.InputLayer)]
    tap_model = keras.Model(ref_model.inputs, [l.output for l in taps])
    outs_tf = tap_model(x, training=True)
    tp = {k: torch.from_numpy(v.copy()).double() for k, v in params.items()}
    k = KerasRef(ndim, params=tp, dtype=torch.float64, training=True, strict=True)
    outs = oracle(k, torch.from_numpy(x).double())
    for layer, y in zip(taps, outs_tf):
        if layer.name in k.acts and tuple(k.acts[layer.name].shape) == tuple(y.shape):
            check(f"forward {layer.name}", _rel(k.acts[layer.name].detach().numpy(), y.numpy()))
    # ---- 4. loss, gradients, one Adam step, moving statistics
    names = [o.name.split("/")[0] for o in ref_model.outputs]
    rngt = np.random.default_rng(3)
    targets = []
    for i, o in enumerate(outs):
        shp = tuple(o.shape)
        if i == 0 and spec["loss"] == "bce":
            targets.append((rngt.random(shp) > 0.6).astype(np.float32))
        elif i == 0 and spec["loss"] == "cce":
            targets.append(np.eye(shp[-1], dtype=np.float32)[rngt.integers(0, shp[-1], shp[:-1])])
        else:
            targets.append(rngt.standard_normal(shp).astype(np.float32))
    kinds = [spec["loss"]] + ["mse"] * (len(outs) - 1)
    tf_losses = {"bce": keras.losses.BinaryCrossentropy(), "cce": keras.losses.CategoricalCrossentropy(), "mse": keras.losses.MeanSquaredError()}
    opt = keras.optimizers.Adam(1e-3)
    with tf.GradientTape() as tape:
        ys = ref_model(x, training=True)
        ys = ys if isinstance(ys, (list, tuple)) else [ys]
        loss_tf = tf.add_n([tf_losses[kd](t, y) for kd, t, y in zip(kinds, targets, ys)])
    grads_tf = tape.gradient(loss_tf, ref_model.trainable_variables)
    total = sum(keras_loss(kd, o, torch.from_numpy(t).double(), logits=k.logits.get(nm)) for kd, o, t, nm in zip(kinds, outs, targets, names))
    total.backward()
    check("loss", abs(float(total) - float(loss_tf)) / max(1.0, abs(float(loss_tf))))
    key_of = {}
    for layer in ref_model.layers:
        for v in layer.trainable_weights:
            key_of[v.ref()] = f"{layer.name}/{v.name.split('/')[-1].split(':')[0]}"
    gmax = max(float(np.abs(g.numpy()).max()) for g in grads_tf if g is not None)
    for v, g in zip(ref_model.trainable_variables, grads_tf):
        key = key_of[v.ref()]
        want = np.zeros(v.shape, np.float32) if g is None else g.numpy()
        got = tp[key].grad.numpy() if tp[key].grad is not None else np.zeros(v.shape)
        check(f"gradient {key}", float(np.abs(got - want).max()) / gmax)
    opt.apply_gradients([(g, v) for g, v in zip(grads_tf, ref_model.trainable_variables) if g is not None])
    for v in ref_model.trainable_variables:
        key = key_of[v.ref()]
        if tp[key].grad is None:
            continue
        w = tp[key].detach().clone()
        keras_adam_step(w, tp[key].grad, torch.zeros_like(w), torch.zeros_like(w), 1, lr=1e-3)
        check(f"adam {key}", float(np.abs(w.numpy() - v.numpy()).max()) / 1e-3)           # in units of the learning rate
    for layer in ref_model.layers:
        for v in layer.non_trainable_weights:
            key = f"{layer.name}/{v.name.split('/')[-1].split(':')[0]}"
            if key in k.new_moving:
                check(f"moving {key}", _rel(k.new_moving[key].numpy(), v.numpy()))
    return report


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", default=os.environ.get("B2SEG_REFERENCE", "/root/reference"))
    ap.add_argument("--case", action="append", choices=sorted(CASES))
    ap.add_argument("--tol", type=float, default=1e-4)
    a = ap.parse_args()
    bad = 0
    for name in a.case or sorted(CASES):
        print(f"[{name}]")
        rep = run_case(name, CASES[name], a.reference, a.tol)
        for what, err in rep["failures"]:
            print(f"  FAIL {what}: {err}")
        bad += len(rep["failures"])
    print("oracle pinned against the reference" if not bad else f"{bad} disagreements: the oracle's reading of Keras-2 is wrong where listed")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
