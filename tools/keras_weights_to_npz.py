"""Convert a Keras weight file (.h5 from save_weights / model.save, or a .keras archive) into the .npz that b2seg's
Model.load_weights reads.  Run it where h5py is available (the reference's own environment has it); it needs nothing else of the
reference.  For .keras archives, whose variables carry no names, pass the builder call that makes the receiving model, e.g.

    python tools/keras_weights_to_npz.py best.keras best.npz --builder "unet_model_builder('UNet', 256, 256, 64, 5, train_mode='from_scratch').ResNet50()"
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tf-1d-2d-segmentation-end2endpipelines_b200"))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--builder", default=None, help="Python expression building the receiving b2seg model (names from b2seg.models1d / models2d)")
    a = ap.parse_args()
    from b2seg.keras_io import read_keras_weights, select_for_model
    specs = []
    if a.builder:
        import b2seg.models1d as m1
        import b2seg.models2d as m2
        ns = {k: getattr(m, k) for m in (m1, m2) for k in dir(m) if not k.startswith("_")}
        specs = eval(a.builder, ns).graph.param_specs()   # noqa: S307 (the user's own command line)
    found = read_keras_weights(a.src, specs)
    if specs:
        found, extra = select_for_model(found, specs)
        if extra:
            print(f"ignored {len(extra)} arrays the model does not have, e.g. {extra[:4]}")
    np.savez(a.dst, **{k.replace("/", "::"): v for k, v in found.items()})
    print(f"{a.dst}: {len(found)} arrays, {sum(v.size for v in found.values())} values")


if __name__ == "__main__":
    main()
