"""Run a few training steps of one BASELINE config (for profilers: ncu / compute-sanitizer wrap this, not bench.py).
usage: python tools/run_step.py --config 3 [--steps 2] [--batch B] [--predict]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tf-1d-2d-segmentation-end2endpipelines_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--predict", action="store_true")
    a = ap.parse_args()
    if a.config == 2:
        from b2seg.model import Adam
        from b2seg.models2d import unet_model_builder
        B = a.batch or 32
        m = unet_model_builder("UNet", 256, 256, 64, 5, num_channels=3, output_nums=1, dense_loop=1, is_transconv=True, train_mode="from_scratch").ResNet50()
        m.compile(loss="binary_crossentropy", optimizer=Adam(2e-4))
        x, y = bench.synth_batch(B, 256, 2)
    else:
        m, x, y, _wl, _gf, B = bench.other_config(a.config, a.batch or None, 0)
    if a.predict:
        for _ in range(a.steps):
            m.predict(x, batch_size=B)
    else:
        for _ in range(a.steps):
            loss = m.train_on_batch(x, y)
        print("loss", loss)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
