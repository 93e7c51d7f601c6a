"""Deep-supervision target pyramids — the reference's `prepareTrainDict` helpers, host side.

2D: TensorFlow/2DCNN/utils/helper_functions.py:359-380 (called per batch by the data generator, DataGenerator.py:113):
    'out' = the mask; 'level{i}' = MaxPooling2D(2^i) of the mask for model_type 'UNet' (levels taken before each
    up-sampling), the mask itself for 'UNetPP' (full-resolution levels).
1D: TensorFlow/1DCNN/1D_Segmentation.ipynb cell 31: the same with a window MEAN over 2^i samples.

`Model.compile(..., ds_targets='UNet' | 'UNetPP')` does the same on the device (csrc/stream_kernels2.cu: target_pool_kernel,
C ABI b2seg_target_pool), so that only the mask crosses PCIe; these NumPy versions serve callers that keep the reference's
data generator, and `Model.evaluate`.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def _window_reduce(a: np.ndarray, ph: int, pw: int, how: str) -> np.ndarray:
    """a: (N, H, W, C) -> (N, H // ph, W // pw, C); trailing rows / columns that do not fill a window are dropped ('valid')"""
    N, H, W, C = a.shape
    Ho, Wo = H // ph, W // pw
    v = a[:, :Ho * ph, :Wo * pw].reshape(N, Ho, ph, Wo, pw, C)
    return v.max(axis=(2, 4)) if how == "max" else v.mean(axis=(2, 4))


def prepareTrainDict(image_batch, model_depth: int, model_type: str) -> Dict[str, np.ndarray]:
    """2D (helper_functions.py:359-380).  Like the reference, an unknown model_type raises KeyError on the first level."""
    image_batch = np.array(image_batch)
    if image_batch.ndim == 3:
        image_batch = np.expand_dims(image_batch, axis=3)
    out = {"out": image_batch}
    for i in range(1, model_depth + 1):
        if model_type == "UNet":
            out[f"level{i}"] = _window_reduce(image_batch, 2 ** i, 2 ** i, "max")
        elif model_type == "UNetPP":
            out[f"level{i}"] = image_batch
        else:
            raise KeyError(f"level{i}")
    return out


def prepareTrainDict1D(y, model_depth: int, signal_length: int, model_name: str, num_channel: int = 1) -> Dict[str, np.ndarray]:
    """1D (1D_Segmentation.ipynb cell 31): window mean; y is (N, signal_length, num_channel)"""
    y = np.array(y)
    out = {"out": y}
    for i in range(1, model_depth + 1):
        if model_name == "UNet":
            w = 2 ** i
            out[f"level{i}"] = _window_reduce(y[:, None, :signal_length, :num_channel].astype(np.float64), 1, w, "mean")[:, 0]
        elif model_name == "UNetPP":
            out[f"level{i}"] = y
        else:
            raise KeyError(f"level{i}")
    return out


def derive_targets_host(mask: np.ndarray, shapes, ndim: int):
    """targets of every output from the mask alone: same shape -> the mask, smaller -> window max (2D) / mean (1D); mask and
    shapes are (N, H, W, C) (1D: H = 1)"""
    outs = []
    for shp in shapes:
        if tuple(shp) == tuple(mask.shape):
            outs.append(mask)
            continue
        ph, pw = mask.shape[1] // shp[1], mask.shape[2] // shp[2]
        if shp[3] != mask.shape[3] or ph * shp[1] != mask.shape[1] or pw * shp[2] != mask.shape[2]:
            raise ValueError(f"cannot derive a target of shape {tuple(shp)} from a mask of shape {tuple(mask.shape)}")
        outs.append(_window_reduce(mask, ph, pw, "max" if ndim == 2 else "mean").astype(np.float32))
    return outs
