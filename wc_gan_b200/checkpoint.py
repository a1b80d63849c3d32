"""Checkpoint interchange with the reference: Keras weight names <-> state_dict keys for the WC
layers, so that a generator trained with the reference can be evaluated through this implementation (inference path a9)
and vice versa.

The reference saves ``generator.save_weights(...)`` / ``model.save(...)`` HDF5 files (gan/train.py:62-73) whose datasets are
named ``<layer name>/<weight name>:0``.  For a stack created by ``create_norm(...)(axis, name)`` the Keras layers and weights
are (gan/conditional_layers.py:634-643,773-786,952-968; generator.py:23-78):

    <name>_npart        moving_mean (C,1) | moving_variance (C,C)
    <name>_repart       kernel, bias                     (uconv: Conv2D (1,1,C,C),(C,) | cconv: (K,1,1,C,C),(K,C))
    <name>_repart_c     kernel, bias | class_matrix, kernel   (ucconv / ufconv conditional branch; fconv has no bias)
    <name>_repart_u     kernel, bias                     (unconditional branch)

Shapes are identical on both sides (the modules keep Keras' shapes), so conversion is renaming only.  HDF5 I/O needs h5py,
which this image does not ship: ``load_keras_h5`` / ``save_keras_h5`` raise ImportError without it, and the same flat
``{keras name: array}`` mapping can be exchanged as an ``.npz`` file instead (``numpy.savez``).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
from torch import nn

from .layers import WhiteningColoring

_SUB = {'npart': '_npart', 'repart': '_repart', 'repart_c': '_repart_c', 'repart_u': '_repart_u'}


def keras_names(model: nn.Module) -> Dict[str, str]:
    """{state_dict key: Keras dataset name} for every WC stack inside ``model``."""
    out = {}
    for prefix, m in model.named_modules():
        if not isinstance(m, WhiteningColoring):
            continue
        base = m.layer_name or prefix
        for sub, suffix in _SUB.items():
            child = getattr(m, sub, None)
            if child is None:
                continue
            for key in child.state_dict().keys():
                sd_key = '.'.join(x for x in (prefix, sub, key) if x)
                out[sd_key] = f'{base}{suffix}/{key}:0'
    return out


def export_keras_weights(model: nn.Module) -> Dict[str, np.ndarray]:
    sd = model.state_dict()
    return {kname: sd[key].detach().cpu().numpy().copy() for key, kname in keras_names(model).items()}


def load_keras_weights(model: nn.Module, weights: Dict[str, np.ndarray], strict: bool = True) -> None:
    """Copy Keras-named arrays into the model's WC stacks (group_size must be None: the reference has no grouping)."""
    sd = model.state_dict()
    missing = []
    with torch.no_grad():
        for key, kname in keras_names(model).items():
            if kname not in weights:
                missing.append(kname)
                continue
            w = torch.as_tensor(np.asarray(weights[kname]), dtype=sd[key].dtype)
            if tuple(w.shape) != tuple(sd[key].shape):
                raise ValueError(f'{kname}: shape {tuple(w.shape)} does not match {tuple(sd[key].shape)}')
            sd[key].copy_(w)
    if strict and missing:
        raise KeyError(f'missing Keras weights: {missing}')


def save_npz(model: nn.Module, path: str) -> None:
    np.savez(path, **{k.replace('/', '|'): v for k, v in export_keras_weights(model).items()})


def load_npz(model: nn.Module, path: str, strict: bool = True) -> None:
    with np.load(path) as f:
        load_keras_weights(model, {k.replace('|', '/'): f[k] for k in f.files}, strict=strict)


def load_keras_h5(model: nn.Module, path: str, strict: bool = False) -> None:
    import h5py                                         # not in this image: raises ImportError
    weights = {}
    with h5py.File(path, 'r') as f:
        root = f['model_weights'] if 'model_weights' in f else f
        for lname in root:
            grp = root[lname]

            def visit(name, obj, lname=lname):
                if isinstance(obj, h5py.Dataset):
                    weights[f"{lname}/{name.split('/')[-1]}"] = np.asarray(obj)
            grp.visititems(visit)
    load_keras_weights(model, weights, strict=strict)


def save_keras_h5(model: nn.Module, path: str) -> None:
    import h5py
    with h5py.File(path, 'w') as f:
        for kname, arr in export_keras_weights(model).items():
            lname, wname = kname.split('/')
            f.require_group(lname).require_group(lname).create_dataset(wname, data=arr)
