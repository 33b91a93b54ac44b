"""Dispatch of the 3D convolutions of networks/networks.py to the hand-written sm_100a kernels.

``conv3d(x, module)`` / ``conv_transpose3d(x, module)`` return the result computed by this library's kernels, or
``None`` when no kernel covers the layer (shape / dtype / device / autograd mode) — the caller then runs the layer
the way the reference does (torch.nn -> cuDNN).  Which layers are covered is stated in DESIGN.md.
"""
import os

import torch

ENABLED = os.environ.get("GENRE_B200_CONV", "1") != "0"


def conv3d(x, m):
    return None


def conv_transpose3d(x, m):
    return None
