"""SphericalBackProjection — mirrors toolbox/cam_bp/cam_bp/functions/sperical_to_tdf.py:10-47.

``SphericalBackProjection().apply(spherical, grid, res)`` -> (tdf, cnt); backward takes
(grad_output, grad_phony) and returns (grad_depth, None, None).  The reference's NaN asserts
(:37,:42-46) force a device sync and ``np.isnan`` on a CUDA tensor fails on current PyTorch; they are
replaced by an opt-in check (GENRE_B200_CHECK_NAN=1) with the same failure (AssertionError).
"""
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .._ext import cam_bp_lib

_CHECK_NAN = os.environ.get("GENRE_B200_CHECK_NAN", "0") == "1"


class SphericalBackProjection(Function):

    @staticmethod
    def forward(ctx, spherical, grid, res=128):
        assert spherical.dim() == 4
        assert grid.dim() == 5
        assert spherical.size(0) == grid.size(0)
        assert spherical.size(1) == grid.size(1)
        assert spherical.size(2) == grid.size(2)
        assert spherical.size(3) == grid.size(3)
        assert grid.size(4) == 3
        assert spherical.is_cuda
        assert grid.is_cuda
        in_shape = spherical.shape
        tdf = spherical.new_empty((in_shape[0], in_shape[1], res, res, res))
        cnt = torch.empty_like(tdf)
        cam_bp_lib.spherical_back_proj_forward(spherical, grid, tdf, cnt)
        ctx.save_for_backward(spherical.detach(), grid, cnt)
        ctx.depth_shape = in_shape
        ctx.mark_non_differentiable(cnt)
        return tdf, cnt

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output, grad_phony):
        assert grad_output.is_cuda
        if _CHECK_NAN:
            assert not bool(torch.isnan(grad_output).any())
        spherical, grid, cnt = ctx.saved_tensors
        grad_depth = grad_output.new_empty(ctx.depth_shape)
        cam_bp_lib.spherical_back_proj_backward(spherical, grid, cnt, grad_output.contiguous(), grad_depth)
        if _CHECK_NAN:
            assert not bool(torch.isnan(grad_depth).any())
        return grad_depth, None, None
