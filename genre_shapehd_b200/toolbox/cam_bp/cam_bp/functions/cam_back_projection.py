"""CameraBackProjection — mirrors toolbox/cam_bp/cam_bp/functions/cam_back_projection.py:9-46.

Same call: ``CameraBackProjection.apply(depth_t, fl, cam_dist, res)`` -> tdf [N,C,res,res,res];
backward returns (grad_depth, grad_fl, grad_camdist, None), once-differentiable.
Differences that do not change results: the volumes are written once by the kernel (no zero_() / +1/res
passes), and the count volume is only materialised when a gradient can flow back.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .._ext import cam_bp_lib


def _forward(ctx, depth_t, fl, cam_dist, res, shift):
    assert depth_t.dim() == 4
    assert fl.dim() == 2 and fl.size(1) == depth_t.size(1)
    assert cam_dist.dim() == 2 and cam_dist.size(1) == depth_t.size(1)
    assert cam_dist.size(0) == depth_t.size(0)
    assert fl.size(0) == depth_t.size(0)
    assert depth_t.is_cuda
    assert fl.is_cuda
    assert cam_dist.is_cuda
    in_shape = depth_t.shape
    need_cnt = any(ctx.needs_input_grad[:3])
    tdf = depth_t.new_empty((in_shape[0], in_shape[1], res, res, res))
    cnt = torch.empty_like(tdf) if need_cnt else None
    cam_bp_lib.back_projection_forward(depth_t, cam_dist, fl, tdf, cnt, shift=shift)
    ctx.save_for_backward(depth_t, fl, cam_dist)
    ctx.cnt_forward = cnt
    ctx.depth_shape = in_shape
    ctx.res = res
    return tdf


def _backward(ctx, grad_output):
    assert grad_output.is_cuda
    depth_t, fl, cam_dist = ctx.saved_tensors
    cnt = ctx.cnt_forward
    grad_depth = grad_output.new_empty(ctx.depth_shape)
    grad_fl = grad_output.new_empty((ctx.depth_shape[0], ctx.depth_shape[1]))
    grad_camdist = grad_output.new_empty((ctx.depth_shape[0], ctx.depth_shape[1]))
    cam_bp_lib.back_projection_backward(
        depth_t, fl, cam_dist, cnt, grad_output.contiguous(), grad_depth, grad_camdist, grad_fl)
    return grad_depth, grad_fl, grad_camdist


class CameraBackProjection(Function):

    @staticmethod
    def forward(ctx, depth_t, fl, cam_dist, res=128):
        return _forward(ctx, depth_t, fl, cam_dist, res, shift=False)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return _backward(ctx, grad_output) + (None,)


class CameraBackProjectionShifted(Function):
    """CameraBackProjection followed by Camera_back_projection_layer.shift_tdf (1 - res * tdf),
    fused into the kernel's output stage (camera_backprojection_module.py:22-28)."""

    @staticmethod
    def forward(ctx, depth_t, fl, cam_dist, res=128):
        return _forward(ctx, depth_t, fl, cam_dist, res, shift=True)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        return _backward(ctx, grad_output * (-float(ctx.res))) + (None,)
