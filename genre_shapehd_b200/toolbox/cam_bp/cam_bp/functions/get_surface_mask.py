"""get_vox_surface_cnt / get_surface_mask — mirror toolbox/cam_bp/cam_bp/functions/get_surface_mask.py:8-40."""
import torch

from .._ext import cam_bp_lib


def _as_map_param(v, n, nc, device):
    if type(v) == float:
        return torch.full((1, 1), v, dtype=torch.float32, device=device).expand(n, nc)
    return v


def get_vox_surface_cnt(depth_t, fl, cam_dist, res=128):
    assert depth_t.dim() == 4
    assert fl.dim() == 2 and fl.size(1) == depth_t.size(1)
    assert cam_dist.dim() == 2 and cam_dist.size(1) == depth_t.size(1)
    assert cam_dist.size(0) == depth_t.size(0)
    assert fl.size(0) == depth_t.size(0)
    assert depth_t.is_cuda
    assert fl.is_cuda
    assert cam_dist.is_cuda
    in_shape = depth_t.shape
    cnt = depth_t.new_empty((in_shape[0], in_shape[1], res, res, res))
    tdf = torch.empty_like(cnt)
    cam_bp_lib.back_projection_forward(depth_t, cam_dist, fl, tdf, cnt)
    return cnt


def get_surface_mask(depth_t, fl=784.4645406, cam_dist=2.0, res=128):
    n = depth_t.size(0)
    nc = depth_t.size(1)
    fl = _as_map_param(fl, n, nc, depth_t.device)
    cam_dist = _as_map_param(cam_dist, n, nc, depth_t.device)
    cnt = get_vox_surface_cnt(depth_t, fl, cam_dist, res)
    mask = torch.empty_like(cnt)
    cam_bp_lib.get_surface_mask(depth_t, cam_dist, fl, cnt, mask)
    surface_vox = torch.clamp(cnt, min=0.0, max=1.0)
    return surface_vox, mask
