"""gen_sph_grid / sph_pad / render_spherical — mirror toolbox/spherical_proj.py:6-72.

``render_spherical`` keeps the reference's registered buffers (``grid`` [S,S,Z,3] and
``depth_weight`` [Z]; they are part of depth_pred_with_sph_inpaint.Net's state_dict) but its forward
is ONE fused kernel (trilinear gather + clamp + stop-probability scan + expected depth, one warp per
ray) instead of grid_sample + clamp + CalcStopProb + matmul + prod over four [N,1,S,S,Z] tensors.
"""
import numpy as np
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from genre_shapehd_b200 import _lib
from .calc_prob.calc_prob.functions.calc_prob import CalcStopProb


def _sph_dirs(res):
    """Unit directions [res, res, 3] in fp64, exactly the arithmetic of spherical_proj.py:8-16."""
    pi = np.pi
    phi = np.linspace(0, 180, res * 2 + 1)[1::2]
    theta = np.linspace(0, 360, res + 1)[:-1]
    p = phi * pi / 180
    t = theta * pi / 180
    grid = np.zeros([res, res, 3])
    proj = np.sin(p)[:, None]
    grid[:, :, 2] = np.cos(p)[:, None]
    grid[:, :, 0] = proj * np.cos(t)[None, :]
    grid[:, :, 1] = proj * np.sin(t)[None, :]
    return grid


def gen_sph_grid(res=128):
    grid = np.reshape(_sph_dirs(res), (1, 1, res, res, 3))
    return torch.from_numpy(grid).float()


def sph_pad(sph_tensor, padding_margin=16):
    F = torch.nn.functional
    pad2d = (padding_margin, padding_margin, padding_margin, padding_margin)
    rep_padded_sph = F.pad(sph_tensor, pad2d, mode='replicate')
    _, _, h, w = rep_padded_sph.shape
    rep_padded_sph[:, :, :, 0:padding_margin] = rep_padded_sph[:, :, :, w - 2 * padding_margin:w - padding_margin]
    rep_padded_sph[:, :, :, h - padding_margin:] = rep_padded_sph[:, :, :, padding_margin:2 * padding_margin]
    return rep_padded_sph


def render_forward(vox, n, res, dirs64, sph_res, z_res, depth_weight, out, pre=None):
    """the fused renderer with empty-space skipping (csrc/render_sph.cu); pre = (scale, lo, hi) renders
    clamp(vox * scale, lo, hi) without materialising it"""
    nbytes = _lib.load().genre_b200_render_spherical_workspace_bytes(n, res)
    ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=vox.device)
    sc, lo, hi = pre if pre is not None else (1.0, 0.0, 0.0)
    _lib.call("genre_b200_render_spherical_forward_skip", vox.data_ptr(), n, res, dirs64.data_ptr(), sph_res, z_res,
              depth_weight.data_ptr(), 1 if pre is not None else 0, float(sc), float(lo), float(hi), out.data_ptr(),
              ws.data_ptr(), nbytes, _lib.stream_ptr(vox))
    return out


class _RenderSpherical(Function):
    @staticmethod
    def forward(ctx, vox, dirs64, depth_weight, sph_res, z_res):
        assert vox.dim() == 5 and vox.size(1) == 1, "render_spherical expects [N,1,R,R,R]"
        assert vox.size(2) == vox.size(3) == vox.size(4)
        _lib.require_cuda(vox, dirs64, depth_weight)
        _lib.require_f32(vox, depth_weight)
        vox = vox.contiguous()
        n, res = vox.size(0), vox.size(2)
        out = vox.new_empty((n, 1, sph_res, sph_res))
        render_forward(vox, n, res, dirs64, sph_res, z_res, depth_weight, out)
        ctx.save_for_backward(vox, dirs64, depth_weight)
        ctx.sph_res, ctx.z_res = sph_res, z_res
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        vox, dirs64, depth_weight = ctx.saved_tensors
        grad_vox = torch.zeros_like(vox)
        _lib.call("genre_b200_render_spherical_backward", vox.data_ptr(), vox.size(0), vox.size(2), dirs64.data_ptr(),
                  ctx.sph_res, ctx.z_res, depth_weight.data_ptr(), grad_out.contiguous().data_ptr(),
                  grad_vox.data_ptr(), _lib.stream_ptr(vox))
        return grad_vox, None, None, None, None


class render_spherical(torch.nn.Module):
    def __init__(self, sph_res=128, z_res=256):
        super().__init__()
        self.sph_res = sph_res
        self.z_res = z_res
        self.gen_grid()
        self.calc_stop_prob = CalcStopProb.apply
        self._dirs64 = {}

    def gen_grid(self):
        res = self.sph_res
        z_res = self.z_res
        dirs = _sph_dirs(res)
        self._dirs_np = dirs
        grid = np.reshape(dirs * 2, (res, res, 3))
        alpha = np.zeros([1, 1, z_res, 1])
        alpha[0, 0, :, 0] = np.linspace(0, 1, z_res)
        grid = grid[:, :, np.newaxis, :]
        grid = grid * (1 - alpha)
        grid = torch.from_numpy(grid).float()
        depth_weight = torch.linspace(0, 1, self.z_res)
        self.register_buffer('depth_weight', depth_weight)
        self.register_buffer('grid', grid)

    def _dirs_on(self, device):
        t = self._dirs64.get(device)
        if t is None:
            t = torch.from_numpy(np.ascontiguousarray(self._dirs_np)).to(device)
            self._dirs64[device] = t
        return t

    def forward(self, vox):
        return _RenderSpherical.apply(vox, self._dirs_on(vox.device), self.depth_weight, self.sph_res, self.z_res)

    def forward_unfused(self, vox):
        """The reference's op-by-op composition (spherical_proj.py:62-72) on this package's
        CalcStopProb; kept for tests of the fused kernel against the composed path."""
        grid = self.grid.expand(vox.shape[0], -1, -1, -1, -1)
        vox = vox.permute(0, 1, 4, 3, 2)
        prob_sph = torch.nn.functional.grid_sample(vox, grid, mode='bilinear', padding_mode='zeros',
                                                   align_corners=True)
        prob_sph = torch.clamp(prob_sph, 1e-5, 1 - 1e-5)
        sph_stop_prob = self.calc_stop_prob(prob_sph)
        exp_depth = torch.matmul(sph_stop_prob, self.depth_weight)
        back_groud_prob = torch.prod(1.0 - prob_sph, dim=4)
        back_groud_prob = back_groud_prob * 1.0
        exp_depth = exp_depth + back_groud_prob
        return exp_depth
