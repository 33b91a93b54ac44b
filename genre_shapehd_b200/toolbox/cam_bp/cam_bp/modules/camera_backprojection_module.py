"""Camera_back_projection_layer — mirrors toolbox/cam_bp/cam_bp/modules/camera_backprojection_module.py:6-28.

``Camera_back_projection_layer()(depth_t, fl=418.3, cam_dist=2.2, shift=True)``.  Scalar intrinsics
become a cached one-element device tensor expanded with stride 0 (the kernels are stride-aware), so
no fill kernel runs per call; the shift ``1 - res * tdf`` is fused into the kernel's output stage.
"""
import torch
from torch import nn

from ..functions import CameraBackProjection
from ..functions.cam_back_projection import CameraBackProjectionShifted


class Camera_back_projection_layer(nn.Module):
    def __init__(self, res=128):
        super(Camera_back_projection_layer, self).__init__()
        assert res == 128
        self.res = 128
        self._scalar_cache = {}

    def _scalar(self, value, n, device):
        key = (float(value), device)
        t = self._scalar_cache.get(key)
        if t is None:
            t = torch.full((1, 1), float(value), dtype=torch.float32, device=device)
            self._scalar_cache[key] = t
        return t.expand(n, 1)

    def forward(self, depth_t, fl=418.3, cam_dist=2.2, shift=True):
        n = depth_t.size(0)
        if type(fl) == float:
            fl = self._scalar(fl, n, depth_t.device)
        if type(cam_dist) == float:
            cam_dist = self._scalar(cam_dist, n, depth_t.device)
        if shift:
            return CameraBackProjectionShifted.apply(depth_t, fl, cam_dist, self.res)
        return CameraBackProjection.apply(depth_t, fl, cam_dist, self.res)

    @staticmethod
    def shift_tdf(input_tdf, res=128):
        out_tdf = 1 - res * (input_tdf)
        return out_tdf


class camera_backprojection(nn.Module):
    """Kept importable for API parity; the reference version (:31-39) reads an undefined attribute
    (self.voxel_res) and can never have run.  This one does what it evidently meant."""

    def __init__(self, vox_res=128):
        super(camera_backprojection, self).__init__()
        self.vox_res = vox_res

    def forward(self, depth, fl, camdist):
        return CameraBackProjection.apply(depth, fl, camdist, self.vox_res)
