"""GPU versions of the GenRe data pipeline's per-sample CPU post-processing.

``voxel_surface`` is the voxel branch of ``Model.preprocess`` (models/genre_full_model.py:86-96) for a whole batch resident on
the device: the reference runs ``scipy.ndimage.binary_erosion`` on every sample inside its DataLoader workers.
"""
import torch

from . import _lib


def voxel_surface(voxel, iterations=2, transpose_flip=True):
    """voxel [B,1,R,R,R] or [B,R,R,R] fp32 on CUDA -> surface voxels of the same shape:
         val = flip(transpose(v, (0,2,1)), 2)   per sample (transpose_flip, genre_full_model.py:89-90)
         clip(val - binary_erosion(val != 0, ones((3,3,3)), iterations), 0, 1)            (:91-93)
    Bit-exact against scipy.  R must be a multiple of 32 (the reference uses 128)."""
    _lib.require_cuda(voxel)
    _lib.require_f32(voxel)
    shape = voxel.shape
    v = voxel.reshape(-1, shape[-3], shape[-2], shape[-1]).contiguous()
    n, r = v.shape[0], v.shape[1]
    if not (v.shape[1] == v.shape[2] == v.shape[3]):
        raise ValueError("voxel_surface expects cubic volumes, got %s" % (tuple(shape),))
    nbytes = _lib.load().genre_b200_voxel_surface_workspace_bytes(n, r)
    if nbytes == 0:
        raise ValueError("voxel_surface: resolution %d unsupported (a multiple of 32, at most 256)" % r)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=v.device)
    out = torch.empty_like(v)
    _lib.call("genre_b200_voxel_surface", v.data_ptr(), n, r, int(iterations), 1 if transpose_flip else 0, out.data_ptr(),
              ws.data_ptr(), nbytes, _lib.stream_ptr(v))
    return out.view(shape)
