"""Mirror of the reference's cffi module ``nndistance._ext.my_lib`` (toolbox/nndistance/src/my_lib.h,
my_lib_cuda.h): same function names and argument orders, in-place outputs, bound to libgenre_b200.so.

The reference also ships a single-threaded CPU implementation (nnd_forward / nnd_backward, my_lib.c);
this package is CUDA-only by contract, so those two raise instead of computing on the host.
"""
import torch

from genre_shapehd_b200 import _lib


def _check(xyz1, xyz2):
    _lib.require_cuda(xyz1, xyz2)
    _lib.require_f32(xyz1, xyz2)
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.size(2) != 3 or xyz2.size(2) != 3 or xyz1.size(0) != xyz2.size(0):
        raise ValueError("expected xyz1 [B,N,3] and xyz2 [B,M,3], got %s and %s" % (tuple(xyz1.shape), tuple(xyz2.shape)))
    if not (xyz1.is_contiguous() and xyz2.is_contiguous()):
        raise ValueError("nndistance inputs must be contiguous")


def nnd_forward_cuda(xyz1, xyz2, dist1, dist2, idx1, idx2):
    """my_lib_cuda.h:1."""
    _check(xyz1, xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    assert dist1.shape == (b, n) and dist2.shape == (b, m) and idx1.shape == (b, n) and idx2.shape == (b, m)
    assert idx1.dtype == torch.int32 and idx2.dtype == torch.int32
    _lib.require_cuda(dist1, dist2, idx1, idx2)
    _lib.call("genre_b200_nnd_forward", xyz1.data_ptr(), xyz2.data_ptr(), b, n, m, dist1.data_ptr(),
              dist2.data_ptr(), idx1.data_ptr(), idx2.data_ptr(), _lib.stream_ptr(xyz1))
    return 1


def nnd_backward_cuda(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2):
    """my_lib_cuda.h:4.  gradxyz1 / gradxyz2 are fully overwritten."""
    _check(xyz1, xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    _lib.require_cuda(gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2)
    assert gradxyz1.shape == xyz1.shape and gradxyz2.shape == xyz2.shape
    assert gradxyz1.is_contiguous() and gradxyz2.is_contiguous()
    assert graddist1.is_contiguous() and graddist2.is_contiguous()
    _lib.call("genre_b200_nnd_backward", xyz1.data_ptr(), xyz2.data_ptr(), b, n, m, graddist1.data_ptr(),
              graddist2.data_ptr(), idx1.data_ptr(), idx2.data_ptr(), gradxyz1.data_ptr(), gradxyz2.data_ptr(),
              _lib.stream_ptr(xyz1))
    return 1


def nnd_forward(*args):
    raise RuntimeError("nndistance: CPU tensors are not supported by genre_shapehd_b200 (CUDA-only, no CPU fallback)")


def nnd_backward(*args):
    raise RuntimeError("nndistance: CPU tensors are not supported by genre_shapehd_b200 (CUDA-only, no CPU fallback)")
