"""Mirror of the reference's cffi module ``calc_prob._ext.calc_prob_lib``
(toolbox/calc_prob/calc_prob/src/calc_prob.h:1-2), bound to libgenre_b200.so."""
from genre_shapehd_b200 import _lib


def _check(*ts):
    _lib.require_cuda(*ts)
    _lib.require_f32(*ts)
    shape = ts[0].shape
    for t in ts:
        if t.dim() != 5:
            raise ValueError("5D input tensor expected but got: %s" % (tuple(t.shape),))
        if t.shape != shape:
            raise ValueError("shape mismatch: %s vs %s" % (tuple(t.shape), tuple(shape)))
        if not t.is_contiguous():
            raise ValueError("calc_prob tensors must be contiguous")


def calc_prob_forward(prob_in, stop_prob):
    """calc_prob.h:1.  stop_prob is fully overwritten (no pre-zeroing needed)."""
    _check(prob_in, stop_prob)
    z = prob_in.shape[4]
    _lib.call("genre_b200_calc_prob_forward", prob_in.data_ptr(), stop_prob.data_ptr(), prob_in.numel() // z, z,
              _lib.stream_ptr(prob_in))
    return 1


def calc_prob_backward(prob_in, stop_prob_weighted, grad_out):
    """calc_prob.h:2.  grad_out is fully overwritten."""
    _check(prob_in, stop_prob_weighted, grad_out)
    z = prob_in.shape[4]
    _lib.call("genre_b200_calc_prob_backward", prob_in.data_ptr(), stop_prob_weighted.data_ptr(),
              grad_out.data_ptr(), prob_in.numel() // z, z, _lib.stream_ptr(prob_in))
    return 1
