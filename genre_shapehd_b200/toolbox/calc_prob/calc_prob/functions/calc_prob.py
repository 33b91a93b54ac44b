"""CalcStopProb — mirrors toolbox/calc_prob/calc_prob/functions/calc_prob.py:9-29
(``CalcStopProb().apply(prob_in)`` / ``CalcStopProb.apply``)."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .._ext import calc_prob_lib


class CalcStopProb(Function):
    @staticmethod
    def forward(ctx, prob_in):
        assert prob_in.dim() == 5
        assert prob_in.dtype == torch.float32
        assert prob_in.is_cuda
        prob_in = prob_in.contiguous()
        stop_prob = torch.empty_like(prob_in)
        calc_prob_lib.calc_prob_forward(prob_in, stop_prob)
        ctx.save_for_backward(prob_in, stop_prob)
        return stop_prob

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_in):
        prob_in, stop_prob = ctx.saved_tensors
        grad_out = torch.empty_like(prob_in)
        stop_prob_weighted = stop_prob * grad_in
        calc_prob_lib.calc_prob_backward(prob_in, stop_prob_weighted.contiguous(), grad_out)
        return grad_out
