"""Gradient reversal (reference ``functions.py:4-16``).

The fused train step folds the reversal into the backward as a ``-alpha`` factor on the classifier
branch; this autograd Function is kept for callers that compose the layer themselves."""
import torch
from torch.autograd import Function


class ReverseLayerF(Function):
    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.neg() * ctx.alpha, None
