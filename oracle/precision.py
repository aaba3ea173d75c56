"""Precision models for the network oracle (TEST INFRASTRUCTURE ONLY).

The reference computes everything in float32 (TF 1.12).  north_star prescribes bf16 tensor-core arithmetic with fp32
accumulation for the B200 path, so the CUDA path rounds activations, weights and activation gradients to bf16 at
well-defined points.  With batch-statistics BN on random-init weights those roundings are amplified
layer by layer, which makes a whole-step comparison against the exact oracle uninformative (a mis-wired tape and bf16
drift look the same).  `Bf16Faithful` makes the ORACLE round at the same points, so that what is left is accumulation
order noise: a whole-step gradient comparison then pins the wiring of the backward tape.

Hooks (all identity in `Exact`):
  w(x)   weights as the tensor cores see them: rounded forward, gradient passes through (fp32 master gradients)
  a(x)   an activation stored as bf16: rounded forward, and its gradient is stored as bf16 too (rounded backward)
  g(x)   a tensor kept in fp32 whose GRADIENT buffer is bf16 (head logits, translator heads, conv outputs before a
         leaky/ReLU mask, the activation before an x2 resize): identity forward, rounded backward
"""
import torch


def _bf16(x):
    return x.to(torch.bfloat16).to(x.dtype)


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class Exact:
    """The reference's arithmetic: no rounding anywhere."""
    name = "exact"

    def w(self, x):
        return x

    def a(self, x):
        return x

    def g(self, x):
        return x


class Bf16Faithful:
    """Rounds where the CUDA path rounds (see module docstring)."""
    name = "bf16"

    def w(self, x):
        return _RoundFwd.apply(x)

    def a(self, x):
        return _RoundBoth.apply(x)

    def g(self, x):
        return _RoundGrad.apply(x) if x.requires_grad else x
