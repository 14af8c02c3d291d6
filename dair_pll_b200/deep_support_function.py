"""Homogeneous input-convex support-function network (learned convex geometry).

Mirror of ``dair_pll/deep_support_function.py:125-266`` (``HomogeneousICNN``): same parameters
(``hidden_weights`` / ``input_weights`` ParameterLists, ``output_weight``), same initialisation
distributions, same ``forward`` semantics -- the support point of the shape in direction d is the
input-Jacobian of f(d) = |w_out| . s(|W_h|^T s(W_d0^T d) + W_d1^T d), s = LeakyReLU(0.5), evaluated by
the hand-written Jacobian recursion of ``:238-266``.

Device path: :class:`ICNNSupport` is a ``torch.autograd.Function`` with an explicit forward and
backward.  The work is GEMM-shaped with weights shared by the whole batch -- two
(D x W x W) products forward, two backward -- and runs on cuBLAS DGEMM through ``torch.matmul``;
only the activation masks are kept between the passes (no autograd graph of elementwise ops).
A hand-written FP64 GEMM for this layer is future work; FP64 tensor cores on B200
have the same peak as the CUDA cores, so the library GEMM is already the right roofline.
Mesh extraction (``extract_mesh`` / ``extract_obj``, :19-122) is logging/export code and stays
with the reference.  Depth is fixed at 2 (the reference's default, ``geometry.py:50``).
"""
import math
from typing import List, Tuple

import torch
from torch import Tensor
from torch.nn import Module, Parameter, ParameterList


class ICNNSupport(torch.autograd.Function):
    """p (D,3) = d f / d direction for directions d (D,3); differentiable w.r.t. the four weights."""

    @staticmethod
    def forward(ctx, d, Wd0, Wd1, Wh, wout, slope):
        # deep_support_function.py:251-264 with hj = |w_out| * m1 folded into the small matrices, so the only
        # (D x width) intermediates are the two slope masks and a0
        if d.is_cuda:
            # fused memory-bound layers (csrc/cn_icnn.cu) around the FP64 GEMMs
            from dair_pll_b200 import ops
            p, h0aug, m1, a0 = ops.icnn_support_forward(d, Wd0, Wd1, Wh, wout, float(slope))
            ctx.save_for_backward(Wd0, Wd1, Wh, wout, h0aug, m1, a0)
            ctx.fused, ctx.slope = True, float(slope)
            return p
        ctx.fused = False
        Wh_a, wo = Wh.abs(), wout.abs()
        lin0 = d @ Wd0
        m0 = torch.where(lin0 > 0, 1.0, slope).to(d.dtype)
        lin1 = torch.addmm(d @ Wd1, lin0.mul_(m0), Wh_a)
        m1 = torch.where(lin1 > 0, 1.0, slope).to(d.dtype)
        a0 = (m1 @ (wo[:, None] * Wh_a.t())).mul_(m0)
        p = torch.addmm(m1 @ (wo[:, None] * Wd1.t()), a0, Wd0.t())
        ctx.save_for_backward(Wd0, Wd1, Wh, wout, m0, m1, a0)
        return p

    @staticmethod
    def backward(ctx, gp):
        Wd0, Wd1, Wh, wout, m0, m1, a0 = ctx.saved_tensors
        gp = gp.contiguous()
        Wh_a, wo = Wh.abs(), wout.abs()
        if ctx.fused:
            from dair_pll_b200 import ops
            g1, gWd0, G = ops.icnn_support_backward(gp, m0, m1, a0, Wd0, ctx.slope)     # m0 slot holds h0aug
        else:
            g1 = gp.t() @ m1                          # (3, width): d/dW_d1 before the |w_out| column scale
            gWd0 = gp.t() @ a0
            t = (gp @ Wd0).mul_(m0)                   # adjoint of (hj |W_h|^T); the masks are constants
            G = t.t() @ m1                            # (width, width): d/d|W_h| before the column scale
        # p is linear in |w_out|_j through column j of hj only: no third (D x width x width) product
        gwout = torch.sign(wout) * ((Wd1 * g1).sum(0) + (Wh_a * G).sum(0))
        return None, gWd0, g1 * wo, torch.sign(Wh) * (G * wo), gwout, None


class HomogeneousICNN(Module):
    """Positively homogeneous ICNN; ``forward(directions)`` returns support points (*, 3)."""

    def __init__(self, depth: int, width: int, negative_slope: float = 0.5, scale=1.0) -> None:
        assert 0.0 <= negative_slope < 1.0
        if depth != 2:
            raise NotImplementedError('the device path implements the reference default depth = 2')
        super().__init__()
        # same distributions as deep_support_function.py:166-186 (values are RNG dependent)
        scale_hidden = 2 * (2.0 / (1 + negative_slope ** 2)) ** 0.5 / width
        hidden = [Parameter(2 * (torch.rand((width, width), dtype=torch.float64) - 0.5) * scale_hidden)]
        inputs = []
        for layer in range(depth):
            w = torch.empty((3, width), dtype=torch.float64)
            torch.nn.init.kaiming_uniform_(w)
            if layer > 0:
                w = w * 2 ** (-0.5)
            inputs.append(Parameter(w))
        scale_out = float(scale) * 2 * (2.0 / (width * (1 + negative_slope ** 2))) ** 0.5
        self.hidden_weights = ParameterList(hidden)
        self.input_weights = ParameterList(inputs)
        self.output_weight = Parameter(2 * (torch.rand(width, dtype=torch.float64) - 0.5) * scale_out)
        self.negative_slope = negative_slope

    def abs_weights(self) -> Tuple[List[Tensor], Tensor]:
        return [w.abs() for w in self.hidden_weights], self.output_weight.abs()

    def forward(self, directions: Tensor) -> Tensor:
        shape = directions.shape
        # CUDA tensors always take the kernel path (float64 arithmetic, as the fp32 variant of the loss kernels);
        # CPU tensors run the same algebra in torch -- host mirror for CPU-only checks, never used by the CUDA path
        dt = torch.float64 if directions.is_cuda else directions.dtype
        p = ICNNSupport.apply(directions.reshape(-1, 3).to(dt), self.input_weights[0].to(dt), self.input_weights[1].to(dt),
                              self.hidden_weights[0].to(dt), self.output_weight.to(dt), self.negative_slope)
        return p.reshape(shape).to(directions.dtype)
