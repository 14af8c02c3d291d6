"""Homogeneous input-convex support-function network (learned convex geometry).

Mirror of ``dair_pll/deep_support_function.py:125-266`` (``HomogeneousICNN``): same parameters
(``hidden_weights`` / ``input_weights`` ParameterLists, ``output_weight``), same initialisation
distributions, same ``forward`` semantics -- the support point of the shape in direction d is the
input-Jacobian of f(d) = |w_out| . s(|W_h|^T s(W_d0^T d) + W_d1^T d), s = LeakyReLU(0.5), evaluated by
the hand-written Jacobian recursion of ``:238-266``.

Device path: :class:`ICNNSupport` is a ``torch.autograd.Function`` with an explicit forward and
backward.  The forward over all D rows is one tensor-core kernel (csrc/cn_icnn_tc.cu: the layer
Jacobians as exact int8 slice products on tcgen05, ``ops.icnn_support_points``); the backward visits only the rows
with a non-zero cotangent.
Depth 2 (the reference's default, ``geometry.py:50``, and every BASELINE configuration) is the kernel path.  Other depths
run the same Jacobian recursion layer by layer on the device as a composition of library products
(:func:`support_points_layerwise`): correct and differentiable, not tuned.
``extract_mesh`` (:95-122; the summary meshes of ``MultibodyTerms.scalars_and_meshes``, multibody_terms.py:566-580)
evaluates the network on the device and builds the hull on the host.
"""
import math
from typing import Callable, List, Sequence, Tuple

import torch
from torch import Tensor
from torch.nn import Module, Parameter, ParameterList

from dair_pll_b200.system import MeshSummary


def surface_directions() -> Tensor:
    """The 296 unit directions ``extract_mesh`` samples: the boundary cells of an 8 x 8 x 8 grid on [-1, 1]^3, normalised,
    in single precision as the reference's module constant (deep_support_function.py:13-16) so that the values agree."""
    axis = torch.linspace(-1, 1, steps=8)
    grid = torch.cartesian_prod(axis, axis, axis)
    shell = grid[grid.abs().max(dim=-1).values >= 1.0]
    return shell / shell.norm(dim=-1, keepdim=True)


def outward_normal_hyperplanes(vertices: Tensor, faces: Tensor):
    """(N, 3) vertices, (M, 3) triangle indices of a convex polytope -> (M, 3) outward unit normals, (M,) whether the
    face as listed winds clockwise seen from outside, (M,) plane offsets n . v_a  (the unbatched form of
    ``extract_outward_normal_hyperplanes``, deep_support_function.py:56-92)."""
    a, b, c = (vertices[faces[:, i]] for i in range(3))
    normals = torch.linalg.cross(b - a, c - a)
    normals = normals / normals.norm(dim=-1, keepdim=True)
    backwards = (normals * (a - vertices.mean(dim=0, keepdim=True))).sum(-1) < 0.0
    normals = torch.where(backwards[:, None], -normals, normals)
    return normals, backwards, (a * normals).sum(-1)


def extract_mesh(support_function: Callable[[Tensor], Tensor], device=None) -> MeshSummary:
    """Vertex / face mesh of the convex shape behind ``support_function`` (deep_support_function.py:95-122): its support
    points in :func:`surface_directions` (evaluated on ``device`` -- the network's own when it is a Module), exact
    duplicates dropped in first-occurrence order, the convex hull's triangles wound counter-clockwise seen from outside."""
    from scipy.spatial import ConvexHull
    if device is None and isinstance(support_function, Module):
        device = next(support_function.parameters()).device
    directions = surface_directions().to(torch.float64)
    if device is not None:
        directions = directions.to(device)
    with torch.no_grad():
        points = support_function(directions).detach().cpu()
    seen, keep = set(), []
    for i, row in enumerate(points.numpy()):
        key = row.tobytes()
        if key not in seen:
            seen.add(key)
            keep.append(i)
    vertices = points[keep]
    faces = torch.as_tensor(ConvexHull(vertices.numpy()).simplices, dtype=torch.long)
    _, backwards, _ = outward_normal_hyperplanes(vertices, faces)
    faces = torch.where(backwards[:, None], faces.flip(-1), faces)
    return MeshSummary(vertices=vertices, faces=faces)


def extract_obj(support_function: Callable[[Tensor], Tensor]) -> str:
    """Wavefront .obj text of :func:`extract_mesh` (deep_support_function.py:19-52): vertices, one normal per face, faces
    as ``v//n`` triples."""
    mesh = extract_mesh(support_function)
    normals, _, _ = outward_normal_hyperplanes(mesh.vertices, mesh.faces)
    lines = ['v ' + ' '.join(str(x.item()) for x in v) for v in mesh.vertices] + ['', '']
    lines += ['vn ' + ' '.join(str(x.item()) for x in n) for n in normals] + ['', '']
    lines += ['f ' + ' '.join(f'{i.item() + 1}//{k + 1}' for i in face) for k, face in enumerate(mesh.faces)]
    return '\n'.join(lines) + '\n'


def support_points_layerwise(d: Tensor, input_wts: Sequence[Tensor], hidden_wts: Sequence[Tensor], wout: Tensor,
                             slope: float) -> Tensor:
    """Support points of a homogeneous ICNN of ANY depth: (D, 3) directions -> (D, 3) input-Jacobian of
    f(d) = |w_out| . h_last,  h_0 = s(d W_d0),  h_l = s(h_{l-1} |W_h,l| + d W_d,l)  (deep_support_function.py:216-266).
    The network is piecewise linear, so the Jacobian is a product of constant matrices and 0/1 slope masks: forward
    pass for the masks (no gradient flows through them, as the reference's ``activation_jacobian``), then the reverse
    recursion  J_last = |w_out| o m_last,  p += J_l W_d,l^T,  J_{l-1} = (J_l |W_h,l|^T) o m_{l-1}.  Composed of library
    products on the device and differentiable with respect to every weight by autograd; the depth-2 networks of the
    BASELINE configurations never come here (:class:`ICNNSupport`)."""
    if not d.is_cuda:
        raise RuntimeError('dair_pll_b200 has no CPU path: support-function networks are evaluated on a CUDA device')
    hidden_abs = [w.abs() for w in hidden_wts]
    with torch.no_grad():
        masks = []
        h = torch.nn.functional.leaky_relu(d @ input_wts[0], slope)
        masks.append(torch.where(h > 0, 1.0, slope).to(d.dtype))
        for wh, wd in zip(hidden_abs, input_wts[1:]):
            h = torch.nn.functional.leaky_relu(h @ wh + d @ wd, slope)
            masks.append(torch.where(h > 0, 1.0, slope).to(d.dtype))
    jac = wout.abs()[None, :] * masks[-1]
    p = torch.zeros_like(d)
    for layer in range(len(hidden_abs), 0, -1):
        p = p + jac @ input_wts[layer].t()
        jac = (jac @ hidden_abs[layer - 1].t()) * masks[layer - 1]
    return p + jac @ input_wts[0].t()


def icnn_weight_gradients(Wd1: Tensor, Wh: Tensor, wout: Tensor, g1: Tensor, gWd0: Tensor, G: Tensor):
    """Chain rule from the three reductions the backward kernels produce -- g1 = gp^T m1 (3,W), gWd0 = gp^T a0 (3,W),
    G = t^T m1 (W,W) with t = (gp Wd0) o m0 -- to the gradients of (Wd0, Wd1, Wh, wout).  p is linear in |w_out|_j
    through column j of hj = |w_out| o m1 only, so d/d w_out needs no third (D x W x W) product."""
    Wh_a, wo = Wh.abs(), wout.abs()
    gwout = torch.sign(wout) * ((Wd1 * g1).sum(0) + (Wh_a * G).sum(0))
    return gWd0, g1 * wo, torch.sign(Wh) * (G * wo), gwout


class ICNNSupport(torch.autograd.Function):
    """p (D,3) = d f / d direction for directions d (D,3); differentiable w.r.t. the four weights.  CUDA only (there is
    no CPU path).

    Forward: every row.  Nothing of size (D x width) is kept for the backward: a row whose cotangent is exactly zero
    contributes exactly zero to every weight gradient, and in the ContactNets loss only the witness points of contacts
    that carry force or penetrate have a non-zero cotangent (3.7% of the rows of the config-3 batch,
    tools/exp_active_rows.py) -- so the backward gathers those rows, re-evaluates their activation masks and runs the
    reductions on them alone.  Width 256: both passes on the tensor cores (csrc/cn_icnn_tc.cu, cn_icnn_tc_bwd.cu), no host
    read; other widths: the FP64 layer kernels of csrc/cn_icnn.cu around library GEMMs (one host read for the row count)."""

    @staticmethod
    def forward(ctx, d, Wd0, Wd1, Wh, wout, slope):
        if not d.is_cuda:
            raise RuntimeError('dair_pll_b200 has no CPU path: support-function networks are evaluated on a CUDA device')
        from dair_pll_b200 import ops
        ctx.slope = float(slope)
        ctx.tensor_cores = Wd0.shape[1] == ops.ICNN_TC_WIDTH and not ops.ICNN_FORCE_FP64_PATH
        if ctx.tensor_cores:
            prepared = ops.icnn_tc_prepare(Wd0, Wd1, Wh, wout, ctx.slope)
            ctx.save_for_backward(d, Wd0, Wd1, Wh, wout, *prepared)
            return ops.icnn_support_points_tc(d, Wd0, Wd1, Wh, wout, ctx.slope, prepared)
        ctx.save_for_backward(d, Wd0, Wd1, Wh, wout)
        return ops.icnn_support_forward(d, Wd0, Wd1, Wh, wout, ctx.slope)[0]

    @staticmethod
    def backward(ctx, gp):
        from dair_pll_b200 import ops
        if ctx.tensor_cores:
            d, Wd0, Wd1, Wh, wout, image, consts = ctx.saved_tensors
            s = ctx.slope
            # no host read: compaction, mask re-evaluation and the (rows x W x W) contraction all take the live-row count
            # from device memory, so the whole training step can be captured into a CUDA graph
            C, R1, S = ops.icnn_support_backward_tc(d, gp.contiguous(), Wh, s, (image, consts))
            Wh_a, wo = Wh.abs(), wout.abs()
            g1 = s * S[:, None] + (1 - s) * R1                              # gp^T m1
            G = (Wd0[:, :, None] * C).sum(0)                                # t^T m1, t = (gp Wd0) o m0
            gWd0 = (C * (Wh_a * wo[None, :])[None]).sum(-1)                 # gp^T a0, a0 = (m1 (|wout| * |Wh|^T)) o m0
            return (None,) + icnn_weight_gradients(Wd1, Wh, wout, g1, gWd0, G) + (None,)
        d, Wd0, Wd1, Wh, wout = ctx.saved_tensors
        rows = torch.nonzero((gp != 0).any(-1)).reshape(-1)        # one host read (the row count sizes the launches)
        if rows.numel() < gp.shape[0]:
            d, gp = d.index_select(0, rows), gp.index_select(0, rows)
        _, h0aug, m1, a0 = ops.icnn_support_forward(d, Wd0, Wd1, Wh, wout, ctx.slope)
        g1, gWd0, G = ops.icnn_support_backward(gp.contiguous(), h0aug, m1, a0, Wd0, ctx.slope)
        return (None,) + icnn_weight_gradients(Wd1, Wh, wout, g1, gWd0, G) + (None,)


class HomogeneousICNN(Module):
    """Positively homogeneous ICNN; ``forward(directions)`` returns support points (*, 3)."""

    def __init__(self, depth: int, width: int, negative_slope: float = 0.5, scale=1.0) -> None:
        assert 0.0 <= negative_slope < 1.0 and depth >= 1
        super().__init__()
        # same distributions as deep_support_function.py:166-186 (values are RNG dependent)
        scale_hidden = 2 * (2.0 / (1 + negative_slope ** 2)) ** 0.5 / width
        hidden = [Parameter(2 * (torch.rand((width, width), dtype=torch.float64) - 0.5) * scale_hidden)
                  for _ in range(depth - 1)]
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
        # float64 arithmetic whatever the storage type (as the fp32 variant of the loss kernels)
        dt = torch.float64
        if len(self.input_weights) != 2:
            p = support_points_layerwise(directions.reshape(-1, 3).to(dt), [w.to(dt) for w in self.input_weights],
                                         [w.to(dt) for w in self.hidden_weights], self.output_weight.to(dt),
                                         self.negative_slope)
            return p.reshape(shape).to(directions.dtype)
        p = ICNNSupport.apply(directions.reshape(-1, 3).to(dt), self.input_weights[0].to(dt), self.input_weights[1].to(dt),
                              self.hidden_weights[0].to(dt), self.output_weight.to(dt), self.negative_slope)
        return p.reshape(shape).to(directions.dtype)
