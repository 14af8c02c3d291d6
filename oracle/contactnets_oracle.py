"""ORACLE (test infrastructure, not product code).

Self-contained CPU restatement (batched fp64 PyTorch + oracle/cone_qp.c) of the
reference's ContactNets hot path, for use where ``/root/reference`` does not
exist (the GPU box).  It is pinned by ``tests/test_oracle_golden.py`` against
golden vectors that ``oracle/gen_golden.py`` produced by running the
reference's own Python through ``oracle/ref_shim.py``; the two un-vendored
boundaries inside it (symbolic callables, QP solver) remain "parity unpinned"
(see oracle/callables.py and oracle/cone_qp.py).

Follows, function by function:
  * theta -> pi_o -> pi_cm -> [m, c, I_cm/m]      dair_pll/inertia.py:205-234, 304-331, 376-382
  * LagrangianTerms.forward                      dair_pll/multibody_terms.py:214-237
  * ContactTerms.forward (+ friction combine)    dair_pll/multibody_terms.py:428-521, 321-324, 401-426
  * Box vertices / top-k support / plane-convex  dair_pll/geometry.py:162-202, 394-403, 553-582
  * contactnets_loss                             dair_pll/multibody_learnable_system.py:104-197
  * forward_dynamics / sim_step                  dair_pll/multibody_learnable_system.py:199-313
  * VelocityIntegrator.step, Integrator.simulate dair_pll/integrator.py:75-99, 153-162
  * FloatingBaseSpace.exponential                dair_pll/state_space.py:466-486
  * quaternion exp / multiply / sinc             dair_pll/quaternion.py:89-104, 208-229, 276-309
  * HomogeneousICNN.forward                      dair_pll/deep_support_function.py:238-266
"""
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from oracle.callables import TreeCallables, TreeSpec, skew
from oracle.cone_qp import OracleSAPSolver

LOSS_EPS = 1e-3   # multibody_learnable_system.py:130
STEP_EPS = 1e-4   # multibody_learnable_system.py:283,298
N_QUERY = 4       # geometry.py:48-49, 491
FORCE_CLIP = 1e3  # multibody_learnable_system.py:187

_CORNERS = torch.tensor([[sx, sy, sz] for sx in (-1., 1.) for sy in (-1., 1.) for sz in (-1., 1.)],
                        dtype=torch.float64)  # geometry.py:39-41 ordering


# --------------------------------------------------------------------------
# inertia parameterisation (inertia.py)
# --------------------------------------------------------------------------
def theta_to_inertia_vector(theta: Tensor) -> Tensor:
    """theta (n_b,10) -> the 10-vector handed to the generated callables:
    [m, c, Ixx,Iyy,Izz,Ixy,Ixz,Iyz] with I := I_cm / m (inertia.py:376-382)."""
    al, d1, d2, d3, s12, s23, s13, t1, t2, t3 = theta.unbind(-1)
    e1, e2, e3 = torch.exp(d1), torch.exp(d2), torch.exp(d3)
    scale = torch.exp(2 * al)
    m = scale * (t1 * t1 + t2 * t2 + t3 * t3 + 1)
    mc = scale[..., None] * torch.stack((t1 * e1, t1 * s12 + t2 * e2, t1 * s13 + t2 * s23 + t3 * e3), -1)
    Ixx = scale * (s12 * s12 + s23 * s23 + s13 * s13 + e2 * e2 + e3 * e3)
    Iyy = scale * (s13 * s13 + s23 * s23 + e1 * e1 + e3 * e3)
    Izz = scale * (s12 * s12 + e1 * e1 + e2 * e2)
    Ixy = scale * (-s12 * e1)
    Ixz = scale * (-s13 * e1)
    Iyz = scale * (-s12 * s13 - s23 * e2)
    c = mc / m[..., None]
    I_o = torch.stack((torch.stack((Ixx, Ixy, Ixz), -1), torch.stack((Ixy, Iyy, Iyz), -1),
                       torch.stack((Ixz, Iyz, Izz), -1)), -2)
    Sc = skew(c)
    I_cm = I_o + m[..., None, None] * (Sc @ Sc)   # parallel axis, origin -> com
    vec = torch.stack((I_cm[..., 0, 0], I_cm[..., 1, 1], I_cm[..., 2, 2],
                       I_cm[..., 0, 1], I_cm[..., 0, 2], I_cm[..., 1, 2]), -1)
    return torch.cat((m[..., None], c, vec / m[..., None]), -1)


def pi_cm_to_theta(pi_cm: Tensor) -> Tensor:
    """Inverse map used to initialise theta from URDF values (inertia.py:236-302, 334-360)."""
    m = pi_cm[..., 0]
    c = pi_cm[..., 1:4] / m[..., None]
    Ixx, Iyy, Izz, Ixy, Ixz, Iyz = pi_cm[..., 4:].unbind(-1)
    I_cm = torch.stack((torch.stack((Ixx, Ixy, Ixz), -1), torch.stack((Ixy, Iyy, Iyz), -1),
                        torch.stack((Ixz, Iyz, Izz), -1)), -2)
    Sc = skew(c)
    I_o = I_cm - m[..., None, None] * (Sc @ Sc)
    oxx, oyy, ozz = I_o[..., 0, 0], I_o[..., 1, 1], I_o[..., 2, 2]
    oxy, oxz, oyz = I_o[..., 0, 1], I_o[..., 0, 2], I_o[..., 1, 2]
    mc = m[..., None] * c
    ae1 = torch.sqrt(0.5 * (oyy + ozz - oxx))
    as12 = -oxy / ae1
    as13 = -oxz / ae1
    ae2 = torch.sqrt(ozz - ae1 ** 2 - as12 ** 2)
    as23 = (-oyz - as12 * as13) / ae2
    ae3 = torch.sqrt(oyy - ae1 ** 2 - as13 ** 2 - as23 ** 2)
    at1 = mc[..., 0] / ae1
    at2 = (mc[..., 1] - at1 * as12) / ae2
    at3 = (mc[..., 2] - at1 * as13 - at2 * as23) / ae3
    ea = torch.sqrt(m - at1 ** 2 - at2 ** 2 - at3 ** 2)
    return torch.stack((torch.log(ea), torch.log(ae1 / ea), torch.log(ae2 / ea), torch.log(ae3 / ea),
                        as12 / ea, as23 / ea, as13 / ea, at1 / ea, at2 / ea, at3 / ea), -1)


# --------------------------------------------------------------------------
# learnable parameters
# --------------------------------------------------------------------------
@dataclass
class OracleParams:
    """Same learnable tensors (names, shapes) as the reference module tree."""
    inertial_parameters: Tensor          # theta, (n_b, 10)
    friction_params: Tensor              # (n_g,), body geometries first, ground last
    length_params: List[Tensor]          # one (1,3) per Box geometry
    icnn: Optional[List[dict]] = None    # per DeepSupportConvex: dict(Wd0,Wd1,Wh,wout,perturbations)

    def leaves(self) -> List[Tensor]:
        out = [self.inertial_parameters, self.friction_params] + list(self.length_params)
        if self.icnn:
            for net in self.icnn:
                out += [net['Wd0'], net['Wd1'], net['Wh'], net['wout']]
        return out

    def requires_grad_(self, flag: bool = True):
        for t in self.leaves():
            t.requires_grad_(flag)
        return self


def icnn_support(net: dict, d: Tensor) -> Tensor:
    """Support point = input-Jacobian of the homogeneous ICNN (deep_support_function.py:238-266)."""
    Wh, wout = net['Wh'].abs(), net['wout'].abs()
    lrelu = torch.nn.functional.leaky_relu
    h0 = lrelu(d @ net['Wd0'], 0.5)
    h1 = lrelu(h0 @ Wh + d @ net['Wd1'], 0.5)
    m1 = torch.where(h1 > 0, torch.ones_like(h1), 0.5 * torch.ones_like(h1))
    m0 = torch.where(h0 > 0, torch.ones_like(h0), 0.5 * torch.ones_like(h0))
    a1 = wout * m1
    a0 = (a1 @ Wh.t()) * m0
    return a1 @ net['Wd1'].t() + a0 @ net['Wd0'].t()


def _support_points(params: OracleParams, g: int, d: Tensor) -> Tensor:
    """(*,3) direction in the geometry frame -> (*, 4, 3) witness points."""
    if params.icnn is not None and params.icnn[g] is not None:
        net = params.icnn[g]
        dirs = d[..., None, :] + net['perturbations']            # geometry.py:321-324
        dirs = dirs / dirs.norm(dim=-1, keepdim=True)
        return icnn_support(net, dirs)
    half = params.length_params[g].abs().reshape(3)
    verts = _CORNERS.to(d.dtype) * half                            # (8,3)
    dots = d @ verts.t()                                           # (*,8)
    sel = torch.topk(dots, N_QUERY, dim=-1, sorted=False).indices  # geometry.py:196-197
    sel, _ = torch.sort(sel, dim=-1)                               # canonical order (reference: unspecified)
    return verts[sel]


def contact_terms(calls: TreeCallables, params: OracleParams, q: Tensor) -> Tuple[Tensor, Tensor]:
    """phi (*, n_c), J (*, 3 n_c, n_v) in the reference's [J_n ; mu J_t interleaved] order."""
    tree = calls.tree
    R_WG = calls.geometry_rotations(q)
    p_WG = calls.geometry_translations(q)
    Jv_G = calls.geometry_spatial_jacobians(q)
    ground = len(tree.geometry_body) - 1
    mu_all = params.friction_params.abs()
    phis, Jn, Jt = [], [], []
    for g in range(ground):
        mu = 2 * mu_all[ground] * mu_all[g] / (mu_all[ground] + mu_all[g])
        R_AB = R_WG[..., ground, :, :].transpose(-1, -2) @ R_WG[..., g, :, :]
        p_AB = ((p_WG[..., g, :] - p_WG[..., ground, :])[..., None, :] @ R_WG[..., ground, :, :])[..., 0, :]
        d_B = -R_AB[..., 2, :]                                     # = -(R_BA)[:, 2]
        p_B = _support_points(params, g, d_B)                      # (*,4,3) in B
        p_A = p_B @ R_AB.transpose(-1, -2) + p_AB[..., None, :]
        phis.append(p_A[..., 2])
        p_W = p_B @ R_WG[..., g, :, :].transpose(-1, -2)           # contact point rel. geometry origin, world
        Jw = Jv_G[..., g, 0:3, :][..., None, :, :]
        Jl = Jv_G[..., g, 3:6, :][..., None, :, :]
        Jpt = Jl - skew(p_W) @ Jw                                  # (*,4,3,n_v), body B minus ground (zero)
        Jn.append(Jpt[..., 2, :])
        Jt.append(mu * Jpt[..., 0:2, :].reshape(Jpt.shape[:-3] + (2 * N_QUERY, Jpt.shape[-1])))
    return torch.cat(phis, -1), torch.cat(Jn + Jt, -2)


def multibody_terms(calls: TreeCallables, params: OracleParams, q: Tensor, v: Tensor):
    """M, J, phi, contact-free acceleration (multibody_terms.py:584-609, without the Delassus)."""
    inertia = theta_to_inertia_vector(params.inertial_parameters)
    inertia = inertia.expand(q.shape[:-1] + inertia.shape)
    M = calls.mass_matrix(q, inertia)
    F = calls.lagrangian_forces(q, v, None, inertia)
    acc = torch.linalg.solve(M, F[..., None])[..., 0]
    phi, J = contact_terms(calls, params, q)
    return M, J, phi, acc


def _to_sappy(vec_or_rows: Tensor, n_c: int, dim: int) -> Tensor:
    """[n..., (tx,ty)...] -> [(tx,ty,n)...] along ``dim`` (tensor_utils.py:460-497)."""
    idx = []
    for c in range(n_c):
        idx += [n_c + 2 * c, n_c + 2 * c + 1, c]
    return vec_or_rows.index_select(dim, torch.tensor(idx))


def _from_sappy(f_s: Tensor, n_c: int) -> Tensor:
    idx = [3 * c + 2 for c in range(n_c)]
    for c in range(n_c):
        idx += [3 * c, 3 * c + 1]
    return f_s.index_select(-1, torch.tensor(idx))


def _whiten(M: Tensor, J: Tensor) -> Tensor:
    """A with A A^T = J M^-1 J^T (any such factor gives the same QP, cf. :153-154)."""
    L = torch.linalg.cholesky(M)
    return torch.linalg.solve_triangular(L, J.transpose(-1, -2), upper=False).transpose(-1, -2)


def contactnets_loss(calls: TreeCallables, params: OracleParams, x: Tensor, x_plus: Tensor, dt: float,
                     solver: Optional[OracleSAPSolver] = None, return_force: bool = False):
    """(*,) ContactNets loss (multibody_learnable_system.py:104-197)."""
    solver = solver or OracleSAPSolver()
    n_q = calls.tree.n_q
    v = x[..., n_q:]
    q_plus, v_plus = x_plus[..., :n_q], x_plus[..., n_q:]
    M, J, phi, acc = multibody_terms(calls, params, q_plus, v_plus)
    n_c = phi.shape[-1]
    dv = v_plus - (v + dt * acc)
    Jt_v = (J[..., n_c:, :] @ v_plus[..., None])[..., 0]
    speeds = Jt_v.reshape(Jt_v.shape[:-1] + (n_c, 2)).norm(dim=-1)
    qvec = (-(J @ dv[..., None])[..., 0]
            + torch.cat((phi.abs(), torch.zeros_like(Jt_v)), -1)
            + dt * torch.cat((speeds, Jt_v), -1))
    const = 0.5 * (dv[..., None, :] @ M @ dv[..., None])[..., 0, 0] + (torch.clamp(-phi, min=0) ** 2).sum(-1)
    A = _to_sappy(_whiten(M, J), n_c, -2)
    f = _from_sappy(solver.apply(A, _to_sappy(qvec, n_c, -1), LOSS_EPS).detach(), n_c)
    bad = ((f.abs() > FORCE_CLIP) | f.isnan() | f.isinf()).any(-1)
    f = torch.where(bad[..., None], torch.zeros_like(f), f)
    const = torch.where(bad, torch.zeros_like(const), const)
    z = (J.transpose(-1, -2) @ f[..., None])[..., 0]
    y = torch.linalg.solve(M, z[..., None])[..., 0]
    loss = 0.5 * (z * y).sum(-1) + 0.5 * LOSS_EPS * (f * f).sum(-1) + (f * qvec).sum(-1) + const
    return (loss, f) if return_force else loss


PRIMAL_UPDATE = False   # see forward_dynamics


def forward_dynamics(calls: TreeCallables, params: OracleParams, q: Tensor, v: Tensor, dt: float,
                     solver: Optional[OracleSAPSolver] = None, return_force: bool = False):
    """Next velocity (multibody_learnable_system.py:199-304; the 1e6 contact filter is a no-op).

    ``PRIMAL_UPDATE`` (module switch, default off = the reference's form): evaluate the same update as
    v- + L^-T w with w = A^T f the primal optimum of the QP (A = J L^-T, L L^T = M, so M^-1 J^T f = L^-T A^T f).
    Identical in exact arithmetic; in floating point the reference's form loses ~cond(Q) eps ~ 1e-9 of the
    DERIVATIVES (f = Pi(-(A w + q)/eps) amplifies the rounding of w by 1/eps = 1e4), which is the resolution limit
    of any comparison against it.  Tests use the switch to show that a ~2e-9 disagreement with the reference
    form is that rounding and not an error of the kernels."""
    solver = solver or OracleSAPSolver()
    M, J, phi, acc = multibody_terms(calls, params, q, v)
    n_c = phi.shape[-1]
    v_minus = v + dt * acc
    q_full = (J @ v_minus[..., None])[..., 0] + torch.cat((phi, torch.zeros_like(phi).repeat_interleave(2, -1)), -1) / dt
    A = _to_sappy(_whiten(M, J), n_c, -2)
    if PRIMAL_UPDATE and not return_force:
        w = solver.apply(A, _to_sappy(q_full, n_c, -1), STEP_EPS, return_primal=True)
        L = torch.linalg.cholesky(M)
        return v_minus + torch.linalg.solve_triangular(L.transpose(-1, -2), w[..., None], upper=True)[..., 0]
    f = _from_sappy(solver.apply(A, _to_sappy(q_full, n_c, -1), STEP_EPS), n_c)
    v_next = v_minus + torch.linalg.solve(M, (J.transpose(-1, -2) @ f[..., None]))[..., 0]
    return (v_next, f) if return_force else v_next


def quat_exp(r: Tensor) -> Tensor:
    ang = r.norm(dim=-1, keepdim=True)
    half = ang / 2
    sinc = torch.where(half.abs() > 0, torch.sin(half) / torch.where(half.abs() > 0, half, torch.ones_like(half)),
                       torch.ones_like(half))
    return torch.cat((torch.cos(half), r * sinc / 2), -1)


def quat_mul(a: Tensor, b: Tensor) -> Tensor:
    aw, av, bw, bv = a[..., :1], a[..., 1:], b[..., :1], b[..., 1:]
    return torch.cat((aw * bw - (av * bv).sum(-1, keepdim=True),
                      aw * bv + bw * av + torch.cross(av, bv, dim=-1)), -1)


def sim_step(calls: TreeCallables, params: OracleParams, x: Tensor, dt: float,
             solver: Optional[OracleSAPSolver] = None) -> Tensor:
    """x -> x_next (integrator.py:153-162 + state_space.py:466-486). No quaternion renormalisation."""
    n_q = calls.tree.n_q
    q, v = x[..., :n_q], x[..., n_q:]
    v_next = forward_dynamics(calls, params, q, v, dt, solver)
    dq = v_next * dt
    quat = quat_mul(q[..., :4], quat_exp(dq[..., :3]))
    return torch.cat((quat, q[..., 4:] + dq[..., 3:], v_next), -1)


def simulate(calls: TreeCallables, params: OracleParams, x0: Tensor, dt: float, steps: int,
             solver: Optional[OracleSAPSolver] = None) -> Tensor:
    """(*, n_x) -> (*, steps+1, n_x) (integrator.py:75-99)."""
    traj = [x0]
    x = x0
    for _ in range(steps):
        x = sim_step(calls, params, x, dt, solver)
        traj.append(x)
    return torch.stack(traj, -2)
