"""ORACLE (test infrastructure, not product code).

Restatement, in batched fp64 PyTorch, of the five callables that the reference
obtains from the un-vendored ``drake_pytorch.sym_to_pytorch`` code generator
applied to pydrake symbolic expressions (un-pinned third-party dependencies,
``setup.py:32-41``; call sites ``dair_pll/multibody_terms.py:134-153`` and
``:299-310``):

    mass_matrix(q, inertia)                  multibody_terms.py:131-138
    lagrangian_forces(q, v, u, inertia)      multibody_terms.py:140-153
    geometry_rotations(q)                    multibody_terms.py:299-300, 365
    geometry_translations(q)                 multibody_terms.py:302-303, 367
    geometry_spatial_jacobians(q)            multibody_terms.py:305-310, 369-376

PARITY UNPINNED at this boundary: neither pydrake nor drake_pytorch can be
imported here and the reference ships no tests.  The expressions those tools
generate are the rigid-body equations of motion of the URDF's kinematic tree in
the reference's state coordinates (``drake_state_converter.py:51-73``:
quaternion w-first, world position, BODY-frame angular velocity, WORLD-frame
linear velocity of the body origin, joint angle/rate), so they are determined by
physics; tests/ check them by identities (M symmetric PD and equal to the
Hessian of the kinetic energy, energy and momentum conservation in free flight,
Jacobians equal to finite differences of the point kinematics).

Model class covered: one floating-base kinematic tree whose other joints are
revolute (enough for both shipped assets, contactnets_cube.urdf and
contactnets_elbow.urdf).  ``inertia`` rows are the reference's "drake spatial
inertia" 10-vectors [m, c(3), Ixx, Iyy, Izz, Ixy, Ixz, Iyz] that the generated
code treats *literally* as (mass, com offset, central rotational inertia), cf.
``multibody_terms.py:198-201`` -- see SURVEY.md section 7 hard part 2.
"""
from dataclasses import dataclass, field
from typing import List, Tuple

import torch
from torch import Tensor

GRAVITY = 9.81  # drake default UniformGravityFieldElement, -z


def skew(p: Tensor) -> Tensor:
    z = torch.zeros_like(p[..., 0])
    return torch.stack((torch.stack((z, -p[..., 2], p[..., 1]), -1),
                        torch.stack((p[..., 2], z, -p[..., 0]), -1),
                        torch.stack((-p[..., 1], p[..., 0], z), -1)), -2)


def quat_to_rot(quat: Tensor) -> Tensor:
    """Rotation matrix of a (not necessarily unit) quaternion, w first.

    Drake builds RotationMatrix(quaternion) with the 2/|q|^2 scaling, so the
    result is orthonormal for any non-zero quaternion."""
    w, x, y, z = quat.unbind(-1)
    s = 2.0 / (w * w + x * x + y * y + z * z)
    return torch.stack((
        torch.stack((1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y)), -1),
        torch.stack((s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x)), -1),
        torch.stack((s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)), -1)), -2)


def axis_rot(axis: Tensor, angle: Tensor) -> Tensor:
    """Rodrigues rotation about a fixed unit axis by a batched angle."""
    K = skew(axis)
    eye = torch.eye(3, dtype=angle.dtype)
    s = torch.sin(angle)[..., None, None]
    c = torch.cos(angle)[..., None, None]
    return eye + s * K + (1 - c) * (K @ K)


@dataclass
class TreeSpec:
    """Kinematic tree: body 0 floats; body i>0 hangs off ``parent[i]`` by a
    revolute joint at ``joint_origin[i]`` (parent frame) about ``axis[i]``.
    ``geometry_body[g]``/``geometry_offset[g]`` place the collision geometries
    (body index -1 = world)."""
    parent: List[int]
    joint_origin: List[Tuple[float, float, float]]
    axis: List[Tuple[float, float, float]]
    geometry_body: List[int] = field(default_factory=list)
    geometry_offset: List[Tuple[float, float, float]] = field(default_factory=list)
    joint_rpy: List[Tuple[float, float, float]] = field(default_factory=list)   # fixed rotation parent -> joint frame
                                                                                # (URDF <origin rpy>); empty = none
    geometry_rpy: List[Tuple[float, float, float]] = field(default_factory=list)  # fixed rotation link -> collision frame
                                                                                # (URDF <collision><origin rpy>); empty = none
    prismatic: List[bool] = field(default_factory=list)    # joint i slides along its axis instead of turning about it
                                                           # (URDF type="prismatic"); empty = all revolute

    def is_prismatic(self, i: int) -> bool:
        return bool(self.prismatic) and bool(self.prismatic[i])

    def geometry_rotation(self, g: int, dtype) -> Tensor:
        """Fixed rotation of geometry g's collision frame in its link (R = Rz(yaw) Ry(pitch) Rx(roll))."""
        if not self.geometry_rpy:
            return torch.eye(3, dtype=dtype)
        r, p, y = (torch.tensor(v, dtype=dtype) for v in self.geometry_rpy[g])
        ex, ey, ez = (torch.tensor(v, dtype=dtype) for v in ((1., 0., 0.), (0., 1., 0.), (0., 0., 1.)))
        return axis_rot(ez, y) @ axis_rot(ey, p) @ axis_rot(ex, r)

    def joint_rotation(self, i: int, dtype) -> Tensor:
        """Fixed rotation of joint i's frame in its parent link (URDF convention R = Rz(yaw) Ry(pitch) Rx(roll))."""
        if not self.joint_rpy:
            return torch.eye(3, dtype=dtype)
        r, p, y = (torch.tensor(v, dtype=dtype) for v in self.joint_rpy[i])
        ex, ey, ez = (torch.tensor(v, dtype=dtype) for v in ((1., 0., 0.), (0., 1., 0.), (0., 0., 1.)))
        return axis_rot(ez, y) @ axis_rot(ey, p) @ axis_rot(ex, r)

    @property
    def n_bodies(self) -> int:
        return len(self.parent)

    @property
    def n_joints(self) -> int:
        return len(self.parent) - 1

    @property
    def n_q(self) -> int:
        return 7 + self.n_joints

    @property
    def n_v(self) -> int:
        return 6 + self.n_joints


# contactnets_cube.urdf: one body, box geometry at the body origin, ground.
CUBE_TREE = TreeSpec(parent=[-1], joint_origin=[(0., 0., 0.)], axis=[(0., 0., 1.)],
                     geometry_body=[0, -1], geometry_offset=[(0., 0., 0.), (0., 0., 0.)])
# contactnets_elbow.urdf:35-71: hinge at (-0.035, 0.06, 0) about y; child box at (0.035,0,0).
ELBOW_TREE = TreeSpec(parent=[-1, 0], joint_origin=[(0., 0., 0.), (-0.035, 0.06, 0.)],
                      axis=[(0., 0., 1.), (0., 1., 0.)],
                      geometry_body=[0, 1, -1],
                      geometry_offset=[(0., 0., 0.), (0.035, 0., 0.), (0., 0., 0.)])


# a three-link chain with an off-axis, rotated second joint (tests of the generic chain kernels; not a reference asset)
# one floating body whose collision geometry sits in a frame of its own, offset and rotated in the link (tests of the
# collision-frame handling of the single-body path; not a reference asset)
FRAMED_BODY_TREE = TreeSpec(parent=[-1], joint_origin=[(0., 0., 0.)], axis=[(0., 0., 1.)],
                            geometry_body=[0, -1], geometry_offset=[(0.012, -0.02, 0.015), (0., 0., 0.)],
                            geometry_rpy=[(0.3, -0.2, 0.5), (0., 0., 0.)])


CHAIN3_TREE = TreeSpec(parent=[-1, 0, 1],
                       joint_origin=[(0., 0., 0.), (-0.035, 0.06, 0.), (0.07, 0.01, -0.02)],
                       axis=[(0., 0., 1.), (0., 1., 0.), (0.6, 0., 0.8)],
                       geometry_body=[0, 1, 2, -1],
                       geometry_offset=[(0., 0., 0.), (0.035, 0., 0.), (0.03, -0.01, 0.), (0., 0., 0.)],
                       joint_rpy=[(0., 0., 0.), (0., 0., 0.), (0.3, -0.2, 0.5)])


# CHAIN3_TREE with every box in a collision frame ROTATED against its link (URDF <collision><origin rpy>)
CHAIN3R_TREE = TreeSpec(parent=CHAIN3_TREE.parent, joint_origin=CHAIN3_TREE.joint_origin, axis=CHAIN3_TREE.axis,
                        geometry_body=CHAIN3_TREE.geometry_body, geometry_offset=CHAIN3_TREE.geometry_offset,
                        joint_rpy=CHAIN3_TREE.joint_rpy,
                        geometry_rpy=[(0.4, 0., 0.), (0.1, -0.3, 0.2), (0., 0.5, -0.4), (0., 0., 0.)])


# three links: a hinge, then a SLIDING (prismatic) joint in a rotated joint frame (tests of the prismatic branch of the generic
# tree kernels and of this module's own prismatic kinematics; not a reference asset)
SLIDER3_TREE = TreeSpec(parent=[-1, 0, 1],
                        joint_origin=[(0., 0., 0.), (-0.035, 0.06, 0.), (0.05, 0.01, -0.02)],
                        axis=[(0., 0., 1.), (0., 1., 0.), (0.6, 0., 0.8)],
                        geometry_body=[0, 1, 2, -1],
                        geometry_offset=[(0., 0., 0.), (0.035, 0., 0.), (0.02, -0.01, 0.), (0., 0., 0.)],
                        joint_rpy=[(0., 0., 0.), (0., 0., 0.), (0.3, -0.2, 0.5)],
                        prismatic=[False, False, True])


# a branching four-link tree: links 1 and 2 hang off the root, link 3 off link 2 (rotated joint frames on two joints); tests
# of the generic tree kernels, not a reference asset
TREE4_TREE = TreeSpec(parent=[-1, 0, 0, 2],
                      joint_origin=[(0., 0., 0.), (-0.035, 0.06, 0.), (0.05, -0.04, 0.01), (0.06, 0.0, -0.015)],
                      axis=[(0., 0., 1.), (0., 1., 0.), (1., 0., 0.), (0.6, 0., 0.8)],
                      geometry_body=[0, 1, 2, 3, -1],
                      geometry_offset=[(0., 0., 0.), (0.035, 0., 0.), (0.0, -0.03, 0.), (0.03, -0.01, 0.), (0., 0., 0.)],
                      joint_rpy=[(0., 0., 0.), (0., 0., 0.), (0.2, 0.1, -0.3), (0.3, -0.2, 0.5)])


# TREE4_TREE's kinematics with the boxes distributed UNEVENLY over the links: two on the root, none on links 1 and 2, one on
# link 3 (tests of the box-slot handling of the generic tree kernels; not a reference asset)
TREE4G_TREE = TreeSpec(parent=[-1, 0, 0, 2],
                       joint_origin=[(0., 0., 0.), (-0.035, 0.06, 0.), (0.05, -0.04, 0.01), (0.06, 0.0, -0.015)],
                       axis=[(0., 0., 1.), (0., 1., 0.), (1., 0., 0.), (0.6, 0., 0.8)],
                       geometry_body=[0, 0, 3, -1],
                       geometry_offset=[(0.02, 0., 0.), (-0.04, 0.01, 0.015), (0.03, -0.01, 0.), (0., 0., 0.)],
                       joint_rpy=[(0., 0., 0.), (0., 0., 0.), (0.2, 0.1, -0.3), (0.3, -0.2, 0.5)],
                       geometry_rpy=[(0., 0., 0.), (0.3, 0., 0.2), (0., 0., 0.), (0., 0., 0.)])


# a six-link tree: three limbs off the root, two of them with a second segment (tests of the generic tree kernels at their
# largest instantiation; not a reference asset)
TREE6_TREE = TreeSpec(parent=[-1, 0, 0, 0, 1, 3],
                      joint_origin=[(0., 0., 0.), (0.05, 0.04, 0.), (-0.05, 0.04, 0.), (0.0, -0.05, 0.01),
                                    (0.06, 0.0, -0.01), (0.0, -0.06, 0.0)],
                      axis=[(0., 0., 1.), (0., 1., 0.), (1., 0., 0.), (0., 0., 1.), (0.6, 0., 0.8), (0., 1., 0.)],
                      geometry_body=[0, 1, 2, 3, 4, 5, -1],
                      geometry_offset=[(0., 0., 0.), (0.03, 0., 0.), (-0.03, 0., 0.), (0., -0.03, 0.), (0.03, -0.01, 0.),
                                       (0., -0.03, 0.01), (0., 0., 0.)],
                      joint_rpy=[(0., 0., 0.), (0., 0., 0.3), (0.2, 0.1, -0.3), (0., 0., 0.), (0.3, -0.2, 0.5), (0.1, 0., 0.)])


class TreeCallables:
    """The five callables for a :class:`TreeSpec` (see module docstring)."""

    def __init__(self, tree: TreeSpec):
        self.tree = tree

    # -- kinematics -------------------------------------------------------
    def kinematics(self, q: Tensor, v: Tensor = None):
        """Per body: world rotation R_i, origin o_i, velocity maps (Tw_i, Tv_i)
        with omega_i(body frame) = Tw_i v, vel(origin, world) = Tv_i v, and, if
        ``v`` is given, body twists and the bias terms (d/dt T_i) v."""
        tree = self.tree
        nv = tree.n_v
        batch = q.shape[:-1]
        dt = q.dtype
        R0 = quat_to_rot(q[..., :4])
        eye6 = torch.eye(nv, dtype=dt)
        Tw0 = eye6[0:3].expand(batch + (3, nv))
        Tv0 = eye6[3:6].expand(batch + (3, nv))
        R, o, Tw, Tv = [R0], [q[..., 4:7]], [Tw0], [Tv0]
        om, al, be = [], [], []
        if v is not None:
            om.append(v[..., 0:3])
            al.append(torch.zeros(batch + (3,), dtype=dt))
            be.append(torch.zeros(batch + (3,), dtype=dt))
        for i in range(1, tree.n_bodies):
            p = tree.parent[i]
            a = torch.tensor(tree.axis[i], dtype=dt)
            pj = torch.tensor(tree.joint_origin[i], dtype=dt)
            if tree.is_prismatic(i):
                # sliding joint: the child keeps the joint frame's orientation and moves by d = q_i along the axis;
                #   o_i = o_p + R_p (pJ + Rfix a d),  w_i = w_p,  v_i = v_p + w_p x (o_i - o_p) + (R_p Rfix a) d'
                # velocity-product acceleration of the origin: ... + w_p x (w_p x r) + 2 w_p x (R_p Rfix a) d'
                Rfix = tree.joint_rotation(i, dt)
                RjT = Rfix.transpose(-1, -2).expand(batch + (3, 3))
                e = eye6[5 + i]
                af = Rfix @ a                                         # axis in the parent link's frame
                rp = pj + af * q[..., 6 + i][..., None]               # (*, 3) child origin in the parent frame
                R.append(R[p] @ Rfix)
                o.append(o[p] + (R[p] @ rp[..., None])[..., 0])
                Tw.append(RjT @ Tw[p])
                Tv.append(Tv[p] - R[p] @ skew(rp) @ Tw[p] + (R[p] @ af)[..., None] * e)
                if v is not None:
                    rate = v[..., 5 + i]
                    w_p = om[p]
                    om.append((RjT @ w_p[..., None])[..., 0])
                    al.append((RjT @ al[p][..., None])[..., 0])
                    wxwxr = torch.cross(w_p, torch.cross(w_p, rp, dim=-1), dim=-1)
                    cor = 2 * rate[..., None] * torch.cross(w_p, af.expand_as(w_p), dim=-1)
                    be.append(be[p] + (R[p] @ (wxwxr - torch.cross(rp, al[p], dim=-1) + cor)[..., None])[..., 0])
                continue
            Rj = tree.joint_rotation(i, dt) @ axis_rot(a, q[..., 6 + i])
            RjT = Rj.transpose(-1, -2)
            e = eye6[5 + i]
            R.append(R[p] @ Rj)
            o.append(o[p] + (R[p] @ pj))
            Tw.append(RjT @ Tw[p] + a[:, None] * e[None, :])
            Tv.append(Tv[p] - R[p] @ skew(pj) @ Tw[p])
            if v is not None:
                rate = v[..., 5 + i]
                w_p = om[p]
                w_in_child = (RjT @ w_p[..., None])[..., 0]
                om.append(w_in_child + a * rate[..., None])
                al.append((RjT @ al[p][..., None])[..., 0]
                          - rate[..., None] * torch.cross(a.expand_as(w_in_child), w_in_child, dim=-1))
                wxwxp = torch.cross(w_p, torch.cross(w_p, pj.expand_as(w_p), dim=-1), dim=-1)
                be.append(be[p] + (R[p] @ (wxwxp - torch.cross(pj.expand_as(w_p), al[p], dim=-1))[..., None])[..., 0])
        return R, o, Tw, Tv, om, al, be

    @staticmethod
    def _body_mass_matrix(R: Tensor, inertia: Tensor) -> Tensor:
        m = inertia[..., 0][..., None, None]
        c = inertia[..., 1:4]
        Ixx, Iyy, Izz, Ixy, Ixz, Iyz = inertia[..., 4:].unbind(-1)
        I_sym = torch.stack((torch.stack((Ixx, Ixy, Ixz), -1), torch.stack((Ixy, Iyy, Iyz), -1),
                             torch.stack((Ixz, Iyz, Izz), -1)), -2)
        Sc = skew(c)
        I_o = I_sym - m * (Sc @ Sc)
        RT = R.transpose(-1, -2)
        eye = torch.eye(3, dtype=R.dtype).expand_as(R)
        top = torch.cat((I_o, m * (Sc @ RT)), -1)
        bot = torch.cat((-m * (R @ Sc), m * eye), -1)
        return torch.cat((top, bot), -2), I_o

    def mass_matrix(self, q: Tensor, inertia: Tensor) -> Tensor:
        R, _, Tw, Tv, _, _, _ = self.kinematics(q)
        M = 0.
        for i in range(self.tree.n_bodies):
            Mi, _ = self._body_mass_matrix(R[i], inertia[..., i, :])
            T = torch.cat((Tw[i], Tv[i]), -2)
            M = M + T.transpose(-1, -2) @ Mi @ T
        return M

    def lagrangian_forces(self, q: Tensor, v: Tensor, u: Tensor, inertia: Tensor) -> Tensor:
        del u  # both assets are unactuated (u has width 0)
        R, _, Tw, Tv, om, al, be = self.kinematics(q, v)
        g = torch.tensor([0., 0., -GRAVITY], dtype=q.dtype)
        F = 0.
        for i in range(self.tree.n_bodies):
            ine = inertia[..., i, :]
            Mi, I_o = self._body_mass_matrix(R[i], ine)
            m = ine[..., 0:1]
            c = ine[..., 1:4]
            w = om[i]
            RT = R[i].transpose(-1, -2)
            g_body = (RT @ g)
            tau = -torch.cross(w, (I_o @ w[..., None])[..., 0], dim=-1) + m * torch.cross(c, g_body.expand_as(c), dim=-1)
            wxwxc = torch.cross(w, torch.cross(w, c, dim=-1), dim=-1)
            frc = -m * (R[i] @ wxwxc[..., None])[..., 0] + m * g
            Fi = torch.cat((tau, frc), -1) - (Mi @ torch.cat((al[i], be[i]), -1)[..., None])[..., 0]
            T = torch.cat((Tw[i], Tv[i]), -2)
            F = F + (T.transpose(-1, -2) @ Fi[..., None])[..., 0]
        return F

    # -- geometry kinematics (world geometry = identity frame, zero Jacobian) --
    def geometry_rotations(self, q: Tensor) -> Tensor:
        R, _, _, _, _, _, _ = self.kinematics(q)
        eye = torch.eye(3, dtype=q.dtype).expand(q.shape[:-1] + (3, 3))
        return torch.stack([eye if b < 0 else R[b] @ self.tree.geometry_rotation(g, q.dtype)
                            for g, b in enumerate(self.tree.geometry_body)], -3)

    def geometry_translations(self, q: Tensor) -> Tensor:
        R, o, _, _, _, _, _ = self.kinematics(q)
        out = []
        for b, off in zip(self.tree.geometry_body, self.tree.geometry_offset):
            if b < 0:
                out.append(torch.zeros(q.shape[:-1] + (3,), dtype=q.dtype))
            else:
                out.append(o[b] + (R[b] @ torch.tensor(off, dtype=q.dtype)))
        return torch.stack(out, -2)

    def geometry_spatial_jacobians(self, q: Tensor) -> Tensor:
        """(*, n_g, 6, n_v): [omega_WG_W ; v_WGo_W] w.r.t. the state velocity."""
        R, _, Tw, Tv, _, _, _ = self.kinematics(q)
        nv = self.tree.n_v
        out = []
        for b, off in zip(self.tree.geometry_body, self.tree.geometry_offset):
            if b < 0:
                out.append(torch.zeros(q.shape[:-1] + (6, nv), dtype=q.dtype))
            else:
                p = torch.tensor(off, dtype=q.dtype)
                Jw = R[b] @ Tw[b]
                Jv = Tv[b] - R[b] @ skew(p) @ Tw[b]
                out.append(torch.cat((Jw, Jv), -2))
        return torch.stack(out, -3)
