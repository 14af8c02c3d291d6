"""Inertial parameterisations (host side, per body -- n_bodies x 10, never per sample).

Mirror of ``dair_pll/inertia.py``'s ``InertialParameterConverter`` for the formats on the
hot path (format definitions: inertia.py:17-79):

* ``pi_cm``  [m, m c, Ixx, Iyy, Izz, Ixy, Ixz, Iyz] about the centre of mass,
* ``pi_o``   same about the body origin,
* ``theta``  the unconstrained log-Cholesky coordinates that are the learnable
  ``inertial_parameters`` (theta_to_pi_o :205-234, pi_o_to_theta :236-302),
* ``drake_spatial_inertia`` [m, c, I_cm / m] -- the vector the generated callables (and the
  kernels) receive (:376-382).

Unlike the reference, importing this module does not change torch's default dtype
(inertia.py:96 does); dtype and device follow the inputs.
"""
from typing import Dict, List, Tuple

import torch
from torch import Tensor

INERTIA_SCALARS = ['I_xx', 'I_yy', 'I_zz', 'I_xy', 'I_xz', 'I_yz']
AXES = ['x', 'y', 'z']


def _sym_from_vec(vec: Tensor) -> Tensor:
    xx, yy, zz, xy, xz, yz = vec.unbind(-1)
    return torch.stack((torch.stack((xx, xy, xz), -1), torch.stack((xy, yy, yz), -1),
                        torch.stack((xz, yz, zz), -1)), -2)


def _vec_from_sym(mat: Tensor) -> Tensor:
    return torch.stack((mat[..., 0, 0], mat[..., 1, 1], mat[..., 2, 2],
                        mat[..., 0, 1], mat[..., 0, 2], mat[..., 1, 2]), -1)


def _point_mass_inertia(m: Tensor, c: Tensor) -> Tensor:
    """m (|c|^2 I - c c^T): inertia about a point of a point mass displaced by c."""
    eye = torch.eye(3, dtype=c.dtype, device=c.device)
    return m[..., None, None] * ((c * c).sum(-1)[..., None, None] * eye - c[..., :, None] * c[..., None, :])


class InertialParameterConverter:
    """Conversions between inertial parameter formats (all batched over leading dims)."""

    @staticmethod
    def theta_to_pi_o(theta: Tensor) -> Tensor:
        alpha, d1, d2, d3, s12, s23, s13, t1, t2, t3 = theta.unbind(-1)
        e1, e2, e3 = d1.exp(), d2.exp(), d3.exp()
        rows = (t1 * t1 + t2 * t2 + t3 * t3 + 1,
                t1 * e1,
                t1 * s12 + t2 * e2,
                t1 * s13 + t2 * s23 + t3 * e3,
                s12 * s12 + s23 * s23 + s13 * s13 + e2 * e2 + e3 * e3,
                s13 * s13 + s23 * s23 + e1 * e1 + e3 * e3,
                s12 * s12 + e1 * e1 + e2 * e2,
                -s12 * e1,
                -s13 * e1,
                -s12 * s13 - s23 * e2)
        return (2 * alpha).exp()[..., None] * torch.stack(rows, -1)

    @staticmethod
    def pi_o_to_theta(pi_o: Tensor) -> Tensor:
        m, hx, hy, hz, xx, yy, zz, xy, xz, yz = pi_o.unbind(-1)
        a_e1 = (0.5 * (yy + zz - xx)).sqrt()
        a_s12 = -xy / a_e1
        a_s13 = -xz / a_e1
        a_e2 = (zz - a_e1 ** 2 - a_s12 ** 2).sqrt()
        a_s23 = (-yz - a_s12 * a_s13) / a_e2
        a_e3 = (yy - a_e1 ** 2 - a_s13 ** 2 - a_s23 ** 2).sqrt()
        a_t1 = hx / a_e1
        a_t2 = (hy - a_t1 * a_s12) / a_e2
        a_t3 = (hz - a_t1 * a_s13 - a_t2 * a_s23) / a_e3
        ea = (m - a_t1 ** 2 - a_t2 ** 2 - a_t3 ** 2).sqrt()
        return torch.stack((ea.log(), (a_e1 / ea).log(), (a_e2 / ea).log(), (a_e3 / ea).log(),
                            a_s12 / ea, a_s23 / ea, a_s13 / ea, a_t1 / ea, a_t2 / ea, a_t3 / ea), -1)

    @staticmethod
    def pi_o_to_pi_cm(pi_o: Tensor) -> Tensor:
        pi_o = pi_o.reshape(-1, 10)
        m = pi_o[:, 0]
        c = pi_o[:, 1:4] / m[:, None]
        I_cm = _sym_from_vec(pi_o[:, 4:]) - _point_mass_inertia(m, c)
        return torch.cat((pi_o[:, :4], _vec_from_sym(I_cm)), -1)

    @staticmethod
    def pi_cm_to_pi_o(pi_cm: Tensor) -> Tensor:
        pi_cm = pi_cm.reshape(-1, 10)
        m = pi_cm[:, 0]
        c = pi_cm[:, 1:4] / m[:, None]
        I_o = _sym_from_vec(pi_cm[:, 4:]) + _point_mass_inertia(m, c)
        return torch.cat((pi_cm[:, :4], _vec_from_sym(I_o)), -1)

    @staticmethod
    def theta_to_pi_cm(theta: Tensor) -> Tensor:
        return InertialParameterConverter.pi_o_to_pi_cm(InertialParameterConverter.theta_to_pi_o(theta))

    @staticmethod
    def pi_cm_to_theta(pi_cm: Tensor) -> Tensor:
        return InertialParameterConverter.pi_o_to_theta(InertialParameterConverter.pi_cm_to_pi_o(pi_cm))

    @staticmethod
    def pi_cm_to_drake_spatial_inertia(pi_cm: Tensor) -> Tensor:
        return torch.cat((pi_cm[..., :1], pi_cm[..., 1:] / pi_cm[..., :1]), -1)

    @staticmethod
    def pi_cm_to_urdf(pi_cm: Tensor) -> Tuple[str, str, List[str]]:
        assert pi_cm.dim() == 1
        mass = str(pi_cm[0].item())
        com = ' '.join(str((v / pi_cm[0]).item()) for v in pi_cm[1:4])
        return mass, com, [str(v.item()) for v in pi_cm[4:]]

    @staticmethod
    def pi_cm_to_scalars(pi_cm: Tensor) -> Dict[str, float]:
        scalars = {'m': pi_cm[0].item()}
        scalars.update({f'com_{ax}': (pi_cm[1 + i] / pi_cm[0]).item() for i, ax in enumerate(AXES)})
        scalars.update({name: pi_cm[4 + i].item() for i, name in enumerate(INERTIA_SCALARS)})
        return scalars
