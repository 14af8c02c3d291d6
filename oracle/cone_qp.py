"""ORACLE (test infrastructure, not product code).

Python face of ``oracle/cone_qp.c``: a stand-in for the un-vendored
``sappy.SAPSolver`` (``dair_pll/multibody_learnable_system.py:29,77``), with the
calling convention of its two call sites (``:181-184`` and ``:293-298``):
``SAPSolver.apply(J_M (*,k,n_v), q (*,k), eps) -> f (*,k)`` in sappy ordering
``[t_x, t_y, n]`` per contact (``tensor_utils.py:460-497``).

PARITY UNPINNED (no reference tests or vectors exist for the solver).  The
optimum is unique; :func:`kkt_residual` certifies it independently of the
algorithm, and :func:`solve_apg` is an unrelated second algorithm.

The backward pass (needed by ``forward_dynamics``, where the reference does not
detach the solver output) is implicit differentiation of the optimality
condition w = A^T Pi(-(A w + q)/eps): one *differentiable* Newton correction
around the converged, detached w* reproduces the implicit-function-theorem
derivative exactly, so plain autograd does the rest.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch
from torch import Tensor

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_build', 'libconeqp.so')
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/cone_qp.c (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, 'cone_qp.c')
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src))
    if force or stale:
        subprocess.check_call(['make', '-C', _HERE, '-s'] + (['-B'] if force else []))
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        lib = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        lib.cone_qp_solve_batch.argtypes = [dp, dp, ctypes.c_double, ctypes.c_int64, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                            dp, ctypes.c_int, dp, dp, ip, dp]
        lib.cone_qp_solve_batch.restype = ctypes.c_int
        lib.cone_qp_apg_batch.argtypes = [dp, dp, ctypes.c_double, ctypes.c_int64, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_double, dp, ip]
        lib.cone_qp_apg_batch.restype = ctypes.c_int
        _lib = lib
    return _lib


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def solve(A: np.ndarray, q: np.ndarray, eps: float, max_iter: int = 200, tol: float = 1e-15,
          nthreads: int = 0, ls_tol: float = 1e-14, w0: np.ndarray = None):
    """Batched Newton solve. A (B,k,nv), q (B,k) -> f (B,k), w (B,nv), iters (B), resid (B)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    B, k, nv = A.shape
    f = np.zeros((B, k))
    w = np.zeros((B, nv))
    iters = np.zeros(B, dtype=np.int32)
    resid = np.zeros(B)
    if w0 is not None:
        w0 = np.ascontiguousarray(w0, dtype=np.float64)
    rc = _load().cone_qp_solve_batch(_dptr(A), _dptr(q), eps, B, k // 3, nv, max_iter, tol, ls_tol,
                                     _dptr(w0) if w0 is not None else None, nthreads, _dptr(f), _dptr(w), iters.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                     _dptr(resid))
    if rc != 0:
        raise RuntimeError(f'cone_qp_solve_batch failed rc={rc}')
    return f, w, iters, resid


def solve_apg(A: np.ndarray, q: np.ndarray, eps: float, max_iter: int = 2000000, tol: float = 1e-15):
    """Independent dual accelerated-projected-gradient solve (slow; small sets only)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    B, k, nv = A.shape
    f = np.zeros((B, k))
    iters = np.zeros(B, dtype=np.int32)
    rc = _load().cone_qp_apg_batch(_dptr(A), _dptr(q), eps, B, k // 3, nv, max_iter, tol, _dptr(f),
                                   iters.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    if rc != 0:
        raise RuntimeError(f'cone_qp_apg_batch failed rc={rc}')
    return f, iters


def project_lorentz_sappy(y: Tensor) -> Tensor:
    """Differentiable projection onto prod L3 in sappy ordering (*, 3 n_c)."""
    shape = y.shape
    y3 = y.reshape(shape[:-1] + (-1, 3))
    t, n = y3[..., :2], y3[..., 2:]
    r = t.norm(dim=-1, keepdim=True)
    inside = r <= n
    polar = (~inside) & (r <= -n)
    s = 0.5 * (n + r)
    safe_r = torch.where(r > 0, r, torch.ones_like(r))
    proj = torch.cat((s * t / safe_r, s), -1)
    out = torch.where(inside, y3, torch.where(polar, torch.zeros_like(y3), proj))
    return out.reshape(shape)


def projection_jacobian(y: Tensor) -> Tensor:
    """(*, n_c, 3, 3) Jacobian of the projection at y (*, 3 n_c); constant w.r.t. autograd."""
    y3 = y.detach().reshape(y.shape[:-1] + (-1, 3))
    t, n = y3[..., :2], y3[..., 2]
    r = t.norm(dim=-1)
    inside = r <= n
    polar = (~inside) & (r <= -n)
    safe_r = torch.where(r > 0, r, torch.ones_like(r))
    that = t / safe_r[..., None]
    s = 0.5 * (n + r)
    a = (s / safe_r)[..., None, None]
    eye2 = torch.eye(2, dtype=y.dtype)
    tt = that[..., :, None] * that[..., None, :]
    G = torch.zeros(y3.shape[:-1] + (3, 3), dtype=y.dtype)
    G[..., :2, :2] = a * (eye2 - tt) + 0.5 * tt
    G[..., :2, 2] = 0.5 * that
    G[..., 2, :2] = 0.5 * that
    G[..., 2, 2] = 0.5
    eye3 = torch.eye(3, dtype=y.dtype).expand_as(G)
    G = torch.where(inside[..., None, None], eye3, G)
    G = torch.where(polar[..., None, None], torch.zeros_like(G), G)
    return G


def kkt_residual(A: Tensor, q: Tensor, eps: float, f: Tensor) -> Tensor:
    """Scaled KKT violation of f for min 1/2 f^T(AA^T+eps I)f + q^T f over prod L3.

    max( dist(f, K), dist(s, K), |f.s| ) / max(1, |f|, |q|)  with s = Qf + q; K self-dual."""
    s = (A @ (A.transpose(-1, -2) @ f[..., None]))[..., 0] + eps * f + q
    df = (f - project_lorentz_sappy(f)).norm(dim=-1)
    ds = (s - project_lorentz_sappy(s)).norm(dim=-1)
    fn, sn = f.norm(dim=-1), s.norm(dim=-1)
    comp = (f * s).sum(-1).abs()
    scale = torch.maximum(torch.ones_like(fn), torch.maximum(fn, q.norm(dim=-1)))
    return torch.maximum(torch.maximum(df, ds) / scale, comp / (scale * scale))


class OracleSAPSolver:
    """Drop-in for ``sappy.SAPSolver()``: ``solver.apply(J, q, eps)``."""

    def __init__(self, nthreads: int = 0):
        self.nthreads = nthreads
        self.last_iters = None

    def apply(self, J: Tensor, q: Tensor, eps: float, return_primal: bool = False) -> Tensor:
        """f (..., k), as sappy.  ``return_primal``: return the primal optimum w = J^T f (..., n_v) instead -- the
        same differentiable Newton correction, for callers that evaluate M^-1 J^T f as L^-T w (the better
        conditioned form of the same quantity: f = Pi(-(J w + q)/eps) amplifies rounding in w by 1/eps)."""
        batch = q.shape[:-1]
        k, nv = J.shape[-2], J.shape[-1]
        A = J.reshape((-1, k, nv))
        qq = q.reshape((-1, k))
        _, w_np, iters, _ = solve(A.detach().numpy(), qq.detach().numpy(), eps, nthreads=self.nthreads)
        self.last_iters = iters
        w0 = torch.from_numpy(w_np).to(J.dtype)
        if A.requires_grad or qq.requires_grad:
            # differentiable Newton correction: value unchanged (gradient is ~0 at w0),
            # derivative = implicit function theorem.
            y0 = -((A @ w0[..., None])[..., 0] + qq) / eps
            G = projection_jacobian(y0).detach()
            Ad = A.detach()
            A3 = Ad.reshape(Ad.shape[0], k // 3, 3, nv)
            H = torch.eye(nv, dtype=J.dtype) + (A3.transpose(-1, -2) @ G @ A3).sum(1) / eps
            grad = w0 - (A.transpose(-1, -2) @ project_lorentz_sappy(y0)[..., None])[..., 0]
            w = w0 - torch.linalg.solve(H, grad[..., None])[..., 0]
        else:
            w = w0
        if return_primal:
            return w.reshape(batch + (nv,))
        y = -((A @ w[..., None])[..., 0] + qq) / eps
        f = project_lorentz_sappy(y)
        return f.reshape(batch + (k,))
