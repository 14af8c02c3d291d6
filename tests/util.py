"""Shared helpers for the test-suite."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def rel_err(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def max_rel_to_scale(a, b):
    """max |a-b| / max|b| -- for vectors whose small entries are not individually meaningful."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def oracle_params_from_golden(g, requires_grad=True):
    from oracle import contactnets_oracle as co
    halves = np.atleast_2d(g['half_lengths'])
    p = co.OracleParams(torch.from_numpy(g['theta']).clone(), torch.from_numpy(g['friction_params']).clone(),
                        [torch.from_numpy(h.copy()).reshape(1, 3) for h in halves])
    return p.requires_grad_(requires_grad)


def kernel_level_params(g):
    """inertia (10), mu_pair (1), half (3) numpy arrays from a golden file's learnables."""
    from oracle import contactnets_oracle as co
    inertia = co.theta_to_inertia_vector(torch.from_numpy(g['theta'])).reshape(10).numpy()
    mu = np.abs(g['friction_params'])
    mu_pair = np.array([2 * mu[0] * mu[1] / (mu[0] + mu[1])])
    return inertia, mu_pair, np.abs(g['half_lengths']).reshape(3)


_EMUL = None


def host_emulation_lib():
    """g++ build of the device per-sample math (tests/host_emul/emul.cpp) -- debugging aid so the
    kernel arithmetic can be checked in a container without a GPU.  Never used by the package."""
    global _EMUL
    if _EMUL is None:
        out = os.path.join(tempfile.gettempdir(), f'dpll_emul_{os.getuid()}.so')
        src = os.path.join(ROOT, 'tests', 'host_emul', 'emul.cpp')
        gxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
        subprocess.check_call([gxx, '-O2', '-std=c++17', '-fPIC', '-shared', '-ffp-contract=off',
                               '-o', out, src])
        _EMUL = ctypes.CDLL(out)
    return _EMUL


def dptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)
