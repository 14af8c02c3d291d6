"""Driver of tools/exp_solver_trace.cpp: builds the bench batch's first 131,072 pairs on the CPU (host emulation of the device
step + the bench's noise), writes them to a flat file, compiles the tracer and prints the visit histogram and the traces of a
few solves with the requested visit count.  Usage: python tools/exp_solver_trace.py [visits=11] [how_many=3].  CPU only."""
import ctypes
import os
import struct
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dair_pll_b200 import synthetic  # noqa: E402
from dair_pll_b200.inertia import InertialParameterConverter as IPC  # noqa: E402
from tests.util import dptr, host_emulation_lib  # noqa: E402

B = 131072
lib = host_emulation_lib()
pi, fr, half = synthetic.cube_learnables_perturbed(0)
inertia = IPC.pi_cm_to_drake_spatial_inertia(pi.reshape(1, 10)).reshape(10).numpy().copy()
a, b = fr.abs().numpy()
mu = np.array([2 * a * b / (a + b)])
halfn = half.abs().numpy().reshape(3).copy()
x = synthetic.cube_states(B, seed=0).numpy().copy()
xn = np.empty_like(x)
lib.emul_cube_step_f64(dptr(x), dptr(inertia), dptr(mu), dptr(halfn), ctypes.c_double(0.0068), ctypes.c_double(1e-4),
                       ctypes.c_int64(B), dptr(xn), None, None)
xp = synthetic.perturb_next_state(torch.from_numpy(xn), seed=7919).numpy().copy()
path = os.path.join(tempfile.gettempdir(), 'dpll_exp_batch.bin')
with open(path, 'wb') as f:
    f.write(struct.pack('q', B))
    for arr in (x, xp, inertia, mu, halfn):
        f.write(np.ascontiguousarray(arr, dtype=np.float64).tobytes())
exe = os.path.join(tempfile.gettempdir(), 'dpll_exp_solver_trace')
subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-o', exe, os.path.join(ROOT, 'tools', 'exp_solver_trace.cpp')])
subprocess.check_call([exe, sys.argv[1] if len(sys.argv) > 1 else '11', sys.argv[2] if len(sys.argv) > 2 else '3', path])
