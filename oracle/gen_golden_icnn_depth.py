"""ORACLE (test infrastructure).  Writes tests/golden/icnn_depth.npz: the REFERENCE's own ``HomogeneousICNN``
(dair_pll/deep_support_function.py:125-266, imported through oracle/ref_shim.py) at depths 1, 3 and 4 -- support points of
random directions, the gradients of a fixed linear functional of them with respect to every weight, and the summary mesh
(``extract_mesh``, :95-122) of the depth-3 network.  Needs /root/reference: build container only; the fixture is committed.

    python -m oracle.gen_golden_icnn_depth
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

from oracle import ref_shim  # noqa: E402

WIDTH = 24
DEPTHS = (1, 3, 4)


def main():
    ref_shim.import_reference()
    from dair_pll.deep_support_function import HomogeneousICNN, extract_mesh
    out = {'depths': np.array(DEPTHS), 'width': np.array(WIDTH)}
    g = torch.Generator().manual_seed(5)
    d = torch.randn(96, 3, generator=g, dtype=torch.float64)
    d = d / d.norm(dim=-1, keepdim=True)
    c = torch.randn(96, 3, generator=g, dtype=torch.float64)
    out['directions'], out['cotangent'] = d.numpy(), c.numpy()
    for depth in DEPTHS:
        torch.manual_seed(100 + depth)
        net = HomogeneousICNN(depth, WIDTH, negative_slope=0.5, scale=0.1).double()
        p = net(d)
        (p * c).sum().backward()
        out[f'd{depth}_p'] = p.detach().numpy()
        for i, w in enumerate(net.input_weights):
            out[f'd{depth}_in{i}'] = w.detach().numpy()
            out[f'd{depth}_gin{i}'] = w.grad.numpy()
        for i, w in enumerate(net.hidden_weights):
            out[f'd{depth}_hid{i}'] = w.detach().numpy()
            out[f'd{depth}_ghid{i}'] = w.grad.numpy()
        out[f'd{depth}_out'] = net.output_weight.detach().numpy()
        out[f'd{depth}_gout'] = net.output_weight.grad.numpy()
        if depth == 3:
            # the module-level direction set _SURFACE is single precision; the network runs in double as everywhere else
            mesh = extract_mesh(lambda s: net(s.double()))
            out['d3_mesh_vertices'] = mesh.vertices.numpy()
            out['d3_mesh_faces'] = mesh.faces.numpy()
    path = os.path.join(ROOT, 'tests', 'golden', 'icnn_depth.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: v.shape for k, v in out.items() if k.startswith('d3_')})


if __name__ == '__main__':
    main()
