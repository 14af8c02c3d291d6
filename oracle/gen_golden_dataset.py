"""ORACLE (test infrastructure).  Writes tests/golden/dataset_slices.npz by running the REFERENCE's own
``TrajectorySliceDataset`` (dair_pll/dataset_management.py:17-67, imported through oracle/ref_shim.py) on the
first states of three recorded cube tosses (assets/contactnets_cube/{0,1,2}.pt) for several slice
configurations.  Needs /root/reference: build container only; the fixture is committed.

    python -m oracle.gen_golden_dataset
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

CONFIGS = [(0, 1, 1), (2, 2, 3), (1, 2, 1), (3, 4, 2)]      # (t_skip, t_history, t_prediction)
LENGTHS = [14, 9, 11]


def main():
    ref_shim.import_reference()
    from dair_pll.data_config import TrajectorySliceConfig
    from dair_pll.dataset_management import TrajectorySliceDataset
    out = {'configs': np.array(CONFIGS), 'n_traj': np.array(len(LENGTHS))}
    trajs = []
    for i, n in enumerate(LENGTHS):
        t = torch.load(os.path.join(ref_shim.REFERENCE_ROOT, 'assets', 'contactnets_cube', f'{i}.pt'))[:n].double()
        trajs.append(t)
        out[f'traj{i}'] = t.numpy()
    for c, (skip, hist, pred) in enumerate(CONFIGS):
        ds = TrajectorySliceDataset(TrajectorySliceConfig(t_skip=skip, t_history=hist, t_prediction=pred))
        for t in trajs:
            ds.add_slices_from_trajectory(t)
        out[f'previous{c}'] = torch.stack([ds[i][0] for i in range(len(ds))]).numpy()
        out[f'future{c}'] = torch.stack([ds[i][1] for i in range(len(ds))]).numpy()
    path = os.path.join(ROOT, 'tests', 'golden', 'dataset_slices.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
