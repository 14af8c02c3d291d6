"""dair_pll_b200: B200-native (sm_100a) ContactNets hot path behind dair_pll's Python API.

Importing the package does not load CUDA code; the first kernel call loads
``libdair_pll_b200.so`` (built in-tree by ``dair_pll_b200.build``) and raises if it is
missing -- there is no CPU or PyTorch fallback.
"""
__version__ = '0.1.0'
