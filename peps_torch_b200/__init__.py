"""peps_torch_b200 -- B200-native CTMRG move engine behind the peps-torch API.

Host code (Python) keeps the iPEPS / environment tensors as torch CUDA tensors and drives
hand-written sm_100a kernels through the C ABI of libctmb.so (include/ctmb.h).
The drop-in modules mirror the reference's plugin interface for the hot path:

    peps_torch_b200.ctm.generic.ctmrg         run, ctm_MOVE          (ctm/generic/ctmrg.py)
    peps_torch_b200.ctm.one_site_c4v.ctmrg_c4v  run, ctm_MOVE_sl     (ctm/one_site_c4v/ctmrg_c4v.py)

There is no CPU path: importing the engine without the built CUDA library raises.
"""
__version__ = '0.1.0'
