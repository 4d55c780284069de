"""reconfigisp_b200 -- B200-native (sm_100a) implementation of the ReconfigISP hot path:
the differentiable ISP module stack and the DARTS mixed-op that drives its pipeline search.

Layout:
  csrc/          hand-written CUDA kernels + the C ABI (include/reconfigisp_b200.h)
  _lib.py        ctypes binding (prototypes parsed from the header)
  ops.py         torch.autograd wrappers over the C ABI
  isp_kernels/   drop-in replacements for the five un-shipped plugin modules the reference imports
                 (whitebalance, gamma, demosaic, globaltonemapping, spatialnoisereduction)
  modules/       the reference's module / container classes (same names, signatures, state-dict keys)
"""
__version__ = '0.1.0'
