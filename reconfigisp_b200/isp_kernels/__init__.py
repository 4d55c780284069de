"""The five kernel plugins the reference imports by bare module name (tools_origin.py:8-17).

Put THIS directory on `sys.path` (the reference appends `/DATA/ISP_Kernels/`; symlinking that path
here also works) and `import whitebalance, gamma, demosaic, globaltonemapping,
spatialnoisereduction` resolve to the B200 kernels.  `install()` does it for the current process."""
import os
import sys


def install():
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(os.path.dirname(here))
    for p in (root, here):
        if p not in sys.path:
            sys.path.insert(0, p)
    return here
