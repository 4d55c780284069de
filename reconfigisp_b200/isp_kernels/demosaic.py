"""Drop-in for the un-shipped `/DATA/ISP_Kernels/demosaic.py` (tools_origin.py:15).

    Demosaic().run(img, option, params)
      'nearestneighbor' | 'demosaicnet'  img NCHW (N,1,H,W) in [0,1] -> NCHW (N,3,H,W) BGR, differentiable
                                          (tools_origin.py:278-284, :302-308)
      'bilinear' | 'laplacian'           img NHWC (N,H,W,1) in [0,255] -> NHWC (N,H,W,3) in [0,255]
                                          (:462-473, :496-507)
`demosaicnet` is an external network whose weights the reference does not ship; it is served by a
registered callable (default: an SRCNNDemosaic-architecture stand-in with fixed seeded weights, the
same stand-in the oracle uses -- see oracle/SPEC.md).
"""
import torch

from reconfigisp_b200 import ops
from reconfigisp_b200.isp_kernels._common import nhwc_to_nchw, nchw_to_nhwc

_DEMOSAICNET = {'fn': None}


def register_demosaicnet(fn):
    """fn(raw NCHW (N,1,H,W)) -> (N,3,H,W) BGR."""
    _DEMOSAICNET['fn'] = fn


def _default_demosaicnet(raw):
    from reconfigisp_b200.modules.tools_proxy import demosaicnet_standin
    return demosaicnet_standin(raw.device)(raw, None)


class Demosaic:
    def run(self, img, option, params):
        fmt = params.get('input', {}).get('format', 'RGGB') if isinstance(params, dict) else 'RGGB'
        if fmt != 'RGGB':
            raise ValueError('demosaic: only the RGGB phase is supported (got %r)' % (fmt,))
        if option == 'nearestneighbor':
            return ops.demosaic(img, 'nearest')
        if option == 'demosaicnet':
            return (_DEMOSAICNET['fn'] or _default_demosaicnet)(img)
        if option == 'bilinear':
            return nchw_to_nhwc(ops.demosaic(nhwc_to_nchw(img).detach(), 'bilinear'))
        if option == 'laplacian':
            return nchw_to_nhwc(ops.demosaic(nhwc_to_nchw(img).detach(), 'malvar', 255.0))
        raise ValueError('demosaic: unknown option %r' % (option,))
