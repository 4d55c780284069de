"""Drop-in for the un-shipped `/DATA/ISP_Kernels/gamma.py` (tools_origin.py:14).

    Gamma().run(img_NHWC in [0,1], 'manual', {'gamma': Tensor (N,1) in [0,1]}) -> NHWC, differentiable,
    one gamma per image (tools_origin.py:64-69, :181-194).
"""
from reconfigisp_b200 import ops
from reconfigisp_b200.isp_kernels._common import nhwc_to_nchw, nchw_to_nhwc


class Gamma:
    def run(self, img, option, params):
        if option != 'manual':
            raise ValueError('gamma: unknown option %r' % (option,))
        return nchw_to_nhwc(ops.gamma(nhwc_to_nchw(img), params['gamma']))
