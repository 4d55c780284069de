"""Drop-in for the un-shipped `/DATA/ISP_Kernels/whitebalance.py` (tools_origin.py:13).

    WhiteBalance().run(img_NHWC, option, params) -> NHWC
      'grayworld'  img in [0,1], differentiable, clipped to [0,1]        (tools_origin.py:35-41)
      'manual'     params['gain'] Tensor (N,3), differentiable           (:216-221, :246-250)
      'whiteworld' img in [0,255], params['white_point_ratio'] (N,)      (:655-662)
Arithmetic: sm_100a kernels behind include/reconfigisp_b200.h; definitions in oracle/SPEC.md.
"""
from reconfigisp_b200 import ops
from reconfigisp_b200.isp_kernels._common import nhwc_to_nchw, nchw_to_nhwc, dev_vec


class WhiteBalance:
    def run(self, img, option, params):
        x = nhwc_to_nchw(img)
        if option == 'grayworld':
            y = ops.grayworld(x)
        elif option == 'manual':
            y = ops.gain(x, params['gain'])
        elif option == 'whiteworld':
            y = ops.whiteworld(x, dev_vec(params['white_point_ratio'], x), 255.0)
        else:
            raise ValueError('whitebalance: unknown option %r' % (option,))
        return nchw_to_nhwc(y)
