"""Drop-in for the un-shipped `/DATA/ISP_Kernels/spatialnoisereduction.py` (tools_origin.py:17).

    SpatialNoiseReduction().run(img_NHWC in [0,255], option, params) -> NHWC in [0,255]  (forward only)
      'bilateral'  window_length IntTensor (N,), sigma_color / sigma_space Tensor (N,)  (tools_origin.py:696-710)
      'median'     size int                                                              (:742-751)
      'fastnlm'    not restated (only its SRCNNRes proxy is on the search path)
"""
import torch

from reconfigisp_b200 import ops
from reconfigisp_b200.isp_kernels._common import nhwc_to_nchw, nchw_to_nhwc, dev_vec


class SpatialNoiseReduction:
    def run(self, img, option, params):
        x = nhwc_to_nchw(img)
        if option == 'bilateral':
            win = params['window_length']
            win = win if torch.is_tensor(win) else torch.as_tensor(win)
            y = ops.bilateral(x, win.reshape(-1), dev_vec(params['sigma_color'], x), dev_vec(params['sigma_space'], x))
        elif option == 'median':
            y = ops.median(x, int(params['size']))
        elif option == 'fastnlm':
            raise NotImplementedError('spatialnoisereduction: fastnlm is outside the rebuilt hot path')
        else:
            raise ValueError('spatialnoisereduction: unknown option %r' % (option,))
        return nchw_to_nhwc(y)
