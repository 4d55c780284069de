"""Drop-in for the un-shipped `/DATA/ISP_Kernels/spatialnoisereduction.py` (tools_origin.py:17).

    SpatialNoiseReduction().run(img_NHWC in [0,255], option, params) -> NHWC in [0,255]  (forward only)
      'bilateral'  window_length IntTensor (N,), sigma_color / sigma_space Tensor (N,)  (tools_origin.py:696-710)
      'median'     size int                                                              (:742-751)
      'fastnlm'    block_size / search_block IntTensor (N,), decay_factor Tensor (N,)        (:785-797; oracle/SPEC.md)
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
            as_int = lambda v: (v if torch.is_tensor(v) else torch.as_tensor(v)).reshape(-1)
            y = ops.fastnlm(x, as_int(params['block_size']), as_int(params['search_block']), dev_vec(params['decay_factor'], x))
        else:
            raise ValueError('spatialnoisereduction: unknown option %r' % (option,))
        return nchw_to_nhwc(y)
