"""Drop-in for the un-shipped `/DATA/ISP_Kernels/globaltonemapping.py` (tools_origin.py:16).

    GlobalToneMapping().run(img_NHWC in [0,255], option, params) -> NHWC in [0,255]  (forward only)
      'reinhard'      params: white_point, middle_grey  np.ndarray (N,)   (tools_origin.py:535-546)
      'crysisengine'  params: lum_adapted                                  (:574-584)
      'filmic'        params: white_point, exposure_bias in [1,10]         (:615-626)
The parameters stay on the device once uploaded; no device->host round trip is needed here.
"""
from reconfigisp_b200 import ops
from reconfigisp_b200.isp_kernels._common import nhwc_to_nchw, nchw_to_nhwc, dev_vec


class GlobalToneMapping:
    def run(self, img, option, params):
        x = nhwc_to_nchw(img)
        if option == 'reinhard':
            y = ops.tone_reinhard(x, dev_vec(params['white_point'], x), dev_vec(params['middle_grey'], x), 255.0)
        elif option == 'crysisengine':
            y = ops.tone_crysis(x, dev_vec(params['lum_adapted'], x), 255.0)
        elif option == 'filmic':
            y = ops.tone_filmic(x, dev_vec(params['white_point'], x), dev_vec(params['exposure_bias'], x), 255.0)
        else:
            raise ValueError('globaltonemapping: unknown option %r' % (option,))
        return nchw_to_nhwc(y)
