"""`SuperPruneFifteenDemosFourBayerTwoFt` -- the supernet variant used with online proxy fine-tuning
(codes/models/modules/super_prune_fifteen_demos_four_bayer_two_ft.py:13-273): identical search space and
forward, plus `proxy_ft_flag`, `n_step` and `load_proxy_nets`."""
from .super_prune_fifteen_demos_four_bayer_two import SuperPruneFifteenDemosFourBayerTwo


class SuperPruneFifteenDemosFourBayerTwoFt(SuperPruneFifteenDemosFourBayerTwo):
    def __init__(self, n_step, threshold, module_path, weight_seed=None):
        super().__init__(n_step, threshold, module_path, weight_seed)
        # (name, fine-tune flag) per sRGB candidate (:103-118); reinhard / filmic are off upstream
        # ("has nan bug"), bm3d has no differentiable original
        self.proxy_ft_flag = [('gamma', 0), ('reinhard', 0), ('crysisengine', 1), ('filmic', 0), ('grayworld', 0),
                              ('whiteworld', 1), ('bilateral', 1), ('median', 1), ('fastnlm', 1), ('skip', 0),
                              ('wbmanual', 0), ('path_restore_14l_bgr', 0), ('wbquadratic', 0), ('gtmmanual', 0),
                              ('bm3d', 0)]

    def load_proxy_nets(self, name_net_dict):
        """Copy the fine-tuned proxy weights into every sRGB step's instance of that proxy (:194-209)."""
        for idx, (name, ft_flag) in enumerate(self.proxy_ft_flag):
            if not ft_flag:
                continue
            state = name_net_dict[name].state_dict()
            for k in range(self.n_step):
                self.all_modules[-1 - k][idx].load_state_dict(state)
