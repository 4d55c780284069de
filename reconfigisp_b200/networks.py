"""`define_G(opt)` -- the YAML -> network factory of the reference (codes/models/networks.py:10-48), same
option keys: network_G.{which_model_G, architecture, individual_module_paths, n_step, n_modules,
prune_threshold, conditional_modules}.  Extension keys: `module_path` (the reference hard-codes
'/DATA/module/', :11) and `weight_seed` (seeded stand-in weights when the checkpoints are absent)."""
import logging

logger = logging.getLogger('base')


def define_G(opt):
    opt_net = opt['network_G']
    module_path = opt_net.get('module_path', '/DATA/module/')
    seed = opt_net.get('weight_seed', None)
    cond_kwargs = opt_net.get('conditional_modules', None) or {}
    which_model = opt_net['which_model_G']
    if which_model == 'SuperPruneFifteenDemosFourBayerTwo':
        from .modules.super_prune_fifteen_demos_four_bayer_two import SuperPruneFifteenDemosFourBayerTwo as Net
        return Net(n_step=opt_net['n_step'], threshold=opt_net['prune_threshold'], module_path=module_path, weight_seed=seed)
    if which_model == 'SuperPruneFifteenDemosFourBayerTwoFt':
        from .modules.super_prune_fifteen_demos_four_bayer_two_ft import SuperPruneFifteenDemosFourBayerTwoFt as Net
        return Net(n_step=opt_net['n_step'], threshold=opt_net['prune_threshold'], module_path=module_path, weight_seed=seed)
    if which_model == 'IspUniversal':
        from .modules.isp_universal import IspUniversal
        return IspUniversal(module_path=module_path, indiv_module_paths=opt_net['individual_module_paths'],
                            architecture=opt_net['architecture'], weight_seed=seed, **cond_kwargs)
    if which_model == 'OriginUniversal':
        from .modules.origin_universal import OriginUniversal
        return OriginUniversal(module_path=module_path, architecture=opt_net['architecture'], weight_seed=seed)
    raise NotImplementedError('Generator model [{:s}] not recognized'.format(which_model))


def create_model(opt):
    """codes/models/__init__.py:5-22 for the model types on the rebuilt path."""
    model = opt['model']
    if model == 'darts':
        from .search import DartsModel as M
    elif model == 'darts_ft':
        from .search_ft import DartsFtModel as M
    elif model == 'isp':
        from .tuning import IspModel as M
    else:
        raise NotImplementedError('Model [{:s}] not recognized.'.format(model))
    m = M(opt)
    logger.info('Model [{:s}] is created.'.format(model))
    return m
