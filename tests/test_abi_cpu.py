"""CPU-side checks of the C-ABI boundary and the host logic (no GPU, no compute calls)."""
import ctypes
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from reconfigisp_b200 import _build, _lib
    assert os.path.exists(_build.LIB), 'build the library first (python -m reconfigisp_b200._build)'
    lib = ctypes.CDLL(_build.LIB)
    assert len(_lib.PROTOS) >= 40
    for name in _lib.PROTOS:
        assert hasattr(lib, name), 'header declares %s but the library does not export it' % name
    lib.risp_abi_version.restype = ctypes.c_int
    assert lib.risp_abi_version() == _lib.ENUMS['RISP_ABI_VERSION']
    # argument validation happens before any CUDA call, so it can be exercised without a GPU
    lib.risp_chain_fwd.restype = ctypes.c_int
    lib.risp_last_error.restype = ctypes.c_char_p
    rc = lib.risp_chain_fwd(None, None, 0, ctypes.c_longlong(0), None, None, None, 0, None, 0, ctypes.c_float(1), ctypes.c_float(1), None)
    assert rc == _lib.RISP_E_INVALID and b'risp_chain_fwd' in lib.risp_last_error()


def test_no_cpu_fallback():
    from reconfigisp_b200 import ops
    with pytest.raises(RuntimeError):
        ops.gamma(torch.rand(1, 3, 8, 8), torch.rand(1, 1))
    with pytest.raises(RuntimeError):
        ops.demosaic(torch.rand(1, 1, 8, 8), 'bilinear')


def test_product_never_imports_the_oracle():
    import re
    for base, _, files in os.walk(os.path.join(ROOT, 'reconfigisp_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f


def test_architecture_grammar_and_registry():
    from reconfigisp_b200.modules import registry as R
    st = R.parse_architecture('Bayer_01_Demosaic_03_sRGB_01_13_11')
    assert [(s.step, s.domain, s.name) for s in st] == [(1, 'Bayer', 'path_bayer'), (2, 'Demosaic', 'laplacian'),
                                                      (3, 'sRGB', 'gamma'), (4, 'sRGB', 'wbquadratic'), (5, 'sRGB', 'wbmanual')]
    st = R.parse_architecture('Bayer_01_02_Demosaic_04_sRGB_01_13_11_05_02_03')       # several per domain (origin_universal.py:170)
    assert [s.name for s in st][:3] == ['path_bayer', 'skip', 'demosaicnet'] and st[-1].step == 9
    with pytest.raises(ValueError):
        R.parse_architecture('01_Bayer_02')
    with pytest.raises(AssertionError):
        R.parse_architecture('sRGB_22')
    with pytest.raises(AssertionError):
        R.parse_architecture('sRGB_16', R.N_SRGB_ORIGIN)
    assert len(R.NAMES['sRGB']) == 21 and R.NAMES['sRGB'][12] == 'wbquadratic'
    # default logits give identity-ish parameters (SURVEY §8c anchors)
    sig = lambda v: 1 / (1 + np.exp(-np.array(v)))
    assert np.allclose(sig(R.DEFAULT_LOGITS['wbmanual']) * 5, 1.0, atol=6e-3)
    assert np.allclose(sig(R.DEFAULT_LOGITS['gtmmanual']), [.25, .5, .75], atol=1e-3)


def test_containers_build_without_gpu_and_keep_reference_key_layout(golden):
    from reconfigisp_b200.modules.isp_universal import IspUniversal
    from reconfigisp_b200.modules.origin_universal import OriginUniversal
    from reconfigisp_b200.modules.super_prune_fifteen_demos_four_bayer_two_ft import SuperPruneFifteenDemosFourBayerTwoFt
    g = golden('fixed_pipelines')
    for tag, arch in (('classical', 'Bayer_02_Demosaic_02_sRGB_11_13_01'), ('sid', 'Bayer_01_Demosaic_03_sRGB_01_13_11'),
                      ('s7isp', 'Bayer_01_Demosaic_01_sRGB_04_01_13')):
        net = OriginUniversal('/x/', arch, weight_seed=10)
        assert list(net.state_dict().keys()) == list(g[tag + '_keys'])
    net = IspUniversal('/x/', (None,) * 7, 'Bayer_02_Demosaic_01_sRGB_11_13_01_14_05', weight_seed=10)
    assert list(net.state_dict().keys()) == list(g['isp_keys'])
    kinds = [s[0] for s in net._make_plan()]
    assert kinds == ['chain', 'head', 'module']                 # skip | nearest+4 per-pixel stages fused | grayworld
    sn = SuperPruneFifteenDemosFourBayerTwoFt(3, 0.2, '/x/', weight_seed=10)
    named = list(sn.named_parameters())
    assert len(named) == 41 and sum(p.numel() for _, p in named) == 216       # SURVEY §2a C2 (probed upstream)
    assert len(sn.trainable_parameters) == 45 and len(sn.alphas) == 5
    assert sum(f for _, f in sn.proxy_ft_flag) == 5
    cond = IspUniversal('/x/', (None,) * 3, 'Bayer_02_Demosaic_01_sRGB_16', weight_seed=1, gamma_in_channels=(12, 5))
    assert cond.is_conditional == [False, False, True] and cond.all_params[2].numel() == cond.all_modules[2].total_params


def test_patch_origins_and_synthetic_data():
    from reconfigisp_b200.ops import patch_origins
    from reconfigisp_b200.synthetic import synthetic_frames
    assert patch_origins(3000, 512, 480) == [0, 480, 960, 1440, 1920, 2400, 2488]
    assert len(patch_origins(4000, 512, 480)) == 9
    raw, gt = synthetic_frames(2, 32, 48, seed=10)
    assert raw.shape == (2, 1, 32, 48) and gt.shape == (2, 3, 32, 48)
    assert torch.equal(torch.round(raw * 1023) / 1023, raw) and torch.equal(torch.round(gt * 255) / 255, gt)
    raw2, _ = synthetic_frames(2, 32, 48, seed=10)
    assert torch.equal(raw, raw2) and not torch.equal(raw[0], raw[1])


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the contract's keys."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'MP/s' and line['value'] > 0 and line['higher_is_better'] is True
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'MP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in line['config'] and 'model' not in line['config']
    # ranks other than 0 exit without work under torchrun
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=120, cwd=root, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_header_is_plain_c_and_a_c_host_links():
    """The boundary is a C ABI: the header must compile as C99 and as C++17, and a C program must link and run
    against the shared library (no compute call: there is no GPU here)."""
    import os, shutil, subprocess, tempfile
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, 'reconfigisp_b200')
    src = ('#include "reconfigisp_b200.h"\n#include <stdio.h>\n'
           'int main(void) { size_t ws = risp_pipeline_step_workspace(4, 3000, 4000, 37);\n'
           '  printf("%d %d %zu\\n", risp_abi_version(), risp_conv_tc_supported(64, 64, 3), ws);\n'
           '  return (risp_chain_fwd(0, 0, 1, 16, 0, 0, 0, 0, 0, 0, 1.f, 1.f, 0) == RISP_OK) ? 1 : 0; }\n')
    with tempfile.TemporaryDirectory() as d:
        for ext, cc, std in (('c', 'gcc', '-std=c99'), ('cpp', 'g++', '-std=c++17')):
            path = os.path.join(d, 't.' + ext)
            open(path, 'w').write(src)
            exe = os.path.join(d, 't_' + ext)
            r = subprocess.run([cc, std, '-Wall', '-Werror', '-I', os.path.join(root, 'include'), path, '-L', libdir,
                                '-lreconfigisp_b200', '-Wl,-rpath,' + libdir, '-o', exe], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            r = subprocess.run([exe], capture_output=True, text=True)
            assert r.returncode == 0, (r.stdout, r.stderr)           # null pointers are rejected with an error code
            ver, tc, ws = r.stdout.split()
            assert int(ver) >= 1 and int(tc) == 1 and int(ws) > 0


def test_blocked_layout_group_count_matches_the_library():
    """ops.tc_groups restates risp_conv_tc_padded_channels (host-only entry) on the launch path."""
    from reconfigisp_b200 import _lib as L, ops
    for C in (1, 3, 4, 5, 12, 13, 17, 32, 48, 64):
        assert ops.tc_groups(C) * 4 == L.size('risp_conv_tc_padded_channels', C)
