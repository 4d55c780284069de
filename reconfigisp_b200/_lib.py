"""ctypes binding of `libreconfigisp_b200.so` (the C ABI declared in include/reconfigisp_b200.h).

The prototypes are parsed from the header itself, so the binding cannot drift from the ABI.
There is NO fallback: if the library is missing, importing this module raises, and every op in
`reconfigisp_b200.ops` refuses non-CUDA tensors.
"""
import ctypes
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'reconfigisp_b200.h')
LIB_PATH = os.environ.get('RISP_LIB_PATH', os.path.join(HERE, 'libreconfigisp_b200.so'))

RISP_OK, RISP_E_INVALID, RISP_E_ALIGN, RISP_E_CUDA, RISP_E_UNSUPPORTED, RISP_E_WORKSPACE = 0, -1, -2, -3, -4, -5

_SCALARS = {'int': ctypes.c_int, 'float': ctypes.c_float, 'long long': ctypes.c_longlong,
            'size_t': ctypes.c_size_t, 'risp_stream_t': ctypes.c_void_p}


def parse_header(path=HEADER):
    """-> {name: (restype, [(ctype, argname)])} for every `risp_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(long long|int|size_t|const char\s*\*)\s+(risp_\w+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), ' '.join(m.group(3).split())
        restype = {'int': ctypes.c_int, 'size_t': ctypes.c_size_t, 'long long': ctypes.c_longlong}.get(ret, ctypes.c_char_p)
        argl = []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                if '*' in a:
                    argl.append((ctypes.c_void_p, a.split('*')[-1].strip()))
                else:
                    toks = a.split()
                    ty = ' '.join(t for t in toks[:-1] if t != 'const')
                    argl.append((_SCALARS[ty], toks[-1]))
        protos[name] = (restype, argl)
    return protos


def parse_enums(path=HEADER):
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    out = {}
    for body in re.findall(r'enum\s+\w+\s*\{(.*?)\}', src, flags=re.S):
        for item in body.split(','):
            if '=' in item:
                k, v = item.split('=')
                out[k.strip()] = int(v.strip())
    for k, v in re.findall(r'#define\s+(RISP_\w+)\s+(-?\d+)', src):
        out[k] = int(v)
    return out


PROTOS = parse_header()
ENUMS = parse_enums()
globals().update(ENUMS)


class RispError(RuntimeError):
    pass


_EXC = {RISP_E_INVALID: ValueError, RISP_E_ALIGN: ValueError, RISP_E_CUDA: RispError,
        RISP_E_UNSUPPORTED: NotImplementedError, RISP_E_WORKSPACE: RispError}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError('%s is missing: build it with `python -m reconfigisp_b200._build` '
                          '(there is no CPU / eager fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, args) in PROTOS.items():
        fn = getattr(lib, name)          # raises AttributeError if the .so lacks a declared symbol
        fn.restype = restype
        fn.argtypes = [t for t, _ in args]
    assert lib.risp_abi_version() == ENUMS['RISP_ABI_VERSION'], 'header / library ABI mismatch'
    return lib


lib = _load()


_FN = {}


def call(name, *args):
    """Invoke an `int risp_*` entry point, mapping error codes to the reference's exception types."""
    fn = _FN.get(name)
    if fn is None:
        fn = _FN[name] = getattr(lib, name)
    rc = fn(*args)
    if rc != RISP_OK:
        msg = lib.risp_last_error().decode()
        raise _EXC.get(rc, RispError)('%s failed (%d): %s' % (name, rc, msg))


def size(name, *args):
    return int(getattr(lib, name)(*args))


# The launch path is host-bound at small batches (a search iteration enqueues ~4.5 k kernels), so the per-argument checks use
# the raw C accessors: torch.cuda.current_device() / current_stream() go through several Python layers (~1.3 / ~15 us a call).
_cur_device = getattr(torch._C, '_cuda_getDevice', None) or torch.cuda.current_device
_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def ptr(t):
    """Device pointer of a CUDA fp32/int tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('reconfigisp_b200 ops run on CUDA tensors only (no CPU fallback); got a %s tensor' % t.device)
    if not t.is_contiguous():
        raise ValueError('reconfigisp_b200 ops need contiguous tensors')
    if t.get_device() != _cur_device():
        # kernels are enqueued on the CURRENT device's stream: a tensor of another GPU would be an illegal access
        raise RuntimeError('tensor lives on cuda:%d but the current device is cuda:%d; wrap the call in torch.cuda.device(t.device)'
                           % (t.get_device(), _cur_device()))
    return t.data_ptr()


def stream():
    if _raw_stream is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


def iarr(values):
    """Host int array argument."""
    values = list(values)
    return (ctypes.c_int * max(1, len(values)))(*values)


def parr(tensors):
    """Host array of device pointers."""
    return (ctypes.c_void_p * max(1, len(tensors)))(*[t.data_ptr() for t in tensors])


def parr_opt(tensors):
    """Host array of device pointers; None entries become NULL."""
    return (ctypes.c_void_p * max(1, len(tensors)))(*[None if t is None else t.data_ptr() for t in tensors])


def workspace(nbytes, device):
    return torch.empty(max(1, (nbytes + 3) // 4), dtype=torch.float32, device=device)
