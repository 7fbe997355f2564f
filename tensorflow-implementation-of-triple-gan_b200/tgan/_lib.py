"""ctypes binding of libtgan.so (the sm_100a CUDA library behind the C ABI in include/tgan.h).

The prototypes are parsed from include/tgan.h so that the header is the single source of truth.
There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
_HEADER = os.path.normpath(os.path.join(_PKG, '..', '..', 'include', 'tgan.h'))
_SO = os.environ.get('TGAN_LIBTGAN') or os.path.join(_PKG, 'libtgan.so')      # (override: kernel-variant experiments)

_CT = {'int': ctypes.c_int, 'float': ctypes.c_float, 'int64_t': ctypes.c_int64, 'uint64_t': ctypes.c_uint64,
       'void': None}


def parse_header(path=_HEADER):
    """-> {name: (restype, [argtypes])} for every `tgan_*` prototype declared in the header."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    src = re.sub(r'^\s*#.*$', ' ', src, flags=re.M)
    src = re.sub(r'typedef\s+struct\s*\{.*?\}\s*\w+\s*;', ' ', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'([\w][\w\s\*]*?)\b(tgan_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if '*' in ret:
            restype = ctypes.c_char_p if 'char' in ret else ctypes.c_void_p
        else:
            restype = _CT[ret.replace('const', '').strip()]
        argtypes = []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                if '*' in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    t = a.replace('const', '').split()[0]
                    argtypes.append(_CT[t])
        protos[name] = (restype, argtypes)
    return protos


class TganIgemmArgs(ctypes.Structure):
    _fields_ = [('x', ctypes.c_void_p), ('N', ctypes.c_int), ('H', ctypes.c_int), ('W', ctypes.c_int),
                ('C', ctypes.c_int), ('ldx', ctypes.c_int), ('wp', ctypes.c_void_p), ('T', ctypes.c_int),
                ('Nout', ctypes.c_int), ('Kpad', ctypes.c_int), ('dy', ctypes.c_int * 25), ('dx', ctypes.c_int * 25),
                ('gh', ctypes.c_int), ('gw', ctypes.c_int), ('sy', ctypes.c_int), ('sx', ctypes.c_int),
                ('out', ctypes.c_void_p), ('odt', ctypes.c_int),
                ('OH', ctypes.c_int), ('OW', ctypes.c_int), ('ldo', ctypes.c_int), ('osy', ctypes.c_int),
                ('osx', ctypes.c_int), ('ooy', ctypes.c_int), ('oox', ctypes.c_int), ('vh', ctypes.c_int),
                ('vw', ctypes.c_int), ('bias', ctypes.c_void_p), ('colsum', ctypes.c_void_p), ('nseg', ctypes.c_int), ('seg_end', ctypes.c_int * 4),
                ('act', ctypes.c_int),
                ('alpha', ctypes.c_float), ('ncls', ctypes.c_int), ('cls_T', ctypes.c_int * 4),
                ('cls_ooy', ctypes.c_int * 4), ('cls_oox', ctypes.c_int * 4),
                ('bias_seg', ctypes.c_int), ('clsum', ctypes.c_void_p), ('cls_h', ctypes.c_int), ('cls_w', ctypes.c_int),
                ('mask_out', ctypes.c_void_p), ('mask_in', ctypes.c_void_p), ('mask_alpha', ctypes.c_float)]


class TganWgradArgs(ctypes.Structure):
    _fields_ = [('dz', ctypes.c_void_p), ('N', ctypes.c_int), ('gh', ctypes.c_int), ('gw', ctypes.c_int),
                ('Cout', ctypes.c_int), ('lddz', ctypes.c_int), ('x', ctypes.c_void_p), ('H', ctypes.c_int),
                ('W', ctypes.c_int), ('Cin', ctypes.c_int), ('ldx', ctypes.c_int), ('sy', ctypes.c_int), ('sx', ctypes.c_int),
                ('T', ctypes.c_int),
                ('dy', ctypes.c_int * 25), ('dx', ctypes.c_int * 25), ('dw', ctypes.c_void_p),
                ('beta', ctypes.c_float), ('ws', ctypes.c_void_p), ('ws_bytes', ctypes.c_int64),
                ('cin_store', ctypes.c_int)]


class TganWnDesc(ctypes.Structure):
    _fields_ = [('V', ctypes.c_void_p), ('g', ctypes.c_void_p), ('inv_norm', ctypes.c_void_p), ('scale', ctypes.c_void_p),
                ('dW', ctypes.c_void_p), ('dV', ctypes.c_void_p), ('dg', ctypes.c_void_p),
                ('A', ctypes.c_int), ('Co', ctypes.c_int), ('B', ctypes.c_int), ('eps_mode', ctypes.c_int)]


class TganPackDesc(ctypes.Structure):
    _fields_ = [('src', ctypes.c_void_p), ('dst', ctypes.c_void_p), ('scale', ctypes.c_void_p), ('taps', ctypes.c_void_p),
                ('st', ctypes.c_int64), ('sn', ctypes.c_int64), ('sk', ctypes.c_int64),
                ('T', ctypes.c_int), ('Nr', ctypes.c_int), ('K', ctypes.c_int), ('Kpad', ctypes.c_int),
                ('scale_on', ctypes.c_int), ('scale_mod', ctypes.c_int)]


_lib = None


def load():
    """Load libtgan.so or fail loudly (no CPU / eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise RuntimeError('libtgan.so not found at %s -- build it with `python -c "import __graft_entry__ as g; '
                           'g.build()"` (or `make -C tensorflow-implementation-of-triple-gan_b200/csrc`). '
                           'There is no fallback path.' % _SO)
    lib = ctypes.CDLL(_SO)
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(lib, name)      # AttributeError here == header/library mismatch: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    for cls, fn in ((TganIgemmArgs, lib.tgan_sizeof_igemm_args), (TganWgradArgs, lib.tgan_sizeof_wgrad_args),
                    (TganWnDesc, lib.tgan_sizeof_wn_desc), (TganPackDesc, lib.tgan_sizeof_pack_desc)):
        if ctypes.sizeof(cls) != fn():
            raise RuntimeError('libtgan: struct layout mismatch for %s (%d vs %d)' % (cls.__name__, ctypes.sizeof(cls), fn()))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError('libtgan: ' + load().tgan_last_error().decode())


def call(name, *args):
    check(getattr(load(), name)(*args))
