"""CPU: the C-ABI shared library loads and exports exactly what include/tgan.h declares; argument validation
fails with an error string instead of crashing (no compute is possible without a GPU)."""
import ctypes
import os
import subprocess

from tgan import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 39
    lib = _lib.load()
    for name in protos:
        assert hasattr(lib, name), name
    out = subprocess.run(['nm', '-D', '--defined-only', _lib._SO], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if ' T ' in l and l.split()[-1].startswith('tgan_')}
    assert exported == set(protos), (exported ^ set(protos))
    assert lib.tgan_version() >= 100


def test_struct_layouts_match_the_header():
    lib = _lib.load()
    assert lib.tgan_sizeof_igemm_args() == ctypes.sizeof(_lib.TganIgemmArgs)
    assert lib.tgan_sizeof_wgrad_args() == ctypes.sizeof(_lib.TganWgradArgs)


def test_argument_validation_reports_errors():
    lib = _lib.load()
    assert lib.tgan_fill_f32(None, 0.0, 10, None) != 0
    assert b'fill' in lib.tgan_last_error()
    assert lib.tgan_sgemm(0, 0, 0, 4, 4, 1.0, None, 4, None, 4, 0.0, None, 4, 1, None, None) != 0
    a = _lib.TganIgemmArgs()
    assert lib.tgan_igemm_bf16(ctypes.byref(a), None) != 0 and b'igemm' in lib.tgan_last_error()
    assert lib.tgan_loss_c(None, None, 0, None, None, None, 0, None, None, 0, 10, None, None, None, None, None, None,
                           None) != 0


def test_only_sm100a_code_is_embedded():
    out = subprocess.run(['cuobjdump', '-lelf', _lib._SO], capture_output=True, text=True).stdout
    archs = {l.split('.')[-2] for l in out.splitlines() if 'sm_' in l}
    assert archs == {'sm_100a'}, out


def test_every_kernel_has_the_pdl_entry():
    """all launches carry the programmatic-stream-serialization attribute (common.cuh: pdl_launch), so a kernel that did
    not wait on its predecessor (pdl_entry / pdl_trigger + pdl_wait) would race with it"""
    import glob
    import re
    csrc = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tensorflow-implementation-of-triple-gan_b200', 'csrc')
    n = 0
    for f in sorted(glob.glob(os.path.join(csrc, '*.cu')) + glob.glob(os.path.join(csrc, '*.cuh'))):
        src = open(f).read()
        assert '<<<' not in src, f + ': raw launch (use pdl_launch)'
        for m in re.finditer(r'__global__', src):
            body = src[src.index('{', src.index(')', m.end())):][:400]
            assert 'pdl_entry();' in body or 'pdl_trigger();' in body, (f, src[m.start():m.start() + 120])
            n += 1
    assert n >= 45
