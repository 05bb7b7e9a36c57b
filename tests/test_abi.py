"""The C-ABI library loads on a CPU-only box and exports every symbol include/iiseg.h declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, 'include', 'iiseg.h')) as fh:
        text = fh.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(iiseg_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), 'libiiseg.so does not export %s' % n


def test_binding_table_matches_header(lib):
    from iterative_inference_segm_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_abi_version_and_error_text(lib):
    assert lib.iiseg_abi_version() == 5
    assert isinstance(lib.iiseg_last_error(), bytes)


def test_conv_desc_mirror_has_the_c_layout(lib):
    """The ctypes mirror of struct iiseg_conv_desc: same size, same offset of the last field."""
    import ctypes as C
    from iterative_inference_segm_b200 import _lib
    assert C.sizeof(_lib.ConvDesc) == lib.iiseg_conv_desc_size()
    assert _lib.ConvDesc.upd_cpad.offset == lib.iiseg_conv_desc_last_offset()
    assert _lib.ConvDesc._fields_[-1][0] == 'upd_cpad'


def test_bad_descriptor_is_rejected_without_gpu(lib):
    """Argument validation happens before any CUDA call, so it can be checked here."""
    import ctypes as C
    from iterative_inference_segm_b200 import _lib
    d = _lib.ConvDesc(weight=16, bias=16, out=16, N=1, H=8, W=8, Cout=64, R=3, S=3, pad=1,
                      oh0=0, ow0=0, OH=8, OW=8)
    d.src[0], d.C[0] = 16, 48
    assert lib.iiseg_conv2d_fwd(C.byref(d), None) != 0
    assert b'C0=48' in lib.iiseg_last_error()


def test_no_cpu_fallback_in_product_package():
    """The product package must not import the oracle or compute on the CPU."""
    pkg = os.path.join(ROOT, 'iterative_inference_segm_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert 'import oracle' not in src and 'from oracle' not in src, f
