"""CPU: the C-ABI library loads and exports every symbol include/mcb200.h declares; argument validation that runs
before any CUDA call reports through mc_last_error_string (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from modelcompression_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'mcb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mc_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "libmcb200.so does not export %s" % name
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes prototypes out of sync with include/mcb200.h"


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.mc_version() == 100
    assert isinstance(lib.mc_last_error_string(), bytes)


def test_argument_validation_without_gpu():
    lib = _lib.load()
    rc = lib.mc_conv_fwd(None, None)
    assert rc == -1 and b'null descriptor' in lib.mc_last_error_string()
    d = _lib.mc_conv_desc()
    d.d_in = d.d_wpack = d.d_scale = d.d_shift = d.d_out = 16
    d.ksize = 5
    rc = lib.mc_conv_fwd(ctypes.byref(d), None)
    assert rc == -1 and b'ksize' in lib.mc_last_error_string()
    rc = lib.mc_kth_abs_select(None, None, 0, 0, 0.0, None, None, 0, None)
    assert rc == -1
    with pytest.raises(_lib.McError):
        _lib.check(rc, "mc_kth_abs_select")
    assert lib.mc_workspace_bytes_kth_abs_select(1000) >= 4000


def test_conv_desc_layout_matches_header():
    # field order of the ctypes mirror == struct mc_conv_desc in the header
    text = open(os.path.join(ROOT, 'include', 'mcb200.h')).read()
    body = text[text.index('typedef struct mc_conv_desc {') + len('typedef struct mc_conv_desc {'):
                text.index('} mc_conv_desc;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in body.split(';'):
        for piece in decl.split(','):
            names = re.findall(r'[A-Za-z_][A-Za-z_0-9]*', piece)
            if names:
                fields.append(names[-1])
    assert fields == [f[0] for f in _lib.mc_conv_desc._fields_]


def test_cpu_tensor_fails_loudly(cfg_path):
    import torch
    import modelcompression_b200 as mc
    model = mc.Darknet(cfg_path).eval()
    with pytest.raises(RuntimeError, match="no CPU"):
        model(torch.zeros(1, 3, 416, 416))
    with pytest.raises(RuntimeError, match="no CPU"):
        mc.weight_prune(model, 50.)


def test_conv_desc_offsets_match_the_c_compiler(tmp_path):
    """The ctypes mirror of mc_conv_desc must have the C compiler's layout (size and every field offset): a field added
    on one side only would silently shift the descriptor the kernels read."""
    import subprocess
    fields = [f[0] for f in _lib.mc_conv_desc._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mcb200.h"', 'int main(void) {',
             '  printf("%zu\\n", sizeof(mc_conv_desc));']
    for f in fields:
        lines.append('  printf("%%zu\\n", offsetof(mc_conv_desc, %s));' % f)
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include')
    subprocess.run(['gcc', '-I', inc, str(src), '-o', str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_lib.mc_conv_desc)
    for f, off in zip(fields, out[1:]):
        assert getattr(_lib.mc_conv_desc, f).offset == off, f
