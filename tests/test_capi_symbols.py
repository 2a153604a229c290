"""CPU: the C-ABI shared library loads and exports every symbol include/afft_b200.h declares.
No compute calls here (there is no GPU in the build container); error paths that do not touch a device are
exercised."""
import ctypes
import os
import re
import subprocess

import pytest

from afft_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "afft_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(afft_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    declared = _declared_symbols()
    assert len(declared) >= 15
    lib = _capi.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/afft_b200.h but not exported by {_capi.LIB_PATH}"
    assert sorted(_capi.EXPORTED_SYMBOLS) == declared


def test_exports_are_unmangled_c():
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for name in _declared_symbols():
        assert name in exported


def test_abi_version_and_struct_sizes():
    lib = _capi.lib()
    assert lib.afft_abi_version() == _capi.ABI_VERSION
    # struct layouts mirrored in _capi.py must match the C header (sizes computed from the C compiler)
    code = r'''
    #include "afft_b200.h"
    #include <stdio.h>
    int main(void) { printf("%zu %zu %zu %zu %zu\n", sizeof(afft_gemm_desc), sizeof(afft_layernorm_desc),
                            sizeof(afft_attention_desc), sizeof(afft_config), sizeof(afft_io)); return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(code)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_capi.GemmDesc), ctypes.sizeof(_capi.LayerNormDesc), ctypes.sizeof(_capi.AttentionDesc),
                     ctypes.sizeof(_capi.Config), ctypes.sizeof(_capi.IO)]


def test_argument_errors_do_not_need_a_device():
    lib = _capi.lib()
    assert lib.afft_gemm(None, None) != 0
    assert b"null" in lib.afft_last_error()
    assert lib.afft_layernorm(None, None) != 0
    assert lib.afft_attention(None, None) != 0
    h = ctypes.c_void_p()
    cfg = _capi.Config()
    cfg.fuser_kind = 99
    assert lib.afft_create(ctypes.byref(cfg), ctypes.byref(h)) == 1  # AFFT_ERR_INVALID
    assert not h.value
    assert lib.afft_forward(None, 1, None, None) != 0
    lib.afft_destroy(None)  # no-op


def test_cpu_tensors_are_rejected():
    import torch
    with pytest.raises(_capi.AfftError):
        _capi.gemm(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(128, 64, dtype=torch.bfloat16),
                   out_f32=torch.zeros(128, 128))
