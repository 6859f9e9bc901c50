"""The C-ABI library loads on a CPU-only box and exports every symbol include/mfm_b200.h declares."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mfm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mfm_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from mfm_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mfm_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound in _lib.py but not declared in the header"


def test_struct_layouts_match_header(lib):
    from mfm_b200 import _lib
    # mfm_target_t: 4 ints/floats..., pointers 8-byte aligned; sizes checked against a C compile
    import subprocess, tempfile, textwrap
    code = textwrap.dedent("""
        #include <stdio.h>
        #include "mfm_b200.h"
        int main(){ printf("%zu %zu %zu %zu %zu\\n", sizeof(mfm_target_t), sizeof(mfm_field_t), sizeof(mfm_ode_opts_t),
                           offsetof(mfm_target_t, kinv_diag), offsetof(mfm_field_t, omega)); return 0; }
    """)
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c"); exe = os.path.join(d, "t")
        open(c, "w").write(code)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    assert int(out[0]) == ctypes.sizeof(_lib.TargetDesc)
    assert int(out[1]) == ctypes.sizeof(_lib.FieldDesc)
    assert int(out[2]) == ctypes.sizeof(_lib.OdeOpts)
    assert int(out[3]) == _lib.TargetDesc.kinv_diag.offset
    assert int(out[4]) == _lib.FieldDesc.omega.offset


def test_host_split_matches_oracle(lib):
    from mfm_b200 import random as mr
    from oracle import threefry as tf
    for seed, num in [(0, 2), (1, 3), (59049, 6), (1024, 7)]:
        assert mr.host_split(tf.PRNGKey(seed), num).tolist() == tf.split(tf.PRNGKey(seed), num).tolist()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from mfm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_jax_ffi_shim_binds_declared_symbols():
    """mfm_b200/jax_ffi: shipped as source (no JAX / XLA FFI headers in this image).  Every C-ABI entry point the handlers call is
    declared in include/mfm_b200.h, every handler symbol the registration module names is defined in the .cc, and without JAX the
    module imports cleanly and says why it cannot build."""
    import re
    import pytest
    import mfm_b200.jax_ffi as J
    src = open(os.path.join(ROOT, "mfm_b200", "jax_ffi", "mfm_jax_ffi.cc")).read()
    header = open(os.path.join(ROOT, "include", "mfm_b200.h")).read()
    called = set(re.findall(r"\b(mfm_[a-z0-9_]+)\s*\(", src)) | set(re.findall(r"\?\s*(mfm_[a-z0-9_]+)\s*:\s*(mfm_[a-z0-9_]+)", src).__iter__().__next__())
    assert len(called) >= 8
    for name in called:
        assert re.search(r"\b%s\s*\(" % name, header), f"{name} is not declared in include/mfm_b200.h"
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", src))
    assert defined == set(J.TARGETS.values())
    if not J.HAVE_JAX:
        with pytest.raises(RuntimeError, match="jax is not installed"):
            J.build()
