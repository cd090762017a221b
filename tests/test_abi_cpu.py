"""The C-ABI shared library: builds for sm_100a without a GPU, loads, exports every symbol the header
declares with the layout the Python binding assumes, and refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "streammind_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sm_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(built_library):
    from streammind_b200 import lib
    declared = _declared_symbols()
    assert len(declared) >= 18
    assert sorted(lib.SYMBOLS) == declared, set(lib.SYMBOLS) ^ set(declared)
    for name in declared:
        assert getattr(built_library, name) is not None


def test_config_struct_layout_matches_header(built_library):
    from streammind_b200 import lib
    fields = [n for n, _ in lib.SmConfig._fields_]
    prog = ('#include <stdio.h>\n#include <stddef.h>\n#include "streammind_b200.h"\nint main(){printf("%zu", sizeof(sm_config));'
            + "".join(f'printf(" %zu", offsetof(sm_config, {f}));' for f in fields) + "return 0;}")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        vals = [int(v) for v in subprocess.check_output([exe]).split()]
    assert vals[0] == C.sizeof(lib.SmConfig)
    for f, off in zip(fields, vals[1:]):
        assert getattr(lib.SmConfig, f).offset == off, f


def test_sass_contains_blackwell_tensor_and_tma_instructions(built_library):
    from streammind_b200 import lib
    sass = subprocess.run(["cuobjdump", "-sass", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100a" in sass or "sm_100" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):            # tcgen05.mma, TMA load, tcgen05.ld
        assert mnemonic in sass, mnemonic


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(built_library):
    from streammind_b200 import lib
    from streammind_b200.engine import Engine, EngineConfig
    h = C.c_void_p()
    cfg = EngineConfig().to_c()
    assert built_library.sm_create(C.byref(h), 0, C.byref(cfg)) != 0
    assert b"CUDA" in built_library.sm_last_error(None)
    with pytest.raises(RuntimeError):
        Engine(EngineConfig())


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "streammind_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
