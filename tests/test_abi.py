"""CPU tests: the C-ABI library builds, loads and exports what include/rvlnl.h declares."""
import ctypes
import os
import re

import pytest

from evidence_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rvlnl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rvl_[a-z0-9_]+)\s*\(", text)))


def test_header_and_ctypes_table_agree():
    assert declared_symbols() == sorted(_abi.SYMBOLS)


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert _abi.load().rvl_abi_version() == _abi.RVL_ABI_VERSION


def test_struct_sizes_match_the_header(built_lib):
    # sizeof computed from the header by the C compiler must equal the ctypes mirror
    import subprocess
    import tempfile
    src = ('#include <stdio.h>\n#include "rvlnl.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
           'sizeof(rvl_param),sizeof(rvl_planet_desc),sizeof(rvl_model_desc),'
           'sizeof(rvl_prior_desc),sizeof(rvl_counters_t));printf("%zu\\n",sizeof(rvl_slice_args));return 0;}')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [ctypes.sizeof(t) for t in (_abi.rvl_param, _abi.rvl_planet_desc, _abi.rvl_model_desc,
                                       _abi.rvl_prior_desc, _abi.rvl_counters_t, _abi.rvl_slice_args)]
    assert got == want


def test_no_cpu_fallback_without_a_gpu(built_lib):
    """Without a device the product must fail loudly (never route to a CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _abi.load()
    h = ctypes.c_void_p()
    rc = lib.rvl_create(ctypes.byref(h), -1)
    assert rc == -2 and not h.value  # RVL_ENODEV
    assert b"no CPU fallback" in lib.rvl_last_error(None)
    # the multi-device handle and the handle-less entry points refuse just as loudly
    rc = lib.rvl_create_multi(ctypes.byref(h), None, 0)
    assert rc == -2 and not h.value and b"no CPU fallback" in lib.rvl_last_error(None)
    args = _abi.rvl_slice_args()
    args.k, args.d, args.n_out, args.m = 4, 3, 2, 2
    assert lib.rvl_slice_phase(0, ctypes.byref(args), None) != 0
    from evidence_b200 import synth
    from evidence_b200.rvmodel import DeviceError, RVModel
    case = synth.make_case(1)
    with pytest.raises(DeviceError):
        RVModel(case.fixedpardict, case.datadict(), case.parnames)


def test_product_never_imports_the_oracle():
    """The checker is test infrastructure: nothing under evidence_b200/, examples/, tools/ or include/
    may import, load or run it (only tests/, __graft_entry__.smoke() and bench.py's CPU legs do)."""
    pkg = os.path.join(ROOT, "evidence_b200")
    banned = re.compile(r"^\s*(from|import)\s+oracle\b|rv_oracle|librvoracle|oracle/_ref|oracle/_build|"
                        r"trueanomaly\.so|#include\s+\"[^\"]*oracle", re.M)
    for top in (pkg, os.path.join(ROOT, "examples"), os.path.join(ROOT, "tools"), os.path.join(ROOT, "include")):
        for dirpath, _, files in os.walk(top):
            for f in files:
                if f.endswith((".py", ".cu", ".h", ".cpp", ".sh")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not banned.search(text), os.path.join(dirpath, f)
