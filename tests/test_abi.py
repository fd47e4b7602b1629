"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/embtab_b200.h declares (no compute without a GPU), struct layouts match between the
header, the CUDA side and the ctypes mirror, and the host mirror refuses to run without CUDA."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "embtab_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(etb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from embtab import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def test_library_exports_every_declared_symbol(lib):
    l = C.CDLL(lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(l, name), f"{name} declared in the header but not exported"
    assert sorted(lib.exported_symbols()) == names, "ctypes mirror and header disagree"
    assert lib.lib().etb_version() == 100


def test_struct_layouts_match_header(lib, tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "embtab_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(etb_table),sizeof(etb_lookup_item),sizeof(etb_update_item),sizeof(etb_index_view));return 0;}')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(lib.Table), C.sizeof(lib.LookupItem), C.sizeof(lib.UpdateItem), C.sizeof(lib.IndexView)]


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "embtab_b200.h"\nint main(void){return ETB_OK;}')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_argument_validation_without_gpu(lib):
    l = lib.lib()
    # bad arguments are rejected before any CUDA call is made
    assert l.etb_maplookup(None, -1, None) == 1
    assert b"negative" in l.etb_last_error()
    assert l.etb_maplookup(None, 0, None) == 0
    t = lib.Table(None, None, 10, 0, 0, 0, lib.F32, 0)
    it = lib.LookupItem(t, None, None, 4, 1, 0, 0, lib.I64, 0)
    assert l.etb_maplookup(C.byref(it), 1, None) == 1 and b"dim" in l.etb_last_error()
    assert l.etb_pooled_sum(None, 4, C.byref(t), None, lib.I64, 0, 1, 0, None) == 1


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import embtab
    with pytest.raises(embtab.EmbTabError):
        embtab.SimpleEmbedding(np.zeros((4, 4), np.float32))
    with pytest.raises(embtab.EmbTabError):
        embtab.DeviceArray.empty((2, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "embeddingtables.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f == "README.md", f"{f} mentions the oracle"


def build_abi_smoke(out_path):
    """tests/abi_smoke.c -> a plain-C executable linked against libembtab_b200.so (no Python, no torch)"""
    import subprocess
    lib_dir = os.path.join(ROOT, "embeddingtables.jl_b200", "lib")
    cmd = ["gcc", "-std=c99", "-O1", "-ffp-contract=off", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "abi_smoke.c"), "-L", lib_dir, "-lembtab_b200", f"-Wl,-rpath,{lib_dir}", "-lm",
           "-o", out_path]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return out_path


def test_abi_smoke_compiles_and_links_as_plain_c(tmp_path):
    # the header is enough for a C compiler and every entry point the program uses resolves against the .so
    exe = build_abi_smoke(str(tmp_path / "abi_smoke"))
    assert os.path.exists(exe)
