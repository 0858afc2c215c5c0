"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol that
include/rvtests_b200.h declares, and fails LOUDLY (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    h = open(os.path.join(ROOT, "include", "rvtests_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(rvt_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_exported():
    from rvtests_b200 import engine
    L = engine.load_library()
    names = declared_functions()
    assert len(names) >= 19
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/rvtests_b200.h but not exported"
    assert sorted(engine.EXPORTS) == names


def test_result_struct_layout_matches_header():
    from rvtests_b200 import engine
    # compile a probe with the real header and compare sizeof/offsetof with the ctypes mirror
    src = r'''
#include <stddef.h>
#include <stdio.h>
#include "rvtests_b200.h"
int main(){printf("%zu %zu %zu %zu %zu %zu\n", sizeof(rvt_gene_result), offsetof(rvt_gene_result,davies_fault),
 offsetof(rvt_gene_result,cmc_U), offsetof(rvt_gene_result,zeg_U), offsetof(rvt_gene_result,skato_Q), offsetof(rvt_gene_result,lambda_max));return 0;}
'''
    exe = "/tmp/rvt_abi_probe"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    got = [int(x) for x in subprocess.run([exe], capture_output=True, check=True).stdout.split()]
    G = engine.GeneResult
    want = [C.sizeof(G), G.davies_fault.offset, G.cmc_U.offset, G.zeg_U.offset, G.skato_Q.offset, G.lambda_max.offset]
    assert got == want


def test_library_is_sm100a_native():
    from rvtests_b200 import build
    lib = build.build_lib()
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rvtests_b200
    with pytest.raises(rvtests_b200.RvtError) as ei:
        rvtests_b200.GeneEngine(0)
    assert "no CPU fallback" in str(ei.value)
