"""CPU tests (no GPU): the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/ovvc_b200.h declares; host-only entry points work; device entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ovvc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ovgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from openvvc_b200 import api
    lib = api.lib()
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/ovvc_b200.h but not exported: %s" % missing


def test_host_side_alf_table_build_matches_oracle(oracle):
    from openvvc_b200 import api, synth
    for seed in range(4):
        ap = synth.alf_params(256, 128, 400 + seed)
        t = api.alf_tables_from_aps(ap["luma_aps"], ap["chroma_aps"])
        o = oracle.alf_build_tables(ap["luma_aps"], ap["chroma_aps"], impl="port")
        n = 16 + len(ap["luma_aps"]); na = ap["chroma_aps"]["num_chroma_alts"]
        assert (t.np("luma_coeff")[:n] == o.np("luma_coeff")[:n]).all()
        assert (t.np("luma_clip")[:n] == o.np("luma_clip")[:n]).all()
        assert (t.np("chroma_coeff")[:na] == o.np("chroma_coeff")[:na]).all()
        assert (t.np("chroma_clip")[:na] == o.np("chroma_clip")[:na]).all()


def test_struct_layouts_match_header():
    from openvvc_b200 import api
    assert C.sizeof(api.AlfTables) == 2 * 24 * 1300 * 2 + 2 * 8 * 7 * 2 + 2 * 4 * 8 * 2
    assert api.ALF_CTU.itemsize == 4
    assert api.SAO_CTU.itemsize == 40


def test_no_cpu_fallback_without_gpu():
    """On a machine without a B200 the product must refuse to run rather than fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from openvvc_b200 import api
    with pytest.raises(api.OVGPUError):
        api.Ctx(0)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "openvvc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, f
