"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/pfmds_b200.h declares.
No compute call is made here; without a CUDA device pfmds_create must fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "pfmds_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pfmds_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_are_exported(cuda_lib):
    lib = C.CDLL(cuda_lib)
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_library_is_sm100a_only(cuda_lib):
    out = subprocess.run(["cuobjdump", "-lelf", cuda_lib], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback(cuda_lib):
    import torch
    if torch.cuda.is_available():
        return
    from pfmds_b200.engine import Engine, PfmdsError
    try:
        Engine(np.zeros((4, 3)), np.zeros((4, 3)), np.ones(4), np.ones(3) * 10)
    except PfmdsError as e:
        assert e.code == 2 and "no CPU fallback" in str(e)
    else:
        raise AssertionError("pfmds_create succeeded without a CUDA device")


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under pfmds_b200/ may include, link, import or execute it."""
    pat = re.compile(r'#include\s*[<"][^>"]*oracle|import\s+[^\n]*oracle|from\s+[^\n]*oracle[^\n]*import|liboracle|oracle/|oracle_run_md')
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "pfmds_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h", ".f90")):
                if pat.search(open(os.path.join(d, f), errors="ignore").read()):
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_kernels_launched_with_1024_threads_fit_the_register_file(cuda_lib):
    """A block of 1024 threads can use at most 64 registers per thread (65 536 per SM): a kernel that grows past that fails at launch
    ("too many resources requested"), which only a GPU run would show.  Checked from the cubin's resource usage."""
    out = subprocess.run(["cuobjdump", "-res-usage", cuda_lib], stdout=subprocess.PIPE, text=True).stdout
    big = ("k_nhc_close", "k_nhciPKd", "k_reduce_ke_partials", "k_scan_block", "k_scan_sums", "k_sl_scan_block", "k_sum_partials", "k_sum_to", "k_sums_final")
    seen = 0
    lines = out.splitlines()
    for k, l in enumerate(lines):
        if "Function" in l and any(b in l for b in big):
            m = re.search(r"REG:(\d+)", lines[k + 1])
            assert m, lines[k + 1]
            assert int(m.group(1)) <= 64, (l, lines[k + 1])
            seen += 1
    assert seen >= 8
