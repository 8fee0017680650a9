"""SURVEY.md 8(f) row 3 — the consumers of md(): run_gr_moire_fitting (golden-section fit of the ljc / morsec parameters, two md()
relaxations per evaluation, file renames; fit_gr_moire.f90, run_gr_moire_fitting.f90) and run_gr_analysis
(graphene_on_surface_analysis.f90).  CPU side: the shared host logic on the oracle engine."""
import os
import subprocess

import numpy as np
import pytest

from pfmds_b200 import inputs
from pfmds_b200.build import EXE_ANALYSIS
from pfmds_b200.host_io import read_xyz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_FIT = os.path.join(ROOT, "oracle", "_build", "oracle_run_gr_moire_fitting")
GOLD = (np.sqrt(5.0) - 1.0) / (np.sqrt(5.0) + 1.0)


def cell(seed, jitter, interface="ljc", steps=30):
    c = inputs.graphene_on_cu_small(interface=interface, period=5, seed=seed, jitter=jitter)
    c["integrators"] = [("nvms", 1.0, steps, 10 ** 9, 10)]
    return c


def fit_rows(path, interface="ljc"):
    npar = 3 if interface == "ljc" else 4
    rows = []
    for l in open(path).read().splitlines():
        v = l.split()
        rows.append(dict(sim=int(v[0]), params=[float(x) for x in v[1:1 + npar]], cells=[float(x) for x in v[1 + npar:-1]], error=float(v[-1])))
    return rows


def run_fit(exe, d, interface="ljc", extra=(), gold="0.02"):
    f = inputs.write_fitting_inputs(d, cell(3, 0.02, interface), cell(4, 0.03, interface), interface=interface, be0=-0.05, grd0=(0.03, 0.03),
                                    zero_level=(-1677.4, -1677.2))
    r = subprocess.run([exe, "-fpfn", f, "-op", "1000", "-omp_n", "2", "-delta_error_gold", gold, "-delta_error_fit", "1e9", *extra],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    return r.stdout


def test_fit_loop_on_the_oracle_engine(tmp_path, oracle_lib):
    d = str(tmp_path) + os.sep
    out = run_fit(ORACLE_FIT, d)
    rows = fit_rows(d + "fit_fit_out.txt")
    assert [r["sim"] for r in rows] == list(range(1, len(rows) + 1)) and len(rows) > 12        # brackets + golden-section refinements
    # the first bracket of the first parameter (sigma): min, max, min + gold*w, max - gold*w around the mid point of the others
    lo, hi = 0.8 * 3.0, 1.2 * 3.0
    assert np.allclose([r["params"][1] for r in rows[:4]], [lo, hi, lo + (hi - lo) * GOLD, hi - (hi - lo) * GOLD], atol=1e-6)
    assert all(abs(r["params"][0] - 0.02) < 1e-9 and abs(r["params"][2] - 2.0) < 1e-9 for r in rows[:4])
    # every row: error = (be/be0-1)^2 + (grd1/grd0-1)^2 + (grd2/grd0-1)^2 from its own columns (be bd grd e e | be bd grd e)
    for r in rows:
        c = r["cells"]
        assert abs(c[3] + c[4] + c[8] - r["error"]) < 3e-6
        assert abs((c[0] / -0.05 - 1) ** 2 - c[3]) < 2e-3 * max(1.0, c[3]) and abs((c[2] / 0.03 - 1) ** 2 - c[4]) < 2e-3 * max(1.0, c[4])
    # file protocol: parameter files numbered, start files restored, per-evaluation outputs
    for k in range(1, len(rows) + 1):
        assert os.path.exists(d + "%06dparameters_LJC_C-Cu.txt" % k)
        assert os.path.exists(d + "fit_%06d_final_a.txt" % k) and os.path.exists(d + "fit_%06d_final_cell_b.xyz" % k)
    assert os.path.exists(d + "start_a.xyz") and os.path.exists(d + "start_b.xyz") and not os.path.exists(d + "cell_a.xyz")
    assert not os.path.exists(d + "parameters_LJC_C-Cu.txt")
    # refinement evaluations restart from the previous relaxed cell: that file is handed back after the run
    assert os.path.exists(d + "fit_%06d_final_cell_a.xyz" % (len(rows) - 1))
    # the all_out row + analysis in one record; the binding energy read back from columns 62-81
    l = open(d + "fit_000001_final_a.txt").read().splitlines()[0]
    be = (float(l[61:81]) + 1677.4) / 96
    assert abs(be - rows[0]["cells"][0]) < 1e-6
    assert " parameters: " in out and out.count("delta_error:") >= 3 + (len(rows) - 12)


def test_fit_pair_mode_gives_the_same_rows(tmp_path, oracle_lib):
    a, b = str(tmp_path / "seq") + os.sep, str(tmp_path / "pair") + os.sep
    run_fit(ORACLE_FIT, a, gold="1e9")
    run_fit(ORACLE_FIT, b, gold="1e9", extra=("-pair",))
    ra, rb = fit_rows(a + "fit_fit_out.txt"), fit_rows(b + "fit_fit_out.txt")
    assert len(ra) == len(rb) == 12
    for x, y in zip(ra, rb):
        assert np.allclose(x["params"], y["params"]) and np.allclose(x["cells"], y["cells"], atol=2e-6) and abs(x["error"] - y["error"]) < 3e-6


def test_fit_morsec_parameter_order(tmp_path, oracle_lib):
    d = str(tmp_path) + os.sep
    run_fit(ORACLE_FIT, d, interface="morsec", gold="1e9")
    rows = fit_rows(d + "fit_fit_out.txt", "morsec")
    assert len(rows) == 12
    p = [float(x) for x in open(d + "000001parameters_MorseC_C-Cu.txt").read().split()[:4]]     # d r a delt, written as params(3) (1) (4) (2)
    assert np.allclose(p, rows[0]["params"], atol=1e-6) and np.isclose(p[1], 0.8 * 3.2)
    assert open(d + "000001parameters_MorseC_C-Cu.txt").read().split()[6] == "F"


def test_gr_analysis(tmp_path, cuda_lib):
    case = inputs.graphene_on_cu_small()
    d = str(tmp_path) + os.sep
    inputs.write_xyz(d + "a.xyz", case)
    case["pos"][:, 2] += 0.5
    inputs.write_xyz(d + "b.xyz", case)
    open(d + "filelist.txt", "w").write("2\na.xyz\nb.xyz\n")
    r = subprocess.run([EXE_ANALYSIS, "-path", d, "-z", "0.0"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    names = np.asarray(case["names"])
    x = read_xyz(d + "a.xyz")
    zc, zcu = x["pos"][names == "C", 2], x["pos"][names != "C", 2]
    rows = [l.split() for l in open(d + "outfilename.txt").read().splitlines()]
    assert rows[0][0] == "a.xyz" and rows[1][0] == "b.xyz"
    want = [zc.mean() - zcu.mean(), zc.min() - zcu.mean(), zc.max() - zcu.mean()]
    assert np.allclose([float(v) for v in rows[0][1:]], want, rtol=0, atol=1e-12)
    assert np.allclose([float(v) for v in rows[1][1:]], want, rtol=0, atol=1e-12)      # a rigid shift changes nothing
