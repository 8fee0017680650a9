"""SURVEY.md 8(f) row 3 on the device: run_gr_moire_fitting with every md() relaxation on the GPU against the same fit on the
CPU oracle engine — same evaluations, same fit_out.txt rows to the printed digits.  (Named zz: it runs after the parity tests.)"""
import os

import numpy as np
import pytest

from pfmds_b200.build import EXE_FIT
from test_fitting import ORACLE_FIT, fit_rows, run_fit

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("extra", [(), ("-pair",)])
def test_fit_rows_match_the_cpu_port(tmp_path, cuda_lib, oracle_lib, extra):
    g, c = str(tmp_path / "gpu") + os.sep, str(tmp_path / "cpu") + os.sep
    run_fit(EXE_FIT, g, gold="1e9", extra=extra)
    run_fit(ORACLE_FIT, c, gold="1e9")
    rg, rc = fit_rows(g + "fit_fit_out.txt"), fit_rows(c + "fit_fit_out.txt")
    assert len(rg) == len(rc) == 12
    for x, y in zip(rg, rc):
        assert np.allclose(x["params"], y["params"], atol=1e-6)
        assert np.allclose(x["cells"], y["cells"], rtol=1e-5, atol=3e-6) and abs(x["error"] - y["error"]) < 1e-5 * max(1.0, abs(y["error"]))
