"""Freeze outputs of the CPU oracle for the small parity cases into tests/golden/*.npz.

The reference (Fortran) cannot be built or run in this image and ships no golden data of its own, so
these fixtures pin the ORACLE (C++ restatement), not the reference: they guard the restatement and the
CUDA path against regressions, and travel to the GPU box where /root/reference does not exist.
Run from the repo root:  python tests/golden/make_golden.py [case names...]   (no names: every case)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from util import list_ids, neighbours, next_row_cases, oracle, small_cases  # noqa: E402

if __name__ == "__main__":
    cases = dict(small_cases(), **next_row_cases())
    for name, case in cases.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        o = oracle(case)
        integ = case["integrators"][0][0]
        dt = case["integrators"][0][1]
        o.advance(integ, dt, 0, 1)
        out = {"pos0": case["pos"], "vel0": case["vel"], "frc0": o.download()[2], "e0": o.energies()[0], "ke0": o.energies()[1]}
        for k, j in list_ids(case):
            nl = neighbours(o, case, k, j)
            out["nnum_%d_%d" % (k, j)] = nl[1]
            out["nlist_%d_%d" % (k, j)] = nl[0].astype(np.int16 if nl[0].max() < 32000 else np.int32)
        o.advance(integ, dt, 1, 10)
        p, v, f = o.download()
        out.update(pos10=p, vel10=v, frc10=f, e10=o.energies()[0], ke10=o.energies()[1])
        if case.get("changes"):  # deposition: the sizes of the groups after step 10 and the rebuilt list of that step
            out["group_n10"] = np.array([o.group_size(g + 1) for g in range(len(case["groups"]))])
            nl = neighbours(o, case, 0, 0)
            out["nnum10_0_0"] = nl[1]
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
        print(name, "ok", {k: v.shape for k, v in out.items() if k.startswith("frc")})
