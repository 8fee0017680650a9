"""The host's Fortran emulation: list-directed READ grammar (settings / parameter / xyz files are read with
`read(u,*)`, md_simulation.f90:48-93) and the fixed edit descriptors of the log and xyz writers (fW.D, esW.D, iW.M, AW)."""
import ctypes as C
import os
import subprocess

import pytest

from pfmds_b200 import inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fio") / "libfio.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "fio_host.cpp")], check=True)
    L = C.CDLL(out)
    L.fio_real.restype = C.c_double
    L.fio_real.argtypes = [C.c_char_p]
    L.fio_F.argtypes = [C.c_double, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.fio_ES.argtypes = [C.c_double, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.fio_I.argtypes = [C.c_long, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.fio_A.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.fio_LR.argtypes = [C.c_double, C.c_char_p, C.c_int]
    return L


def _s(fn, *a):
    buf = C.create_string_buffer(4096)
    fn(*a, buf, 4096)
    return buf.value.decode()


def test_fixed_edit_descriptors(lib):
    assert _s(lib.fio_F, 3.14159, 10, 4) == "    3.1416"
    assert _s(lib.fio_F, -0.5, 24, 6) == " " * 15 + "-0.500000"
    assert _s(lib.fio_F, 123456.789, 8, 3) == "********"                # does not fit: asterisks, like Fortran
    assert _s(lib.fio_F, 12.0, 27, 16) == " " * 8 + "12.0000000000000000"
    assert _s(lib.fio_ES, 1e-8, 16, 6) == "    1.000000E-08"
    assert _s(lib.fio_ES, -12345.678, 21, 9) == "     -1.234567800E+04"
    assert _s(lib.fio_I, 42, 9, 0) == "       42"
    assert _s(lib.fio_I, 42, 0, 6) == "000042"                         # i6.6 of snapshot_000042.xyz
    assert _s(lib.fio_I, 7, 0, 4) == "0007"                             # i4.4 of the per-rank prefix
    assert _s(lib.fio_A, b"nvt", 6, 3) == "   nvt"                      # A6 of a trimmed name: right-justified
    assert _s(lib.fio_A, b"md_step_limit:", 32, 128) == "md_step_limit:" + " " * 18   # A32 of a character(128): leftmost 32
    assert _s(lib.fio_A, b"CU_fixed", 12, 32) == "CU_fixed    "
    assert len(_s(lib.fio_LR, -6.2588955742945984e-3)) == 26 and "E-003" in _s(lib.fio_LR, -6.2588955742945984e-3)


def test_list_directed_grammar(lib, tmp_path):
    p = tmp_path / "f.txt"
    p.write_text("label:\t 12 , 3.5d0 'quoted string' T extra tokens are dropped\n"
                 "a b\n"
                 "   c   / the slash ends the record\n"
                 "1,2,,3\n"
                 "\"it''s\" .false. .TRUE. f\n")
    buf = C.create_string_buffer(4096)
    assert lib.fio_records(str(p).encode(), 1, 5, buf, 4096) == 0
    assert buf.value.decode().strip().split("\x1f") == ["label:", "12", "3.5d0", "quoted string", "T"]
    # a record that needs 3 items continues on the next line; '/' leaves the rest untouched
    assert lib.fio_records(str(p).encode(), 2, 3, buf, 4096) == 0
    recs = buf.value.decode().split("\n")
    assert recs[0].split("\x1f") == ["label:", "12", "3.5d0"] and recs[1].split("\x1f") == ["a", "b", "c"]
    assert lib.fio_real(b"3.5d0") == 3.5 and lib.fio_real(b"1.e-8") == 1e-8 and lib.fio_real(b"100.") == 100.0
    assert [lib.fio_logical(t) for t in (b"T", b"F", b".true.", b".FALSE.", b"t", b"x")] == [1, 0, 1, 0, 1, -1]


def test_settings_grammar_is_the_code_order_not_the_readme_order(lib, tmp_path):
    """md() reads all_moving/xyz/z/all_atoms/traj (no termo_atoms line) and nhc_num + one line per thermostat."""
    case = inputs.graphene_on_cu_small()
    d = str(tmp_path) + os.sep
    inputs.write_case(d, case)
    buf = C.create_string_buffer(4096)
    assert lib.fio_settings(d.encode(), b"md_run_settings.txt", buf, 4096) == 0, buf.value
    f = buf.value.decode().split("|")
    assert f[:8] == ["2000", "md_run.log", "init.xyz", "F", "5", "2", "1", "3"]
    assert f[8:] == ["tb:10:1", "ljc:6:3", "rjl:7:1"]
    # the stale README layout (an extra termo_atoms_group_num line) shifts every later read: rejected, not misparsed silently
    txt = open(d + "md_run_settings.txt").read().replace("all_atoms_group_num:", "termo_atoms_group_num: 1\nall_atoms_group_num:")
    open(d + "stale.txt", "w").write(txt)
    assert lib.fio_settings(d.encode(), b"stale.txt", buf, 4096) == 1
    # unknown interaction name: the reference prints 'error: unknown interaction name'
    open(d + "bad.txt", "w").write(open(d + "md_run_settings.txt").read().replace("rjl parameters", "eam parameters"))
    assert lib.fio_settings(d.encode(), b"bad.txt", buf, 4096) == 1 and b"unknown interaction name" in buf.value
