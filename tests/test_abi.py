"""The C-ABI library loads here (no GPU) and exports every symbol include/dart_b200.h declares."""
import ctypes as C
import os
import re

import dart_b200
from tests.helpers import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "dart_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dart_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    lib = C.CDLL(dart_b200._lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dart_b200.h but not exported"


def test_cfg_struct_matches_header(built):
    # dart_default_cfg is pure host code: compare it with the python-side builders field by field
    lib = dart_b200._lib.lib()
    for method, builder in ((0, dart_b200.pmpc_cfg), (1, dart_b200.rmpc_cfg), (2, dart_b200.lmpc_cfg)):
        c = dart_b200.DartCfg()
        assert lib.dart_default_cfg(method, C.byref(c)) == 0
        p = builder()
        for name, _ in dart_b200.DartCfg._fields_:
            a, b = getattr(c, name), getattr(p, name)
            if hasattr(a, "__len__"):
                assert list(a) == list(b), name
            else:
                assert a == b, name


def test_no_device_is_an_error_not_a_fallback(built):
    import torch
    if torch.cuda.is_available():
        return
    try:
        dart_b200.NMPCEngine(dart_b200.pmpc_cfg())
    except dart_b200.DartError as e:
        assert "no CUDA device" in str(e)
    else:
        raise AssertionError("engine creation must fail without a GPU")


def test_yaml_config_roundtrip():
    c = dart_b200.cfg_from_yaml("pmpc")
    assert (c.N, c.Qp, c.Qv, c.R, c.mu, c.u_lo, c.u_hi) == (15, 400.0, 2.0, 0.2, 0.1, -0.6, 0.6)
    c = dart_b200.cfg_from_yaml("rmpc")
    assert (c.N, c.Qp, c.R, c.Rdu, c.du_hi, c.vmax) == (20, 80.0, 0.02, 1.0, 0.06, 0.2)
    c = dart_b200.cfg_from_yaml("lmpc")
    assert list(c.Q)[:4] == [200.0, 2.0, 200.0, 2.0] and list(c.Rl) == [0.1, 0.1, 1.0, 1.0] and c.u_hi == 0.4


def test_header_is_plain_c99(tmp_path):
    """include/dart_b200.h is the boundary a C host binds: it must compile as C99 on its own, with no CUDA or C++."""
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text("""
#include "dart_b200.h"
int use(void) {
    dart_cfg cfg; dart_handle h = 0; double x[6] = {0}, r[6] = {0}, u[2], J[1]; int32_t st[1], it[1];
    if (dart_default_cfg(DART_PMPC, &cfg) != DART_OK) return 1;
    if (dart_create(&h, &cfg, 0) != DART_OK) return 2;
    if (dart_solve_host(h, 1, x, r, 0, 0, 0, u, J, st, it) != DART_OK) return 3;
    return dart_destroy(h);
}
""")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
