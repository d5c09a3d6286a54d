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
