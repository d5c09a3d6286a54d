import os, sys, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch, dart_b200
from importlib import import_module
_lib = dart_b200._lib; L = _lib.lib(); W = dart_b200.workloads
dev = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for method in ("pmpc", "rmpc", "lmpc"):
    if method == "pmpc":
        c = W.pmpc_config2(8, seed=3); aux = np.stack([c["Qp"], c["Qv"], c["R"], c["mu"]], axis=1); x0, ref, cfg = c["state"], c["target"], dart_b200.pmpc_cfg()
    elif method == "rmpc":
        d = W.rmpc_inputs(96, seed=5); x0, ref, aux, cfg = d["x0"], d["ref"], d["aux"], dart_b200.rmpc_cfg()
    else:
        d = W.lmpc_inputs(160, seed=7); x0, ref, aux, cfg = d["x0"], d["ref"], d["aux"], dart_b200.lmpc_cfg()
    B = x0.shape[0]
    eng = dart_b200.NMPCEngine(cfg, device=0)
    rows = torch.zeros((B, 4), dtype=torch.float64, device=dev); eng.set_result_rows(rows)
    gp = C.c_void_p(); gh = C.create_string_buffer(64)
    assert L.dart_peer_alloc(B * 32, C.byref(gp), gh) == 0
    arr = (C.c_void_p * 1)(gp)
    assert L.dart_set_result_rows_peers(eng._h, arr, 1, 0) == 0
    g = torch.as_tensor(dart_b200.parallel._DevView(gp.value, (B, 4), "<f8"), device=dev)
    out = eng.solve_device(t(x0), t(ref), aux=t(aux)); torch.cuda.synchronize()
    print(method, B, bool(torch.equal(g, rows)), float(g.abs().sum()), float(rows.abs().sum()), eng.last_launch_config())
