"""A/B builds of the solve kernels: python tools/build_variant.py NAME [-DFLAG=..]...  ->  build/variants/libdart_b200_NAME.so
(recompiles nmpc_rmpc.cu / nmpc_lmpc.cu / nmpc_pmpc.cu with the extra flags, reuses the other objects of the last build()).
Run a tool against it with DART_B200_LIB=build/variants/libdart_b200_NAME.so (dev tool)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
name, flags = sys.argv[1], sys.argv[2:]
g.build_cuda()
obj = os.path.join(ROOT, "build", "obj")
out = os.path.join(ROOT, "build", "variants"); os.makedirs(out, exist_ok=True)
procs, objs = [], []
for f in sorted(os.listdir(obj)):
    if not f.endswith(".o"):
        continue
    if f in ("nmpc_rmpc.o", "nmpc_lmpc.o", "nmpc_pmpc.o"):
        o = os.path.join(out, f"{name}_{f}")
        procs.append(subprocess.Popen(["nvcc"] + g.NVCC_FLAGS + flags + ["-c", os.path.join(g.CSRC, f[:-2] + ".cu"), "-o", o]))
        objs.append(o)
    else:
        objs.append(os.path.join(obj, f))
for p in procs:
    if p.wait() != 0:
        sys.exit(1)
so = os.path.join(out, f"libdart_b200_{name}.so")
subprocess.check_call(["nvcc", "-shared", "-o", so] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
print(so)
