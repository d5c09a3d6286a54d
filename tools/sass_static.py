"""Static SASS instruction count per solver function of one kernel (nvdisasm -g line info): what the instruction cache has to hold.
python tools/sass_static.py build/obj/nmpc_lmpc.o LmpcAxisELi32ELi20"""
import collections, os, re, subprocess, sys, tempfile
obj, key = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
cur = fn = None
cnt = collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m: fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if fn and key in fn and "solve_kernel" in fn and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        cnt[cur] += 1
src = open(os.path.join(ROOT, "dart-dual-arm-non-prehensile-manipulation_b200/csrc/solver_core.cuh")).read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s*(?:template <class TL>\s*)?DART_HD (?:static )?(?:constexpr )?(?:void|bool|double|int|D2) (\w+)\(", l)
    if m: marks.append((i, m.group(1)))
def phase(ln):
    name = "?"
    for i, n in marks:
        if i <= ln: name = n
    return name
agg = collections.Counter()
for (f, ln), c in cnt.items():
    agg[phase(ln) if f == "solver_core.cuh" else f] += c
print("static SASS instructions:", sum(cnt.values()), f"({sum(cnt.values()) * 16 // 1024} kB)")
for k, v in agg.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 16):
    print(f"{v:6d}  {k}")
