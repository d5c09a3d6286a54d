"""Group the stall samples of an ncu report by solver phase (function in solver_core.cuh): python tools/ncu_phases.py rep"""
import collections, csv, re, subprocess, sys, os
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
sections, cur = [], None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = {"file": r[1], "rows": []}; sections.append(cur)
    elif len(r) > 10 and r[0] == "Line No": cur["hdr"] = r
    elif cur is not None and len(r) > 10 and "hdr" in cur: cur["rows"].append(r)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "dart-dual-arm-non-prehensile-manipulation_b200/csrc/solver_core.cuh")).read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s*(?:template <class TL>\s*)?DART_HD (?:static )?(?:constexpr )?(?:void|bool|double|int) (\w+)\(", l)
    if m: marks.append((i, m.group(1)))
def phase(ln):
    name = "?"
    for i, nme in marks:
        if i <= ln: name = nme
    return name
seen = set(); ph = collections.Counter(); ins = collections.Counter(); tot = toti = 0
for s in sections:
    f = s["file"].split("/")[-1]
    if f in seen: continue
    seen.add(f)
    h = s["hdr"]; iL, iN, iI = h.index("Line No"), h.index("# Samples"), h.index("Instructions Executed")
    for r in s["rows"]:
        try: ln, sm, ii = int(r[iL]), int(r[iN] or 0), int(r[iI] or 0)
        except ValueError: continue
        key = phase(ln) if f == "solver_core.cuh" else f
        ph[key] += sm; ins[key] += ii; tot += sm; toti += ii
print("samples", tot, "warp-inst", toti)
for k, v in ph.most_common(16):
    print(f"{100*v/tot:5.1f}% samples  {100*ins[k]/toti:5.1f}% inst   {k}")
