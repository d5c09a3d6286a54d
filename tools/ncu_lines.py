"""Aggregate an ncu source-page CSV by source line: python tools/ncu_lines.py report.ncu-rep [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
sections, cur = [], None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}; sections.append(cur)
    elif len(r) > 10 and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and len(r) > 10 and "hdr" in cur:
        cur["rows"].append(r)
seen = set(); tot = collections.Counter(); samp = collections.Counter(); srcs = {}
for s in sections:
    f = s["file"].split("/")[-1]
    if f in seen: continue      # the page repeats per function view
    seen.add(f)
    h = s["hdr"]; iL, iS, iI, iN = h.index("Line No"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    for r in s["rows"]:
        try: ln, ins, sm = int(r[iL]), int(r[iI] or 0), int(r[iN] or 0)
        except ValueError: continue
        tot[(f, ln)] += ins; samp[(f, ln)] += sm; srcs[(f, ln)] = r[iS].strip()[:100]
T, S = sum(tot.values()), sum(samp.values())
print("total warp-inst", T, "samples", S)
for k, v in samp.most_common(top):
    print(f"{100*v/max(S,1):5.1f}% samp {100*tot[k]/max(T,1):5.1f}% inst  {k[0]}:{k[1]}  {srcs[k]}")
