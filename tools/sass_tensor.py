"""Count the tcgen05 / TMEM / TMA mnemonics per kernel in the shipped library (cuobjdump -sass): evidence that the tensor-core
kernels are what they say.  python tools/sass_tensor.py > profiles/r2_sass_tensor.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "dart-dual-arm-non-prehensile-manipulation_b200", "lib", "libdart_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UTCATOMSWS|SYNCS|UTMAPF|DFMA|FFMA2)\b")
cnt = collections.defaultdict(collections.Counter)
fn = "?"
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::", "", fn)[:110]
        continue
    m = pat.search(line)
    if m:
        cnt[fn][m.group(1)] += 1
print("# tcgen05 (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld), TMA (UTMALDG), mbarrier (SYNCS) and FP64/FFMA2")
print("# mnemonic counts per kernel of lib/libdart_b200.so (cuobjdump -sass)")
for k in sorted(cnt):
    c = cnt[k]
    if any(x in c for x in ("UTCHMMA", "LDTM", "UTMALDG")) or ("nmpc_solve_kernel" in k and (", 32, 20>" in k or "PmpcAxis, 16, 15>" in k)):
        print(k, dict(c))
