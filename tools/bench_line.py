"""One-line digest of a bench.py JSON line on stdin (dev tool for same-box A/B runs): python bench.py | python tools/bench_line.py TAG"""
import json, sys
d = json.loads([l for l in sys.stdin if l.startswith("{")][-1])
tv = d.get("throughput_variant") or {}
print(sys.argv[1] if len(sys.argv) > 1 else "", f"step {d['ms_per_step']:.5f} ms  value {d['value'] / 1e6:.2f} M  e2e {d['e2e']['value'] / 1e6:.2f} M  "
      f"filled {tv.get('ms', float('nan')):.3f} ms  converged {d.get('converged')}  frac {d['roofline']['frac']:.3f}")
