"""Pretty-print the headline fields of bench.py JSON lines read from stdin: python bench.py | python tools/benchline.py"""
import json, sys
for ln in sys.stdin:
    ln = ln.strip()
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    r = d.get("roofline") or {}
    print("value %.3f G/s  ms/step %.1f  kernel_ms %s  frac %s  e2e %.3f G/s  clocks %s  cpu %s" % (
        d["value"] / 1e9, d["ms_per_step"], r.get("kernel_ms"), r.get("frac"), d["e2e"]["value"] / 1e9,
        d.get("clocks"), (d.get("cpu_baseline") or {}).get("value")))
