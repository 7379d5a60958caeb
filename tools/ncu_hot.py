"""Hottest SASS instructions (warp-stall samples) of a .ncu-rep captured with --import-source on; run here, no GPU needed."""
import csv
import io
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    print(lines[start - 1][:140])
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    tot = sum(float(r["# Samples"]) for r in rows)
    print(len(rows), "instructions,", int(tot), "samples")
    stalls = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(float(r[k] or 0) for r in rows) for k in stalls}
    print("stalls:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    idx = {r["Address"]: i for i, r in enumerate(rows)}
    for r in sorted(rows, key=lambda r: -float(r["# Samples"]))[:top]:
        why = max(stalls, key=lambda k: float(r[k] or 0))
        print(f"{idx[r['Address']]:5d} {100 * float(r['# Samples']) / tot:5.1f}%  {why[6:]:12s} {r['Source'][:100]}")


if __name__ == "__main__":
    main()
