"""Warp-stall samples of a .ncu-rep (captured with --import-source on) aggregated per CUDA source line: the SASS
listing of the report is joined with the line table nvdisasm prints for the same cubin (no GPU needed).
Usage: python tools/ncu_lines.py report.ncu-rep file.cubin kernel-substring [top]"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep, cubin, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
    line_of, cur, on = [], None, False
    for l in dis:
        if l.startswith("\t.text.") or l.startswith(".text."):
            on = kern in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            inl = re.findall(r'inlined at "[^"]+", line (\d+)', m.group(3))
            cur = (m.group(1).split("/")[-1], int(m.group(2)), tuple(int(x) for x in inl))
            continue
        if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", l):
            line_of.append(cur)
    print(len(rows), "instructions in the report,", len(line_of), "in the cubin")
    n = min(len(rows), len(line_of))
    tot = sum(float(r["# Samples"]) for r in rows)
    agg, why = {}, {}
    stalls = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
    for i in range(n):
        key = line_of[i]
        agg[key] = agg.get(key, 0.0) + float(rows[i]["# Samples"])
        w = why.setdefault(key, {})
        for k in stalls:
            w[k] = w.get(k, 0.0) + float(rows[i][k] or 0)
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        w = sorted(why[key].items(), key=lambda kv: -kv[1])[:2]
        print(f"{100 * v / tot:5.1f}%  {key[0]}:{key[1]} {('<- ' + ','.join(map(str, key[2]))) if key[2] else '':14s} "
              + " ".join(f"{k[6:]}={100 * x / max(v, 1):.0f}%" for k, x in w))


if __name__ == "__main__":
    main()
