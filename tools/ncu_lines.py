"""Per CUDA source line: executed warp instructions and stall samples of one launch of an ncu report.
    python tools/ncu_lines.py rep.ncu-rep <launch-skip> [top]"""
import csv, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
ix = {h: k for k, h in enumerate(hdr)}
ex, st = ix["Instructions Executed"], ix["Warp Stall Sampling (All Samples)"]
lines = []
for r in rows:
    if len(r) == len(hdr) and r[0].isdigit():
        lines.append((int(r[0]), r[1], int(r[ex]), int(r[st])))
te = sum(l[2] for l in lines) or 1
ts = sum(l[3] for l in lines) or 1
print("total warp instr %.1f M, stall samples %d" % (te / 1e6, ts))
sel = sorted(set(sorted(range(len(lines)), key=lambda k: -lines[k][2])[:top]) | set(sorted(range(len(lines)), key=lambda k: -lines[k][3])[:top]))
for k in sel:
    l = lines[k]
    print("%4d  ex %5.1f%%  st %5.1f%%  %s" % (l[0], 100.0 * l[2] / te, 100.0 * l[3] / ts, l[1].strip()[:110]))

if len(sys.argv) > 4:      # phase table: "name:lo-hi,name:lo-hi,..."
    print()
    for spec in sys.argv[4].split(","):
        name, rng = spec.split(":")
        lo, hi = (int(x) for x in rng.split("-"))
        e = sum(l[2] for l in lines if lo <= l[0] <= hi)
        s = sum(l[3] for l in lines if lo <= l[0] <= hi)
        print("%-10s lines %4d-%4d  ex %5.1f%%  st %5.1f%%" % (name, lo, hi, 100.0 * e / te, 100.0 * s / ts))
