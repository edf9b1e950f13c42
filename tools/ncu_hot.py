"""Per-instruction hot spots from `ncu --page source --csv` of one kernel.
    python tools/ncu_hot.py rep.ncu-rep kernel_regex [min_pct]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr) and r[0] != "Address"]
ix = {h: i for i, h in enumerate(hdr)}
iex, ist, isrc = ix["Instructions Executed"], ix["Warp Stall Sampling (All Samples)"], ix["Source"]
itag = ix.get("L1 Tag Requests Global"); iw = ix.get("L1 Wavefronts Shared"); ith = ix.get("Avg. Threads Executed")
tot = sum(int(r[iex]) for r in data) or 1; tots = sum(int(r[ist]) for r in data) or 1
print("instructions", tot, "stall samples", tots, "sass lines", len(data))
for k, r in enumerate(data):
    ex, st = int(r[iex]), int(r[ist])
    if 100 * ex / tot >= minp or 100 * st / tots >= 2 * minp:
        print("%4d %-58s ex %5.2f%% stall %5.2f%% thr %5s tags %9s shwf %9s" % (
            k, r[isrc].strip()[:58], 100 * ex / tot, 100 * st / tots, r[ith][:5] if ith else "", r[itag] if itag else "", r[iw] if iw else ""))
