"""Top stall sites per kernel from an `ncu --set full --import-source on` report (SASS view of the source page).
    python tools/ncu_hotspots.py gpurun_out/prof_x.ncu-rep "title" > profiles/rNN_hotspots.md"""
import csv
import subprocess
import sys


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    print("# %s\n" % title)
    print("Source: `%s`, SASS view of `ncu --page source`; per kernel the instructions with the most warp-stall samples "
          "(`st` = share of the kernel's samples, `ex` = share of its executed warp instructions, `thr` = average live "
          "threads). A stall is charged to the instruction that could not issue, i.e. the first *consumer* of a pending "
          "load.\n" % rep)
    i, seen = 0, set()
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1].replace("void ", "").replace("(bool)", "")
            name = name[:name.index(">(") + 1] if ">(" in name else name.split("(")[0]
            hdr = rows[i + 1]
            j = i + 2
            data = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if len(rows[j]) == len(hdr):
                    data.append(rows[j])
                j += 1
            i = j
            if name in seen:
                continue
            seen.add(name)
            ix = {h: k for k, h in enumerate(hdr)}
            ex, st, src, th = (ix["Instructions Executed"], ix["Warp Stall Sampling (All Samples)"], ix["Source"],
                               ix["Avg. Threads Executed"])
            lsb = ix.get("stall_long_sb")
            tot = sum(int(r[ex]) for r in data) or 1
            tots = sum(int(r[st]) for r in data) or 1
            long_sb = sum(int(r[lsb]) for r in data) if lsb is not None else 0
            print("## %s\n" % name)
            print("%d SASS lines, %.1f M warp instructions, %d stall samples (%.0f %% long_scoreboard)\n"
                  % (len(data), tot / 1e6, tots, 100.0 * long_sb / tots))
            print("| line | instruction | st | ex | thr |\n|---:|---|---:|---:|---:|")
            top = sorted(range(len(data)), key=lambda k: -int(data[k][st]))[:10]
            for k in sorted(top):
                r = data[k]
                print("| %d | `%s` | %.1f %% | %.2f %% | %s |" % (k, r[src].strip()[:70], 100.0 * int(r[st]) / tots,
                                                                 100.0 * int(r[ex]) / tot, r[th][:4]))
            print()
        else:
            i += 1


if __name__ == "__main__":
    main()
