"""Summarise an .ncu-rep (ncu --set full) into a markdown table for profiles/.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep "title" > profiles/rNN_x.md"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_bytes.sum", "L2 bytes"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__waves_per_multiprocessor", "waves / SM"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %")]


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    print("# %s\n" % title)
    print("Source: `%s` (`ncu --set full --clock-control none --import-source on`; values per launch)\n" % rep)
    for row in rows:
        name = row[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        print("## %s\n" % name)
        print("| metric | value |\n|---|---|")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print("| %s (`%s`) | %s %s |" % (label, key, row[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(row[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("| top stalls (warps per issue) | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in stalls[:5]))
        print()


if __name__ == "__main__":
    main()
