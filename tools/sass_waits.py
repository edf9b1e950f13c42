"""Scoreboard view of a kernel's SASS: which loads share a scoreboard, and where each scoreboard is waited on.

    python tools/sass_waits.py kanpyo_b200/libkanpyo_b200.so kp_viterbiILi8 [first_line last_line]

Decodes the control bits of the 128-bit instruction words printed by `cuobjdump -sass` (stall count, write /
read barrier index, wait mask: the layout NVIDIA GPUs have used since Volta).  ptxas has six scoreboards per
warp; loads that share one are waited for together, and an instruction that re-arms a scoreboard or overwrites a
register last written through it waits for it to drain.  A load whose scoreboard is waited on a few
instructions after it was issued has its whole latency exposed, whatever the source code suggests
(profiles/r01i_sweep.md is the case that paid)."""
import re
import subprocess
import sys


def decode(lib, fn):
    raw = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
    out, on, i = [], False, 0
    while i < len(raw):
        ln = raw[i]
        if "Function :" in ln:
            on = fn in ln
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", ln)
        if on and m and i + 1 < len(raw):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", raw[i + 1])
            if m2:
                ctrl = (int(m2.group(1), 16) >> 41) & 0x1FFFFF
                out.append({"addr": m.group(1), "text": m.group(2).strip(), "stall": ctrl & 0xF, "wb": (ctrl >> 5) & 7,
                            "rb": (ctrl >> 8) & 7, "wait": (ctrl >> 11) & 0x3F})
                i += 2
                continue
        i += 1
    return out


def main():
    lib, fn = sys.argv[1], sys.argv[2]
    ins = decode(lib, fn)
    lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4]) if len(sys.argv) > 4 else len(ins)
    print("%d instructions in %s" % (len(ins), fn))
    for k, x in enumerate(ins[lo:hi], lo):
        mem = any(t in x["text"] for t in ("LDG", "STG", "LDS", "STS", "ATOM", "RED", "LDL", "STL"))
        if mem or (x["wait"] & 0b111100):
            # next instruction that waits on this load's scoreboard
            nxt = ""
            if mem and x["wb"] != 7:
                for j in range(k + 1, min(k + 400, len(ins))):
                    if ins[j]["wait"] >> x["wb"] & 1:
                        nxt = "  -> first wait on SB%d: +%d (%s)" % (x["wb"], j - k, ins[j]["text"][:40])
                        break
            print("%4d %s %-56s wb:%s rb:%s wait:%s%s" % (k, x["addr"], x["text"][:56], x["wb"] if x["wb"] != 7 else "-",
                                                          x["rb"] if x["rb"] != 7 else "-",
                                                          format(x["wait"], "06b") if x["wait"] else "------", nxt))


if __name__ == "__main__":
    main()
