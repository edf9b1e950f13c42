"""Turn the artefacts of tools/gpu_final.sh (gpurun_out/*_<tag>.*) into the committed evidence under profiles/.
    python tools/make_profiles.py r01c "one-line description of the code state" """
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, note = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# 1. launch list: one step
lines = [l for l in open(os.path.join(G, "launches_%s.csv" % tag)) if l.startswith('"')]
r = list(csv.reader(lines))
hdr, rows = r[0], r[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ks = [(re.sub(r"\(.*", "", x[ki]).replace("void ", ""), float(x[vi].replace(",", ""))) for x in rows]
ks = [k for k in ks if k[0].startswith("kp_")]
idx = [i for i, k in enumerate(ks) if k[0] == "kp_prep_count"]
step = ks[idx[-2]:idx[-1]]
tot = sum(k[1] for k in step)
with open(os.path.join(P, "%s_launches_cfg2.md" % tag), "w") as f:
    f.write("# %s — ncu launch list of one cfg2 step (65 536 sentences, 16.1 MB)\n\n%s\n\n" % (tag, note))
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 1 "
            "--warmup 3 --no-cpu` (per-launch times are cold-cache and serialised: compare shares, not absolutes)\n\n")
    f.write("| kernel | us | share |\n|---|---:|---:|\n")
    for k in step:
        f.write("| %s | %.1f | %.1f%% |\n" % (k[0], k[1] / 1e3, 100 * k[1] / tot))
    f.write("| **total** | %.1f | |\n" % (tot / 1e3))
shutil.copy(os.path.join(G, "launches_%s.csv" % tag), os.path.join(P, "%s_launches_cfg2.csv" % tag))

# 2. full capture summary
rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep,
                     "%s — ncu --set full, cfg2 step. %s" % (tag, note)], capture_output=True, text=True).stdout
open(os.path.join(P, "%s_ncu_full.md" % tag), "w").write(md)

# 3. DRAM traffic of the dominant kernel, per launch
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u = rr[0], rr[1]
traffic = {}
for row in rr[2:]:
    name = row[h.index("Kernel Name")].split("(")[0].replace("void ", "")

    def val(key):
        i = h.index(key)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[i]]
        return float(row[i]) * scale
    traffic[name] = {"dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
                     "source": "profiles/%s_ncu_full.md (dram__bytes_read.sum + dram__bytes_write.sum)" % tag}
json.dump(traffic, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)

# 4. bench lines
for a, b in (("bench_%s.json", "%s_bench.json"), ("bench_ref_%s.json", "%s_bench_reference.json")):
    if os.path.exists(os.path.join(G, a % tag)):
        shutil.copy(os.path.join(G, a % tag), os.path.join(P, b % tag))
for a in ("sanitize_%s.log", "pytest_gpu_%s.log", "smoke_%s.log"):
    if os.path.exists(os.path.join(G, a % tag)):
        shutil.copy(os.path.join(G, a % tag), os.path.join(P, "%s_%s" % (tag, (a % tag).replace("_%s" % tag, ""))))
print(open(os.path.join(P, "%s_launches_cfg2.md" % tag)).read())
print(json.dumps(traffic, indent=1))
