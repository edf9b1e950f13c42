#!/bin/bash
nvidia-smi topo -m 2>/dev/null | head -14
python - <<'PY'
import os
try:
    import pynvml as nv
    nv.nvmlInit()
    for i in range(nv.nvmlDeviceGetCount()):
        h = nv.nvmlDeviceGetHandleByIndex(i)
        try:
            print(i, "affinity", [hex(int(w)) for w in nv.nvmlDeviceGetCpuAffinity(h, 2)])
        except Exception as e:
            print(i, "affinity failed", e)
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        p = "/sys/bus/pci/devices/%s/numa_node" % (bus[4:] if len(bus.split(":")[0]) == 8 else bus)
        print("  ", bus, p, open(p).read().strip() if os.path.exists(p) else "missing")
except Exception as e:
    print("nvml failed", e)
print("allowed cpus", len(os.sched_getaffinity(0)))
PY
