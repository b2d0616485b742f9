"""H2D bandwidth vs NUMA placement of the pinned buffer."""
import os, time, subprocess, torch, numpy as np
n = 132 * 1024 * 1024
def bw(h, d, reps=10):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return n / ((time.perf_counter() - t) / reps) / 1e9
d = torch.empty(n, dtype=torch.uint8, device="cuda")
bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", "0"], stdout=subprocess.PIPE, text=True).stdout.strip().lower()
bus = bus[-12:] if len(bus) > 12 else bus
try:
    node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
except Exception as e:
    node = -1
print("gpu bus", bus, "numa node", node, "affinity size", len(os.sched_getaffinity(0)))
print(subprocess.run("lscpu | grep -i numa", shell=True, stdout=subprocess.PIPE, text=True).stdout)
h1 = torch.from_numpy(np.ones(n, np.uint8)).pin_memory()
print("default placement:", ["%.1f" % bw(h1, d) for _ in range(4)])
for nd in (0, 1):
    try:
        cpus = open("/sys/devices/system/node/node%d/cpulist" % nd).read().strip()
        s = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            s |= set(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, s)
        h = torch.from_numpy(np.ones(n, np.uint8)).pin_memory()
        print("pinned buffer allocated while bound to node", nd, ":", ["%.1f" % bw(h, d) for _ in range(4)])
    except Exception as e:
        print("node", nd, "failed", e)
