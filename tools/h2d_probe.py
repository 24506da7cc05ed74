"""Diagnostic: steady-state pinned H2D bandwidth vs NUMA node of the pinned pages (verified in numa_maps)."""
import glob, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distributed_rl_b200.hostmem import gpu_node_cpus, _parse_cpulist
dev = torch.device("cuda:0")
torch.cuda.init()
all_cpus = os.sched_getaffinity(0)
nodes = {os.path.basename(n): _parse_cpulist(open(n + "/cpulist").read()) & all_cpus
         for n in sorted(glob.glob("/sys/devices/system/node/node*"))}
print("gpu node cpus (sysfs):", sorted(gpu_node_cpus(dev) or [])[:4], "...")
a = torch.randn(4096, 4096, device=dev)
def where(t):
    addr = "%x" % t.data_ptr()
    try:
        for line in open("/proc/self/numa_maps"):
            f = line.split()
            if f[0] == addr or (int(f[0], 16) <= t.data_ptr() < int(f[0], 16) + (1 << 21) and "N" in line):
                return " ".join(x for x in f if x.startswith("N") and "=" in x) or line.strip()[:80]
    except Exception as e:
        return "numa_maps: %r" % e
    return "?"
def bw(h, it=150, busy=False):
    d = torch.empty(h.numel(), dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        for _ in range(30):
            d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(it):
            d.copy_(h, non_blocking=True)
        e1.record(s)
    if busy:
        for _ in range(400):
            a @ a
    torch.cuda.synchronize()
    return h.numel() * it / (e0.elapsed_time(e1) / 1e3) / 1e9
size = 9 << 20
for name, cpus in list(nodes.items()) + [("all", all_cpus)] + list(nodes.items()):
    if not cpus: continue
    os.sched_setaffinity(0, cpus)
    h = torch.empty(size, dtype=torch.uint8, pin_memory=True); h.fill_(1)   # fresh power-of-two bucket each time
    os.sched_setaffinity(0, all_cpus)
    print(name, "size %d MB" % (size >> 20), "pages:", where(h), "| idle %.1f %.1f busy %.1f GB/s" % (bw(h), bw(h), bw(h, busy=True)), flush=True)
    size *= 2
