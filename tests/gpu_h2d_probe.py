"""Pinned host -> device copy bandwidth, with and without binding to the GPU's NUMA node."""
import os, sys, glob, subprocess, time, torch
dev = torch.device("cuda:0")
def bw(label):
    h = torch.empty(400 << 20, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty_like(h, device=dev)
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(label, "H2D GB/s: %.1f" % (5 * h.numel() / e0.elapsed_time(e1) / 1e6), flush=True)
print("cpus allowed:", len(os.sched_getaffinity(0)), "of", os.cpu_count())
bw("default")
try:
    bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
except Exception:
    bus = None
out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
print("bus", out)
node_path = "/sys/bus/pci/devices/%s/numa_node" % out.lower().replace("00000000:", "0000:")
try:
    node = int(open(node_path).read())
except Exception as e:
    node = -1
    print("numa lookup failed", e)
print("numa node of gpu0:", node, "nodes:", sorted(glob.glob("/sys/devices/system/node/node*")))
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
if node >= 0:
    cl = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
    cpus = set()
    for part in cl.split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    cpus &= os.sched_getaffinity(0)
    if cpus:
        os.sched_setaffinity(0, cpus)
        bw("bound to node %d (%d cpus)" % (node, len(cpus)))
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout)
