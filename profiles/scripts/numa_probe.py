"""Host topology of the GPU box and H2D bandwidth of pinned memory allocated on each NUMA node."""
import glob, os, subprocess, sys
import torch

print(subprocess.run("lscpu | grep -i 'numa\\|socket\\|model name\\|^CPU(s)'; nvidia-smi topo -m | head -14", shell=True,
                     capture_output=True, text=True).stdout)
prop = torch.cuda.get_device_properties(0)
bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
path = "/sys/bus/pci/devices/%s/numa_node" % bus
print("gpu0 pci", bus, "numa_node", open(path).read().strip() if os.path.exists(path) else "n/a")
print("affinity now: %d cpus" % len(os.sched_getaffinity(0)))
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
allcpus = os.sched_getaffinity(0)


def cpulist(s):
    out = set()
    for part in s.strip().split(","):
        if "-" in part:
            a, b = part.split("-"); out.update(range(int(a), int(b) + 1))
        elif part:
            out.add(int(part))
    return out


dst = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for nd in nodes:
    cpus = cpulist(open(nd + "/cpulist").read()) & allcpus
    if not cpus:
        continue
    os.sched_setaffinity(0, cpus)
    h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    os.sched_setaffinity(0, allcpus)
    for rep in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dst.copy_(h, non_blocking=True); b.record(); torch.cuda.synchronize()
    print("%s (%d cpus): pinned H2D %.1f GB/s" % (os.path.basename(nd), len(cpus), (256 << 20) / a.elapsed_time(b) / 1e6), flush=True)
    a.record(); h.copy_(dst, non_blocking=True); b.record(); torch.cuda.synchronize()
    print("%s: pinned D2H %.1f GB/s" % (os.path.basename(nd), (256 << 20) / a.elapsed_time(b) / 1e6), flush=True)
    del h
