import os, torch, glob
p = torch.cuda.get_device_properties(0)
bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
print("bdf", bdf)
for f in ("numa_node", "local_cpulist"):
    try:
        print(f, open(f"/sys/bus/pci/devices/{bdf}/{f}").read().strip())
    except Exception as e:
        print(f, "ERR", e)
print("affinity", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], "...")
print("nodes", [(os.path.basename(n), open(n + "/cpulist").read().strip()) for n in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))])
os.system("nvidia-smi topo -m | head -8")
