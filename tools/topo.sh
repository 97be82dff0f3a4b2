nvidia-smi topo -m 2>&1 | head -30
python - <<'PY'
import os
print("affinity", sorted(os.sched_getaffinity(0)))
import glob
for n in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")): print(n, open(n).read().strip())
import torch
for i in range(torch.cuda.device_count()):
    p=torch.cuda.get_device_properties(i)
    bid=f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    try: nn=open(f"/sys/bus/pci/devices/{bid}/numa_node").read().strip()
    except Exception as e: nn=str(e)
    print(i,bid,nn)
PY
