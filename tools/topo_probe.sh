nvidia-smi topo -m 2>&1 | head -30
echo ---- numa
ls /sys/devices/system/node/ 2>&1 | head
for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist); done
echo ---- gpus
python - <<'PY'
import torch, os
for i in range(torch.cuda.device_count()):
    p = torch.cuda.get_device_properties(i)
    bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
    try:
        node = open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip()
    except Exception as e:
        node = "err %s" % e
    print(i, bus, "numa", node)
print("affinity", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8])
PY
nproc; free -g | head -2
