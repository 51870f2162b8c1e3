"""Print the NUMA layout of the box and whether this process may set its memory policy."""
import ctypes, glob, os, subprocess
print("nodes:", sorted(os.path.basename(p) for p in glob.glob("/sys/devices/system/node/node[0-9]*")))
for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*/cpulist")):
    print(p, open(p).read().strip())
print("affinity:", sorted(os.sched_getaffinity(0)))
try:
    out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        idx, bdf = [v.strip() for v in line.split(",")]
        bdf = bdf.lower()[4:] if len(bdf) > 12 else bdf.lower()
        path = f"/sys/bus/pci/devices/{bdf}/numa_node"
        print(idx, bdf, open(path).read().strip() if os.path.exists(path) else "no sysfs entry")
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
except Exception as e:
    print("nvidia-smi:", e)
libc = ctypes.CDLL(None, use_errno=True)
mask = ctypes.c_ulong(1)
rc = libc.syscall(238, 1, ctypes.byref(mask), 64)      # set_mempolicy(MPOL_PREFERRED, {node 0})
print("set_mempolicy rc", rc, "errno", ctypes.get_errno())
