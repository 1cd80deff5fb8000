"""Host topology of a GPU box as the end-to-end pipeline sees it: NUMA nodes, the CPUs this
process may run on, each GPU's PCI address and NUMA node, and the memory policy calls available.

    python tools/host_topology.py > gpurun_out/host_topology.txt

Read by a person, not by the library: gaast_b200.device.bind_host_near_device() does the binding.
"""
import glob
import os
import subprocess
import sys


def read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError as e:
        return f"<{e.strerror}>"


def main():
    print("allowed cpus:", sorted(os.sched_getaffinity(0)))
    print("online nodes:", read("/sys/devices/system/node/online"),
          " has_memory:", read("/sys/devices/system/node/has_memory"),
          " has_cpu:", read("/sys/devices/system/node/has_cpu"))
    for nd in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        mem = read(nd + "/meminfo").splitlines()
        tot = [l for l in mem if "MemTotal" in l or "MemFree" in l]
        print(os.path.basename(nd), "cpus", read(nd + "/cpulist"), "|", " ; ".join(" ".join(t.split()[2:]) for t in tot))
    print("cgroup cpuset:", read("/sys/fs/cgroup/cpuset.cpus.effective"), "mems:", read("/sys/fs/cgroup/cpuset.mems.effective"))
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id,name", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=60).stdout
    except Exception as e:  # noqa: BLE001
        out = f"<nvidia-smi: {e}>"
    for line in out.strip().splitlines():
        parts = [s.strip() for s in line.split(",")]
        if len(parts) != 3:
            print(line)
            continue
        idx, bdf, name = parts
        short = bdf.lower()
        if len(short.split(":")[0]) == 8:
            short = short[4:]
        print(f"gpu {idx} {name} {bdf} numa_node={read('/sys/bus/pci/devices/' + short + '/numa_node')} "
              f"local_cpulist={read('/sys/bus/pci/devices/' + short + '/local_cpulist')} "
              f"link={read('/sys/bus/pci/devices/' + short + '/current_link_speed')} x{read('/sys/bus/pci/devices/' + short + '/current_link_width')}")
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=60).stdout)
    except Exception as e:  # noqa: BLE001
        print(f"<nvidia-smi topo: {e}>")
    print(subprocess.run("lscpu | head -40", shell=True, capture_output=True, text=True).stdout)
    # is set_mempolicy permitted here?
    import ctypes
    libc = ctypes.CDLL(None, use_errno=True)
    SYS_set_mempolicy, SYS_get_mempolicy = 238, 239  # x86-64
    mode = ctypes.c_int(-1)
    mask = (ctypes.c_ulong * 16)()
    rc = libc.syscall(SYS_get_mempolicy, ctypes.byref(mode), mask, ctypes.c_ulong(1024), None, ctypes.c_ulong(0))
    print("get_mempolicy rc", rc, "errno", ctypes.get_errno(), "mode", mode.value, "mask", hex(mask[0]))
    rc = libc.syscall(SYS_set_mempolicy, ctypes.c_int(0), None, ctypes.c_ulong(0))
    print("set_mempolicy(MPOL_DEFAULT) rc", rc, "errno", ctypes.get_errno())


if __name__ == "__main__":
    sys.exit(main())
