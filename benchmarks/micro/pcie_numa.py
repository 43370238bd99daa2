"""Developer tool: pinned-host <-> device copy bandwidth as a function of the NUMA node the pinned pages live on
(set_mempolicy(MPOL_PREFERRED) before the allocation).  bench.py's end-to-end arm is PCIe-bound (326 MB each way per
step), so where the pinned buffers live decides its number.

    python benchmarks/micro/pcie_numa.py
"""
import ctypes
import glob
import os
import subprocess

import torch

libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy = 238          # x86_64
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2


def set_mempolicy(mode, node=None):
    if node is None:
        return libc.syscall(SYS_set_mempolicy, mode, None, 0)
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(mask), 64)


def bw(nbytes=1 << 28, reps=5):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = {}
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        out[name] = round(nbytes * reps / a.elapsed_time(b) / 1e6, 1)
    # both directions at once on two streams
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a); s2.wait_event(a)
    for _ in range(reps):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    out["both_each"] = round(nbytes * reps / a.elapsed_time(b) / 1e6, 1)
    return out


def main():
    print("cpus allowed:", sorted(os.sched_getaffinity(0)))
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    print("numa nodes:", nodes)
    for n in nodes:
        try:
            print(f" node{n} cpulist:", open(f"/sys/devices/system/node/node{n}/cpulist").read().strip())
        except OSError as e:
            print(" ", e)
    try:
        print("mems allowed:", [l for l in open("/proc/self/status") if l.startswith(("Mems_allowed_list", "Cpus_allowed_list"))])
    except OSError:
        pass
    props = torch.cuda.get_device_properties(0)
    bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
    try:
        print("gpu0", bus, "numa_node:", open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
    except OSError as e:
        print("gpu0 numa_node unreadable:", e)
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[:1500])
    except Exception as e:  # noqa: BLE001
        print("topo failed", e)
    torch.cuda.init()
    print("default policy GB/s:", bw())
    for n in nodes:
        rc = set_mempolicy(MPOL_PREFERRED, n)
        print(f"preferred node {n} (rc {rc}) GB/s:", bw())
    set_mempolicy(MPOL_DEFAULT)


if __name__ == "__main__":
    main()
