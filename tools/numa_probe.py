#!/usr/bin/env python
"""Host topology of the GPU box, for the e2e (host-buffer) path: which cores / memory nodes this
process may use, which NUMA node every GPU hangs off, and the pinned H2D rate per GPU.

    python tools/numa_probe.py [--bw]
"""
import glob
import json
import os
import subprocess
import sys


def read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except Exception as exc:  # noqa: BLE001
        return "n/a (%s)" % type(exc).__name__


def main():
    out = {"nproc_affinity": len(os.sched_getaffinity(0)),
           "affinity": sorted(os.sched_getaffinity(0)),
           "cpu_count": os.cpu_count()}
    status = read("/proc/self/status")
    for line in status.splitlines():
        if line.startswith(("Cpus_allowed_list", "Mems_allowed_list")):
            k, v = line.split(":", 1)
            out[k] = v.strip()
    out["nodes"] = {os.path.basename(p): {"cpulist": read(p + "/cpulist"),
                                          "meminfo": read(p + "/meminfo").splitlines()[:2]}
                    for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))}
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id,name", "--format=csv,noheader"],
                           stdout=subprocess.PIPE, text=True, timeout=30).stdout
        gpus = []
        for line in q.strip().splitlines():
            idx, bus, name = [x.strip() for x in line.split(",")]
            bdf = bus.lower()
            if len(bdf.split(":")[0]) == 8:
                bdf = bdf[4:]
            gpus.append({"index": int(idx), "bus": bus, "name": name,
                         "numa_node": read("/sys/bus/pci/devices/%s/numa_node" % bdf),
                         "local_cpulist": read("/sys/bus/pci/devices/%s/local_cpulist" % bdf)})
        out["gpus"] = gpus
        out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, text=True,
                                     timeout=30).stdout.splitlines()
    except Exception as exc:  # noqa: BLE001
        out["gpus"] = "nvidia-smi failed: %s" % exc
    print(json.dumps(out, indent=1))
    if "--bw" in sys.argv:
        import time
        import torch
        n = 168 << 20
        for dev in range(torch.cuda.device_count()):
            torch.cuda.set_device(dev)
            h = torch.empty(n, dtype=torch.uint8).pin_memory()
            d = torch.empty(n, dtype=torch.uint8, device="cuda")
            for _ in range(2):
                d.copy_(h, non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                d.copy_(h, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
            print(json.dumps({"gpu": dev, "h2d_gbs": n / dt / 1e9}))


if __name__ == "__main__":
    main()
