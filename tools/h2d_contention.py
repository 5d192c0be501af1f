"""Host -> device copy bandwidth with every rank copying at once (torchrun, one rank per GPU), with the
process left where the launcher put it and with the process bound to the CPUs of its GPU's NUMA node before the
pinned buffer is allocated.  Explains the gap between the device-timed and the end-to-end bench figures at 8 GPUs."""
import json
import os
import time

import torch
import torch.distributed as dist


def gpu_numa_node(idx):
    bus = torch.cuda.get_device_properties(idx).pci_bus_id if hasattr(torch.cuda.get_device_properties(idx), "pci_bus_id") else None
    import subprocess
    out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(idx)],
                         capture_output=True, text=True).stdout.strip()
    bid = out.lower()
    if bid.startswith("00000000:"):
        bid = bid[4:]
    p = f"/sys/bus/pci/devices/{bid}/numa_node"
    try:
        return int(open(p).read()), bid
    except Exception as e:
        return -1, f"{bid} ({e})"


def node_cpus(node):
    try:
        s = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    except Exception:
        return set()
    cpus = set()
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def measure(dev, mb=142, reps=30):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return mb * (1 << 20) * reps / dt / 1e9


def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    before = sorted(os.sched_getaffinity(0))
    node, bid = gpu_numa_node(lr)
    bw0 = measure(dev)
    allowed = set(before)
    local = node_cpus(node) & allowed if node >= 0 else set()
    bound = False
    if local:
        os.sched_setaffinity(0, local)
        bound = True
    bw1 = measure(dev)
    rec = {"rank": rank, "gpu_bus": bid, "gpu_numa": node, "cpus_allowed": f"{before[0]}..{before[-1]} ({len(before)})",
           "local_cpus": len(local), "bound": bound, "h2d_GBps_unbound": round(bw0, 2), "h2d_GBps_bound": round(bw1, 2)}
    out = [None] * world
    dist.all_gather_object(out, rec)
    if rank == 0:
        for r in out:
            print(json.dumps(r))
        print(json.dumps({"aggregate_unbound": round(sum(r["h2d_GBps_unbound"] for r in out), 1),
                          "aggregate_bound": round(sum(r["h2d_GBps_bound"] for r in out), 1)}))
        try:
            nodes = sorted(n for n in os.listdir("/sys/devices/system/node") if n.startswith("node"))
            print("numa nodes:", {n: open(f"/sys/devices/system/node/{n}/cpulist").read().strip() for n in nodes})
        except Exception as e:
            print("numa:", e)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
