"""Context for the e2e number: the box's PCIe link one way and both ways at once (pinned 1 GiB buffers), the GPU's NUMA node
and this process's CPU affinity."""
import os, time, torch
dev = torch.device('cuda:0')
nb = 1 << 30
ha = torch.empty(nb, dtype=torch.uint8).pin_memory(); hb = torch.empty(nb, dtype=torch.uint8).pin_memory()
da = torch.empty(nb, dtype=torch.uint8, device=dev); db = torch.empty(nb, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=4):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): da.copy_(ha, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): hb.copy_(db, non_blocking=True)
    torch.cuda.synchronize(); return reps * nb / (time.perf_counter() - t) / 1e9
for _ in range(2): run(True, True, 1)
print(f"h2d alone {run(True, False):.1f} GB/s, d2h alone {run(False, True):.1f} GB/s, both at once {run(True, True):.1f} GB/s each way")
try:
    import pynvml
    pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
    bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
    bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
    p = f"/sys/bus/pci/devices/{bdf[-12:].lower()}/numa_node"
    print("gpu", bdf, "numa_node", open(p).read().strip() if os.path.exists(p) else "n/a", "| nodes", sorted(d for d in os.listdir('/sys/devices/system/node') if d.startswith('node')) if os.path.isdir('/sys/devices/system/node') else "n/a")
except Exception as e:  # noqa: BLE001
    print("nvml:", e)
print("affinity", len(os.sched_getaffinity(0)), "cpus of", os.cpu_count())
