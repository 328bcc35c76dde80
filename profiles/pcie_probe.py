import os, subprocess, time, torch, ctypes
print(subprocess.run("nvidia-smi topo -m; nproc; lscpu | grep -i -E 'numa|socket|model name'; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c", shell=True, capture_output=True, text=True).stdout)
print("affinity now:", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], "...")
def bw(tag, nbytes, wc=False):
    n = nbytes // 8
    if wc:
        rt = ctypes.CDLL("libcudart.so.12") if False else None
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    h.fill_(1.0)
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    s = torch.cuda.Stream()
    for direction in ("H2D", "D2H"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            for _ in range(3):
                (d.copy_(h, non_blocking=True) if direction == "H2D" else h.copy_(d, non_blocking=True))
            e0.record()
            for _ in range(10):
                (d.copy_(h, non_blocking=True) if direction == "H2D" else h.copy_(d, non_blocking=True))
            e1.record()
        e1.synchronize()
        print(f"{tag} {direction} {nbytes/1e6:8.1f} MB: {nbytes*10/(e0.elapsed_time(e1)*1e-3)/1e9:6.1f} GB/s")
torch.cuda.init()
for nb in (8 << 20, 256 << 20):
    bw("default-affinity", nb)
import pynvml
pynvml.nvmlInit()
hdl = pynvml.nvmlDeviceGetHandleByIndex(0)
try:
    words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (os.cpu_count() + 63) // 64)
    cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
    print("GPU0 ideal cpus:", len(cpus), cpus[:8], "...")
    os.sched_setaffinity(0, set(cpus) & os.sched_getaffinity(0) or os.sched_getaffinity(0))
    print("affinity set:", len(os.sched_getaffinity(0)))
    for nb in (8 << 20, 256 << 20):
        bw("gpu-local-affinity", nb)
except Exception as e:
    print("affinity probe failed:", e)
