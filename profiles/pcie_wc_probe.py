"""H2D rate of an 8.4 MB copy from default pinned memory and from write-combined pinned memory (cudaHostAllocWriteCombined)."""
import ctypes
import torch

torch.cuda.init()
rt = ctypes.CDLL("libcudart.so.12")
nbytes = 65536 * 16 * 8
d = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")


def alloc(flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    ctypes.memset(p, 1, nbytes)
    return p


def rate(p, tag):
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        for _ in range(3):
            rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), p, ctypes.c_size_t(nbytes), 1, ctypes.c_void_p(s.cuda_stream))
        e0.record()
        for _ in range(20):
            rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), p, ctypes.c_size_t(nbytes), 1, ctypes.c_void_p(s.cuda_stream))
        e1.record()
    e1.synchronize()
    print(f"{tag}: {nbytes * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")


for rep in range(2):
    rate(alloc(0), "pinned default       ")
    rate(alloc(4), "pinned write-combined")
h = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
rate(ctypes.c_void_p(h.data_ptr()), "torch pin_memory     ")
