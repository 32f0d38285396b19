"""Pinned H2D bandwidth: default pinned memory vs write-combined (cudaHostAllocWriteCombined), 632 MB per copy, 1..N GPUs in one process."""
import ctypes as C, sys, time
import torch
rt = C.CDLL("libcudart.so.12")
n = 632_000_000
ngpu = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for flags, name in ((0, "default"), (4, "write-combined")):
    bufs, devs, streams = [], [], []
    for g in range(ngpu):
        torch.cuda.set_device(g)
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
        C.memset(p, 65, n)
        bufs.append(p)
        devs.append(torch.empty(n, dtype=torch.uint8, device="cuda:%d" % g))
        streams.append(torch.cuda.Stream(device=g))
    def run(reps):
        for _ in range(reps):
            for g in range(ngpu):
                torch.cuda.set_device(g)
                assert rt.cudaMemcpyAsync(C.c_void_p(devs[g].data_ptr()), bufs[g], C.c_size_t(n), C.c_int(1), C.c_void_p(streams[g].cuda_stream)) == 0
        for g in range(ngpu):
            torch.cuda.set_device(g); torch.cuda.synchronize()
    run(2)
    t0 = time.perf_counter(); run(10); dt = time.perf_counter() - t0
    print("%s pinned, %d GPU(s): %.1f GB/s aggregate H2D" % (name, ngpu, 10 * ngpu * n / dt / 1e9))
    for p in bufs: rt.cudaFreeHost(p)
