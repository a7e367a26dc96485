"""development: copy-engine rates for the shapes gcmesh_mesh_host uses (strided 2-D H2D of brick runs, strip-sized D2H), alone, together,
and with host threads reading memory beside them."""
import os, sys, time, threading, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
cudart = ctypes.CDLL("libcudart.so.12")
MB = 1 << 20
cols, strips = 72, 18
src = torch.ones(cols * strips * 4 * 131072, dtype=torch.uint8).pin_memory()       # "blocks" bricks of 18 strips of 72 columns (x4 bricks each)
dst = torch.empty_like(src, device="cuda")
out_d = torch.ones(175 * MB, dtype=torch.uint8, device="cuda"); out_h = torch.empty(175 * MB, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def h2d_2d():
    for k in range(strips):
        off = k * cols * 4 * 131072
        e = cudart.cudaMemcpy2DAsync(ctypes.c_void_p(dst.data_ptr() + off), ctypes.c_size_t(4 * 131072), ctypes.c_void_p(src.data_ptr() + off), ctypes.c_size_t(4 * 131072),
                                     ctypes.c_size_t(131072), ctypes.c_size_t(cols), 1, ctypes.c_void_p(s1.cuda_stream))
        assert e == 0
def h2d_lin():
    for k in range(strips):
        off = k * cols * 4 * 131072
        e = cudart.cudaMemcpyAsync(ctypes.c_void_p(dst.data_ptr() + off), ctypes.c_void_p(src.data_ptr() + off), ctypes.c_size_t(cols * 131072), 1, ctypes.c_void_p(s1.cuda_stream))
        assert e == 0
def d2h():
    n = out_d.numel() // strips
    with torch.cuda.stream(s2):
        for k in range(strips): out_h[k * n:(k + 1) * n].copy_(out_d[k * n:(k + 1) * n], non_blocking=True)
def timed(f, k=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k
h2d_bytes = strips * cols * 131072
print("H2D strided 2-D (18 x 72 rows of 128 KB, pitch 512 KB): %.1f GB/s" % (h2d_bytes / timed(h2d_2d) / 1e9))
print("H2D linear pieces of %.1f MB: %.1f GB/s" % (cols * 131072 / 1e6, h2d_bytes / timed(h2d_lin) / 1e9))
print("D2H in 18 pieces: %.1f GB/s" % (out_d.numel() / timed(d2h) / 1e9))
t = timed(lambda: (h2d_2d(), d2h())); print("2-D H2D + D2H together: %.2f ms -> H2D %.1f GB/s, D2H %.1f GB/s" % (t * 1e3, h2d_bytes / t / 1e9, out_d.numel() / t / 1e9))
t = timed(lambda: (h2d_lin(), d2h())); print("linear H2D + D2H together: %.2f ms -> H2D %.1f GB/s, D2H %.1f GB/s" % (t * 1e3, h2d_bytes / t / 1e9, out_d.numel() / t / 1e9))
# host readers beside the copies (numpy sum releases the GIL)
big = np.ones(1 << 30, np.uint8)
stop = False
def reader(i):
    part = big[i * (1 << 26):(i + 1) * (1 << 26)]
    while not stop: part.max()
for nthreads in (4, 8, 15):
    stop = False
    th = [threading.Thread(target=reader, args=(i,)) for i in range(nthreads)]
    [x.start() for x in th]
    time.sleep(0.2)
    t = timed(lambda: (h2d_2d(), d2h()), 3)
    stop = True; [x.join() for x in th]
    print("with %d host threads reading memory: %.2f ms -> H2D %.1f GB/s, D2H %.1f GB/s" % (nthreads, t * 1e3, h2d_bytes / t / 1e9, out_d.numel() / t / 1e9))
