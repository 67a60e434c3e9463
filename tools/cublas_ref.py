"""Reference point for the FP64 tensor pipe: cuBLAS DGEMM / ZGEMM (through torch) on LU-like shapes, TFLOP/s."""
import torch, time
dev = torch.device("cuda", 0)
def bench(f, n_iter=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(n_iter):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for (m, n, k) in [(8192, 8192, 8192), (16384, 16384, 256), (8192, 8192, 256), (30000, 30000, 256), (30000, 30000, 512)]:
    A = torch.randn(m, k, dtype=torch.float64, device=dev); B = torch.randn(k, n, dtype=torch.float64, device=dev); C = torch.randn(m, n, dtype=torch.float64, device=dev)
    ms = bench(lambda: torch.addmm(C, A, B, beta=1.0, alpha=-1.0, out=C))
    print("DGEMM %6d %6d %5d  %8.3f ms  %6.2f TFLOP/s" % (m, n, k, ms, 2.0 * m * n * k / ms / 1e9), flush=True)
    del A, B, C
    if m * n <= 16384 * 16384:
        A = torch.randn(m, k, dtype=torch.complex128, device=dev); B = torch.randn(k, n, dtype=torch.complex128, device=dev); C = torch.randn(m, n, dtype=torch.complex128, device=dev)
        ms = bench(lambda: torch.addmm(C, A, B, beta=1.0, alpha=-1.0, out=C))
        print("ZGEMM %6d %6d %5d  %8.3f ms  %6.2f TFLOP/s (8mnk)" % (m, n, k, ms, 8.0 * m * n * k / ms / 1e9), flush=True)
        del A, B, C
