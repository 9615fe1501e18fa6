import torch, time
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2): c = a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    best = 1e9
    for _ in range(5):
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print("cuBLAS DGEMM", n, "%.2f ms  %.2f TFLOP/s" % (best, 2 * n**3 / best / 1e9))
    L = torch.linalg.cholesky(a @ a.T + n * torch.eye(n, dtype=torch.float64, device="cuda"))
    m = a @ a.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    e0.record(); L = torch.linalg.cholesky(m); e1.record(); torch.cuda.synchronize()
    print("cusolver potrf", n, "%.2f ms  %.2f TFLOP/s" % (e0.elapsed_time(e1), n**3 / 3 / e0.elapsed_time(e1) / 1e9))
