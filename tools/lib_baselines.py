"""Library fp64 reference points on the GPU box (cuBLAS DGEMM, cuSOLVER potrf/potri via torch). Context only:
these size the fp64 roofline denominator and give a vendor-library time for the same N^3 flops."""
import torch, time, json, sys
def t(f, n=3):
    f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(n):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
out = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device='cuda'); b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    ms = t(lambda: a @ b); out[f"dgemm_{n}_tflops"] = 2 * n**3 / ms * 1e-9
    ms = t(lambda: a @ b.T); out[f"dgemm_nt_{n}_tflops"] = 2 * n**3 / ms * 1e-9
    del a, b
# sustained: loop 2s
a = torch.randn(8192, 8192, dtype=torch.float64, device='cuda'); b = torch.randn(8192, 8192, dtype=torch.float64, device='cuda')
torch.cuda.synchronize(); t0 = time.time(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); k = 0
while k < 60: a @ b; k += 1
e1.record(); torch.cuda.synchronize(); out["dgemm_8192_sustained_tflops"] = k * 2 * 8192**3 / e0.elapsed_time(e1) * 1e-9
del a, b
for n in (4096, 16384):
    x = torch.randn(n, 8, dtype=torch.float64, device='cuda')
    K = torch.exp(-0.5 * torch.cdist(x, x)**2 / 8) + 0.01 * torch.eye(n, dtype=torch.float64, device='cuda')
    ms = t(lambda: torch.linalg.cholesky(K), 2); out[f"cusolver_potrf_{n}_ms"] = ms; out[f"cusolver_potrf_{n}_tflops"] = n**3 / 3 / ms * 1e-9
    L = torch.linalg.cholesky(K)
    ms = t(lambda: torch.cholesky_inverse(L), 2); out[f"cusolver_potri_{n}_ms"] = ms
    del K, L
print(json.dumps(out, indent=1))
