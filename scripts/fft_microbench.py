"""Time the transform kernels alone at the RB3D 256^3 shapes (CUDA events, inputs > L2)."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from dedalus_b200.transforms import RealFourierTransform, FastChebyshevTransform

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
G = 3 * N // 2
reps = 5


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
rf = RealFourierTransform(G, N)
ch0 = FastChebyshevTransform(G, N, -0.5, -0.5, -0.5, -0.5, stretch=0.5)
ch2 = FastChebyshevTransform(G, N, 1.5, 1.5, -0.5, -0.5, stretch=0.5)
cases = [
    ("z_bwd_plain", ch0, 'b', (N, N, N), (N, N, G), 2, 0),
    ("z_bwd_deriv", ch0, 'b', (N, N, N), (N, N, G), 2, 1),
    ("z_fwd_conv", ch2, 'f', (N, N, G), (N, N, N), 2, 0),
    ("y_bwd", rf, 'b', (N, N, G), (N, G, G), 1, 0),
    ("y_fwd", rf, 'f', (N, G, G), (N, N, G), 1, 0),
    ("x_bwd", rf, 'b', (N, G, G), (G, G, G), 0, 0),
    ("x_bwd_deriv", rf, 'b', (N, G, G), (G, G, G), 0, 1),
    ("x_fwd", rf, 'f', (G, G, G), (N, G, G), 0, 0),
]
for name, plan, d, sin, sout, axis, deriv in cases:
    a = torch.randn(sin, dtype=torch.float64, device='cuda'); b = torch.empty(sout, dtype=torch.float64, device='cuda')
    if d == 'b':
        fn = lambda: plan.backward(a, b, axis, deriv=deriv)
    else:
        fn = lambda: plan.forward(a, b, axis)
    ms = timeit(fn)
    gb = 8 * (a.numel() + b.numel()) / 1e9
    out[name] = dict(ms=ms, gbps=gb / (ms * 1e-3))
    print(f"{name:14s} {ms:8.3f} ms  {gb / (ms * 1e-3):8.1f} GB/s", flush=True)
a = torch.randn((G, G, G), dtype=torch.float64, device='cuda'); b = torch.empty_like(a)
ms = timeit(lambda: b.copy_(a)); print(f"{'copy':14s} {ms:8.3f} ms  {16 * a.numel() / 1e9 / (ms * 1e-3):8.1f} GB/s")
json.dump(out, open("gpurun_out/fft_microbench.json", "w"), indent=1)
