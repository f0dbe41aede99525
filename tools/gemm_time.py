"""Dev probe (GPU): large-rank reconstruction (cgemm_tc) time at the MeerKAT shape, k = 416, 296 matrices."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
B, m, n, k = 296, 512, 4096, 416
g = torch.Generator(device="cuda:0").manual_seed(1)
U = torch.view_as_complex(torch.randn((B, m, k, 2), device="cuda:0", generator=g)).contiguous()
Vt = torch.view_as_complex(torch.randn((B, k, n, 2), device="cuda:0", generator=g)).contiguous()
S = torch.rand((B, k), device="cuda:0", generator=g)
out = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
eng.reconstruct(U, S, Vt, None, out=out)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.reconstruct(U, S, Vt, None, out=out)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 3
ref = (U[:2] * S[:2, None, :]) @ Vt[:2]
err = float((out[:2] - ref).abs().max() / ref.abs().max())
print(f"{ms:.3f} ms  {8.0 * B * m * n * k / ms / 1e9:.0f} TFLOP/s algorithmic  max rel diff vs torch {err:.2e}", flush=True)
