"""One compress + reconstruct of a few MeerKAT-shaped matrices (for ncu captures of gram_tc / cgemm_tc kernels)."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 74
A = torch.empty((B, 512, 4096), dtype=torch.complex64, device="cuda:0")
eng.synth_fill(A, B, 1, nbl_total=64)
W = eng.gram(A, impl=2)
torch.cuda.synchronize()
g = torch.Generator(device="cuda:0").manual_seed(1)
k = 256
U = torch.view_as_complex(torch.randn((B, 512, k, 2), device="cuda:0", generator=g)).contiguous()
Vt = torch.view_as_complex(torch.randn((B, k, 4096, 2), device="cuda:0", generator=g)).contiguous()
S = torch.rand((B, k), device="cuda:0", generator=g)
out = eng.reconstruct(U, S, Vt, None)
U8, Vt8, S8 = U[:, :, :8].contiguous(), Vt[:, :8].contiguous(), S[:, :8].contiguous()
out8 = eng.reconstruct(U8, S8, Vt8, None)
torch.cuda.synchronize()
print("ok")
