"""Development tool: one reconstruction launch at the BASELINE configs[4] shape for ncu (usage: prof_recon.py B k)."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine

eng = get_engine(0)
B, k = int(sys.argv[1]), int(sys.argv[2])
m, n = 128, 2048
U = torch.randn((B, m, k), dtype=torch.complex64, device="cuda:0") / (2 * m) ** 0.5
Vt = torch.randn((B, k, n), dtype=torch.complex64, device="cuda:0") / (2 * n) ** 0.5
S = torch.rand((B, k), dtype=torch.float32, device="cuda:0") + 0.5
out = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
for _ in range(3):
    eng.reconstruct(U, S, Vt, None, out=out)
torch.cuda.synchronize()
print("ok")
