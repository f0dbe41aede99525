"""Development tool: one compress of B matrices m x n (fixed rank 8) for ncu captures of the tridiagonalisation kernel."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
B, m, n = [int(x) for x in sys.argv[1:4]]
A = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
eng.synth_fill(A, B // 4, 4)
eng.set_option("eig_impl", 2)
if len(sys.argv) > 4:
    eng.set_option("tridiag_impl", int(sys.argv[4]))
for kv in sys.argv[5:]:                                   # further library options as name=value
    eng.set_option(kv.split("=")[0], float(kv.split("=")[1]))
for _ in range(2):
    eng.compress(A, compressionrank=8)
torch.cuda.synchronize()
print("ok")
