"""Dev probe (GPU): per-kernel times of the fixed-rank (leading eigenpairs) path through vk_compress_batched."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine

eng = get_engine(0)
a = [int(x) for x in sys.argv[1:]] or [112, 256, 1024, 8]
for i in range(0, len(a), 4):
    B, m, n, k = a[i:i + 4]
    A = torch.empty((B, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, B // 4, 4)
    eng.compress(A, compressionrank=k)
    eng.set_option("stage_timing", 2)
    U, S, Vt, ranks, stats = eng.compress(A, compressionrank=k)
    torch.cuda.synchronize()
    eng.set_option("stage_timing", 0)
    print("sweeps==0 (leading-pairs path) for", int((stats[:, 2] == 0).sum()), "of", B, "matrices; done", int(stats[:, 3].sum()))
