"""Dev probe (GPU): eigensolver kernel times (stage_timing = 2 prints them on stderr) for the tridiagonalisation variants
("tridiag_impl" 0 = lower triangle + deferred, 1 = undeferred, 2 = full-storage deferred). Usage: tridiag_ab.py B m n [k] ..."""
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
    eng.set_option("eig_impl", 2)
    ref = None
    for impl, var in ((0, -1), (0, 0), (0, 32), (0, 64), (0, 96), (0, 128)):
        eng.set_option("tridiag_impl", impl)
        eng.set_option("tridiag_nts", var)
        eng.set_option("stage_timing", 2)
        kw = dict(compressionrank=k) if k > 0 else dict(decorrelation=0.99)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        res = eng.compress(A, **kw)
        e1.record()
        torch.cuda.synchronize()
        S = res[1] if isinstance(res, (tuple, list)) else res.S
        if ref is None:
            ref = S.clone()
        dev = float(((S - ref).abs() / ref.abs().clamp_min(1e-20)).max())
        print(f"B={B} {m}x{n} k={k} tridiag_impl={impl} variant={var}: compress {e0.elapsed_time(e1):.3f} ms, max rel dS vs impl 1 {dev:.2e}", flush=True)
    eng.set_option("stage_timing", 0)
    eng.set_option("tridiag_impl", 0)
    eng.set_option("tridiag_nts", -1)
