"""Dev probe (GPU): configs[3] cube compress time under option settings ("name=value,..." per run)."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
A = torch.empty((8320, 64, 64), dtype=torch.complex64, device="cuda:0")
eng.synth_fill(A, 2080, 4)
for combo in sys.argv[1:]:
    opts = dict(kv.split("=") for kv in combo.split(",") if kv)
    for k, v in opts.items():
        eng.set_option(k, float(v))
    for _ in range(3):
        eng.compress(A, compressionrank=8)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.compress(A, compressionrank=8)
        e1.record(); e1.synchronize()
        ts.append(round(e0.elapsed_time(e1) / 10, 3))
    eng.set_option("stage_timing", 1)
    eng.compress(A, compressionrank=8)
    torch.cuda.synchronize()
    em = eng.last_eig_ms()
    eng.set_option("stage_timing", 0)
    print(f"{combo or 'default'}: {ts} ms per compress, leading pairs {em['leading_pairs']:.3f} ms", flush=True)
    for k in opts:
        eng.set_option(k, 0)
