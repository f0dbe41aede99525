"""Dev probe (GPU): compress time of the 1040-matrix MeerKAT shard under option settings given as name=value arguments
(comma separated per run; options named in a run are set back to `reset` values afterwards)."""
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
A = torch.empty((1040, 512, 4096), dtype=torch.complex64, device="cuda:0")
eng.synth_fill(A, 260, 4)
RESET = {"split_variant": 1}
for combo in sys.argv[1:]:
    opts = dict(kv.split("=") for kv in combo.split(",") if kv)
    for k, v in opts.items():
        eng.set_option(k, float(v))
    for _ in range(2):
        eng.compress(A, decorrelation=0.99)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        eng.compress(A, decorrelation=0.99)
    e1.record(); e1.synchronize()
    print(combo, round(e0.elapsed_time(e1) / 4, 2), "ms per compress", flush=True)
    for k in opts:
        eng.set_option(k, RESET.get(k, 0))
