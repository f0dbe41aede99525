"""Dev probe run on the GPU box: stage timings of the hot path on a few shapes (not a benchmark)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from visco_b200.engine import get_engine  # noqa: E402


def run(eng, nbl, ncorr, m, n, reps=3, **kw):
    A = torch.empty((nbl * ncorr, m, n), dtype=torch.complex64, device="cuda:0")
    eng.synth_fill(A, nbl, ncorr, nbl_total=max(nbl, 64))
    torch.cuda.synchronize()
    res = {}
    eng.set_option("stage_timing", 1)
    for _ in range(reps):
        U, S, Vt, ranks, stats = eng.compress(A, **kw)
        torch.cuda.synchronize()
        res = eng.last_stage_ms()
    eng.set_option("stage_timing", 0)
    eng.reconstruct(U, S, Vt, ranks)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t2 = torch.cuda.Event(enable_timing=True)
    t0.record()
    U, S, Vt, ranks, stats = eng.compress(A, **kw)
    t1.record()
    out = eng.reconstruct(U, S, Vt, ranks)
    t2.record()
    torch.cuda.synchronize()
    st = stats.cpu()
    nvis = A.numel()
    err = float((A - out).abs().pow(2).sum().sqrt() / A.abs().pow(2).sum().sqrt())
    line = dict(shape=[nbl * ncorr, m, n], kw=kw, stages_ms=res, compress_ms=t0.elapsed_time(t1),
                recon_ms=t1.elapsed_time(t2), gvis_s=nvis / (t0.elapsed_time(t2) * 1e-3) / 1e9,
                rank_mean=float(ranks.float().mean()), rank_max=int(ranks.max()), sweeps_mean=float(st[:, 2].mean()),
                sweeps_max=float(st[:, 2].max()), converged=float(st[:, 3].min()), rel_err=err,
                recon_gbs=(out.numel() * 8 + U.numel() * 8 + Vt.numel() * 8) / (t1.elapsed_time(t2) * 1e-3) / 1e9)
    print(json.dumps(line), flush=True)
    return line


if __name__ == "__main__":
    eng = get_engine(0)
    import os
    for kv in filter(None, os.environ.get("VISCO_OPTS", "").split(",")):
        k_, v_ = kv.split("=")
        eng.set_option(k_, float(v_))
    print(torch.cuda.get_device_name(0), os.environ.get("VISCO_OPTS", ""), flush=True)
    which = sys.argv[1:] or ["c2", "c4", "c3s", "c1"]
    t = time.time()
    if "c2" in which:
        run(eng, 28, 4, 256, 1024, compressionrank=8)
    if "c4" in which:
        run(eng, 2080, 4, 64, 64, compressionrank=8)
        run(eng, 2080, 4, 64, 64, decorrelation=0.99)
    if "c3s" in which:
        run(eng, 8, 4, 512, 4096, reps=1, decorrelation=0.99)
    if "c3m" in which:
        run(eng, 64, 4, 512, 4096, reps=1, decorrelation=0.99)
    if "c1" in which:
        run(eng, 21, 4, 360, 16, decorrelation=0.9)
    if "c5" in which:
        run(eng, 256, 4, 128, 2048, compressionrank=8)
    print("probe wall s", time.time() - t, flush=True)
