"""BASELINE config 5: SKA-Mid-like decompression-only sweep. Reconstruction of [B, 128, 2048] matrices for k = 1..32 from
random factors with decaying S; reports achieved algorithmic GB/s (Bytes_recon = 8mn + 8k(m+n) + 4k per matrix)
against the measured HBM peak. B is a slice of the 78804-matrix cube that fits comfortably (the full cube is 165 GB of
output and would be ring-buffered)."""
import json
import sys
import torch
sys.path.insert(0, ".")
from visco_b200.engine import get_engine
eng = get_engine(0)
HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if len(sys.argv) < 2 else float(sys.argv[1])
B, m, n = 4096, 128, 2048
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(5)
out = torch.empty((B, m, n), dtype=torch.complex64, device=dev)
rows = []
for k in list(range(1, 9)) + [10, 12, 16, 20, 24, 32]:
    U = torch.view_as_complex(torch.randn((B, m, k, 2), device=dev, generator=g)).contiguous()
    Vt = torch.view_as_complex(torch.randn((B, k, n, 2), device=dev, generator=g)).contiguous()
    S = (torch.rand((B, k), device=dev, generator=g) + 0.1) * torch.exp(-torch.arange(k, device=dev) / 4.0)
    for _ in range(3):
        eng.reconstruct(U, S, Vt, None, out=out)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        eng.reconstruct(U, S, Vt, None, out=out)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    by = B * (8.0 * m * n + 8.0 * k * (m + n) + 4.0 * k)
    ref = (U[:2] * S[:2, None, :]) @ Vt[:2]
    err = float((out[:2] - ref).abs().max() / ref.abs().max())
    rows.append(dict(k=k, ms=ms, gbs=by / ms / 1e6, frac=by / ms / 1e6 / HBM, gvis_s=B * m * n / ms / 1e6,
                     kernel="recon_smallk" if k <= 8 else ("cgemm_tc" if k % 2 == 0 else "cgemm_simt"), err=err))
    print(json.dumps(rows[-1]), flush=True)
