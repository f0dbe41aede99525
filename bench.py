#!/usr/bin/env python
"""bench.py — round trip (compress -> reconstruct) throughput of the VISCO hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload kat7|meerkat|small|ska] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic visibilities: vk_compress_batched
(Gram on tcgen05 -> Jacobi eigensolver -> select/truncate -> factor formation) followed by vk_reconstruct_batched.
Default workload = BASELINE.json configs[1] (KAT-7 shape: 28 baselines x 4 corr x 256 time x 1024 chan complex64,
fixed rank k = 8), one such cube per GPU (weak scaling; no collective on the data path, one NCCL all-gather of the
per-matrix ranks/statistics after the timed region).

Prints ONE JSON line (rank 0). `value` = visibilities compressed+reconstructed per second with the cube resident in
HBM, CUDA-event timed, max over ranks. `e2e` = same metric through the host-buffer C ABI (vk_compress_host /
vk_reconstruct_host) with pinned host inputs and outputs, copies inside the timed region.
`--impl reference` times the reference's CPU path (oracle port of np.linalg.svd + svd_flip + energy rule +
(U*S)@Vt, one process per host core, BLAS threads = 1) on the same workload.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (baselines per GPU, ncorr, m, n, kwargs, BASELINE.json config it is)
    "kat7": (28, 4, 256, 1024, dict(compressionrank=8), "configs[1]: synthetic KAT-7 shape 28 bl x 4 corr x 256 x 1024, k=8"),
    "meerkat": (260, 4, 512, 4096, dict(decorrelation=0.99),
                "configs[2]: MeerKAT-64 shape, 260-baseline shard (1/8 of 2080) x 4 corr x 512 x 4096, decorrelation 0.99"),
    "small": (2080, 4, 64, 64, dict(compressionrank=8), "configs[3]: 2080 bl x 4 corr x 64 x 64, one-sided Jacobi path"),
    "ska": (2048, 4, 128, 2048, dict(compressionrank=8), "configs[4]-like: 2048 bl x 4 corr x 128 x 2048, k=8"),
}
METRIC = "visibilities compressed+reconstructed /sec (GVis/s)"


def algorithmic_bytes(B, m, n, kbar):
    """Bytes_compress = Bytes_recon = 8mn + 8k(m+n) + 4k per matrix."""
    return B * (8.0 * m * n + 8.0 * kbar * (m + n) + 4.0 * kbar)


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        p = {"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops": float(j["bf16_tflops"]),
             "bf16_tflops_sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
             "source": "MEASURED_PEAKS.json"}
    except Exception:
        pass
    return p


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------- CPU reference arm
def _cpu_one(args):
    a, kw = args
    from oracle import visco_oracle as vo
    rec, s, k = vo.roundtrip(a, decorrelation=kw.get("decorrelation"), compressionrank=kw.get("compressionrank"))
    return k


def cpu_worker_main(path):
    """Child process: OPENBLAS/OMP threads were pinned to 1 in the environment BEFORE numpy was imported, and there is
    no CUDA context here, so forking a pool is safe. Prints one JSON line {vis_per_s, seconds}."""
    import multiprocessing as mp
    import numpy as np
    with np.load(path, allow_pickle=True) as z:
        cube, kw, procs, reps = z["cube"], z["kw"].item(), int(z["procs"]), int(z["reps"])
    ctx = mp.get_context("fork")
    secs = []
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_one, [(cube[i], kw) for i in range(min(procs, len(cube)))])  # warm the workers
        for _ in range(reps):
            t0 = time.perf_counter()
            pool.map(_cpu_one, [(cube[i], kw) for i in range(len(cube))], chunksize=1)
            secs.append(time.perf_counter() - t0)
    print(json.dumps({"vis_per_s": cube[0].size * len(cube) * reps / sum(secs), "seconds": sum(secs)}), flush=True)
    return 0


def cpu_roundtrip_rate(cube, kw, procs, reps=1):
    """oracle port on `procs` processes, BLAS threads = 1 each (mirrors the reference's -nw N -nt 1), run in a fresh
    interpreter so that neither this process's CUDA context nor its BLAS thread pool is forked.
    Returns (visibilities per second, seconds)."""
    import subprocess
    import tempfile
    import numpy as np
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "sample.npz")
        np.savez(path, cube=cube, kw=np.array(kw, dtype=object), procs=procs, reps=reps)
        outp = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-worker", path], env=env, check=True,
                              capture_output=True, text=True, timeout=1200).stdout
    j = json.loads(outp.strip().splitlines()[-1])
    return j["vis_per_s"], j["seconds"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle.synth_np import synth_cube
    nbl, ncorr, m, n, kw, desc = WORKLOADS[args.workload]
    procs = len(os.sched_getaffinity(0))
    # bounded sample of the workload: ~10-30 s of CPU work per step
    per_matrix_s = 2.5e-9 * m * n * min(m, n)           # ~0.17 s at 256 x 1024 (survey probe)
    nsample = int(max(min(nbl * ncorr, 20.0 / max(per_matrix_s, 1e-6)), min(nbl * ncorr, procs)))
    nsample = max(ncorr, nsample // ncorr * ncorr)
    cube = synth_cube(nsample // ncorr, ncorr, m, n, nbl_total=nbl * args.gpus)
    if args.warmup:
        cpu_roundtrip_rate(cube[:max(1, min(len(cube), procs))], kw, procs, reps=args.warmup)
    rate, _ = cpu_roundtrip_rate(cube, kw, procs, reps=args.steps)
    value = rate / 1e9
    full_vis = nbl * ncorr * m * n * args.gpus
    line = {
        "metric": METRIC, "value": value, "unit": "GVis/s", "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * full_vis / (value * 1e9),   # time the CPU path needs for the full step workload
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex64", "data": "synthetic",
        "config": {"workload": desc, "shape": [nbl * ncorr * args.gpus, m, n], **kw},
        "cpu_baseline": {"value": value, "unit": "GVis/s", "cores": procs, "kind": "port",
                         "sample": f"{len(cube)} of {nbl * ncorr * args.gpus} matrices per step, one process per core, BLAS threads=1, "
                                   f"one SVD evaluation per matrix (the reference as written does 3-5)"},
        "e2e": {"value": value, "unit": "GVis/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------- our arm
def log(msg):
    if os.environ.get("VISCO_BENCH_VERBOSE"):
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from visco_b200.engine import get_engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    eng = get_engine(local)

    nbl, ncorr, m, n, kw, desc = WORKLOADS[args.workload]
    B = nbl * ncorr
    A = torch.empty((B, m, n), dtype=torch.complex64, device=dev)
    eng.synth_fill(A, nbl, ncorr, bl_offset=rank * nbl, nbl_total=nbl * world)
    kmax = eng.rank_bound(m, n, kw.get("compressionrank"), kw.get("decorrelation"))
    fac = (torch.empty((B, m, kmax), dtype=torch.complex64, device=dev), torch.empty((B, kmax), dtype=torch.float32, device=dev),
           torch.empty((B, kmax, n), dtype=torch.complex64, device=dev), torch.empty((B,), dtype=torch.int32, device=dev),
           torch.empty((B, 4), dtype=torch.float32, device=dev))
    out = torch.empty_like(A)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        U, S, Vt, ranks, stats = eng.compress(A, out=fac, kmax=kmax, **kw)
        eng.reconstruct(U, S, Vt, ranks, out=out)

    log('warmup')
    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    log('timed region')

    # ---- timed region: K steps, CUDA events on the launching stream, stage events recorded inside the library ----
    eng.set_option("stage_timing", 1)
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    stage_acc, eig_acc = {}, {}
    launches0 = eng.launch_count
    barrier()
    ev[0].record()
    for i in range(args.steps):
        U, S, Vt, ranks, stats = eng.compress(A, out=fac, kmax=kmax, **kw)
        ev[2 * i + 1].record()
        eng.reconstruct(U, S, Vt, ranks, out=out)
        ev[2 * i + 2].record()
        for k_, v_ in eng.last_stage_ms().items():
            stage_acc[k_] = stage_acc.get(k_, 0.0) + v_
        for k_, v_ in eng.last_eig_ms().items():
            eig_acc[k_] = eig_acc.get(k_, 0.0) + v_
    barrier()
    launches = eng.launch_count - launches0
    clocks = sampler.stop()
    eng.set_option("stage_timing", 0)
    total_ms = ev[0].elapsed_time(ev[-1])
    comp_ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps)) / args.steps
    recon_ms = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(args.steps)) / args.steps
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    nvis_all = float(B) * m * n * world
    value = nvis_all / (ms_per_step * 1e-3) / 1e9

    log('gather')
    # ---- the only collective: gather per-matrix ranks + statistics (after the timed region) ----
    from visco_b200.shard import gather_ranks_stats
    rk, st = gather_ranks_stats(fac[3], fac[4])
    rk_h, st_h = rk.cpu().numpy(), st.cpu().numpy()
    kbar = float(rk_h.mean())

    log('e2e')
    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()
    # host buffers for the whole cube when they fit in ~8 GB of pinned memory, else a leading sub-batch (stated below)
    per_mat = 8.0 * (2 * m * n + kmax * (m + n))
    Be = int(max(1, min(B, 8e9 // per_mat)))
    Ah = pinned((Be, m, n), torch.complex64)
    Ah[...] = A[:Be].cpu().numpy()
    hout = (pinned((Be, m, kmax), torch.complex64), pinned((Be, kmax), torch.float32), pinned((Be, kmax, n), torch.complex64),
            pinned((Be,), torch.int32), pinned((Be, 4), torch.float32))
    rec_h = pinned((Be, m, n), torch.complex64)
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        Uh, Sh, Vh, rh, _ = eng.compress_host(Ah, out=hout, **kw)
        eng.reconstruct_host(Uh, Sh, Vh, rh, out=rec_h)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    fac_bytes = int(hout[0].nbytes + hout[1].nbytes + hout[2].nbytes + hout[3].nbytes)
    e2e = {"value": float(Be) * m * n * world / e2e_s / 1e9, "unit": "GVis/s", "matrices_per_step_per_gpu": Be,
           "h2d_bytes_per_step": int(Ah.nbytes + fac_bytes), "d2h_bytes_per_step": int(fac_bytes + hout[4].nbytes + rec_h.nbytes),
           "api": "vk_compress_host + vk_reconstruct_host (pinned host buffers)", "steps": e2e_steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- rooflines (rank 0's kernels; algorithmic work per SURVEY section 8d) ----
    pk = peaks()
    r = min(m, n)
    stage_ms = {k_: v_ / args.steps for k_, v_ in stage_acc.items()}
    tf32_peak = 0.5 * pk["bf16_tflops"]
    stages = {}
    if stage_ms.get("gram", 0) > 0:
        fl = 8.0 * r * r * max(m, n) * B
        ach = fl / (stage_ms["gram"] * 1e-3) / 1e12
        stages["gram_tcgen05"] = {"bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                                  "ms": stage_ms["gram"], "note": "algorithmic 8 r^2 n flops (no credit for the 3 TF32 MMAs per product); "
                                  "peak = 0.5 x measured dense bf16 (TF32 rate); stage time includes the Gram normalisation pass"}
    rb = algorithmic_bytes(B, m, n, kbar)
    ach = rb / (recon_ms * 1e-3) / 1e9
    stages["reconstruct"] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "ms": recon_ms}
    fb = algorithmic_bytes(B, m, n, kbar)
    if stage_ms.get("factors", 0) > 0:
        ach = fb / (stage_ms["factors"] * 1e-3) / 1e9
        stages["factor_formation"] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                                      "ms": stage_ms["factors"]}
    jac_ms = stage_ms.get("jacobi", 0.0) + stage_ms.get("small", 0.0)
    sweeps = float(st_h[:, 2].mean())
    eig_ms = {k_: v_ / args.steps for k_, v_ in eig_acc.items()}
    if sum(eig_ms.values()) > 0:
        # direct eigensolver (tridiag.cu). Its kernels, by time; the dominant one is reported as `roofline`.
        #  tridiag_kernel: streams the trailing block once per Householder step (read + write): HBM roofline,
        #      algorithmic bytes = B * sum_j (r-j-1)^2 * 16; served from L2 when B r^2 8 bytes fit there.
        #  leading pairs / QL: scalar, latency bound (one lane per eigenvalue / per matrix): no roofline.
        #  rotations: shared-memory bound (one load + one store of 16 bytes per rotation and column pair).
        # matrices per internal pass of the direct solver (api.cu:auto_chunk: 8 GB of scratch)
        per = 2 * r * r * 8 + (1.5 * r * r + 256) * 8 + (8 * r + 64) * 80 + 2 * 32 * r * 4 + 6 * r * 4
        eig_chunk = int(min(B, max(1, (8 << 30) // per)))
        tb = 16.0 * B * sum((r - j - 1) ** 2 for j in range(max(r - 2, 0)))
        t_ms = eig_ms.get("tridiag", 0.0)
        tri = {"bound": "hbm", "kernel": "tridiag_kernel (Householder tridiagonalisation, fused rank-2 update + matvec)",
               "achieved": tb / (t_ms * 1e-3) / 1e9 if t_ms > 0 else None, "peak": pk["hbm_gbs"], "unit": "GB/s",
               "frac": tb / (t_ms * 1e-3) / 1e9 / pk["hbm_gbs"] if t_ms > 0 else None,
               # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full): the KAT-7 cube as captured;
               # MeerKAT: 181.2 GB for a 296-matrix launch, scaled to the matrices of this launch
               "traffic": {"kat7": 194.4e6, "meerkat": 181.2e9 / 296 * min(B, eig_chunk)}.get(args.workload),
               "traffic_source": {"kat7": "profiles/r01_ncu_full_tridiag_kat7.txt",
                                  "meerkat": "profiles/r01_ncu_full_tridiag_r512.txt"}.get(args.workload),
               "avg_launch_ms": t_ms / max(1.0, math.ceil(B / max(1, eig_chunk))), "ms": t_ms,
               "launches_per_step": float(max(1, math.ceil(B / max(1, eig_chunk)))),
               "share_of_step": t_ms / ms_per_step if ms_per_step else None,
               "algorithmic_bytes_per_step": tb,
               "note": "algorithmic bytes = trailing block read + written once per Householder step; "
                       + ("the %d matrices of a pass (%.0f MB) stay in the 126 MB L2, so DRAM traffic is far below it and the "
                          "figure is an L2-bandwidth one" % (min(B, eig_chunk), min(B, eig_chunk) * r * r * 8 / 1e6)
                          if min(B, eig_chunk) * r * r * 8 < 126e6 else
                          "the matrices of a pass (%.0f MB) exceed L2: HBM-bound" % (min(B, eig_chunk) * r * r * 8 / 1e6)),
               "peak_source": pk["source"]}
        stages["eig_tridiag"] = tri
        for k_, label in (("leading_pairs", "bisect/twisted/backtr kernels (leading eigenpairs, fixed rank)"),
                          ("ql", "tql_kernel (implicit QL, one lane per matrix, latency bound)"),
                          ("reflectors", "formq_kernel (reflector accumulation, rows in registers, fp32 SIMT)"),
                          ("rotations", "rotapply_kernel (level-scheduled plane rotations, shared-memory bound)")):
            if eig_ms.get(k_, 0.0) > 0:
                stages["eig_" + k_] = {"kernel": label, "ms": eig_ms[k_], "share_of_step": eig_ms[k_] / ms_per_step}
        top = max(eig_ms, key=lambda q: eig_ms[q])
        if top == "tridiag":
            dominant = tri
        else:
            dominant = {"bound": "hbm", "kernel": stages["eig_" + top]["kernel"], "achieved": None, "peak": pk["hbm_gbs"],
                        "unit": "GB/s", "frac": None, "traffic": None, "ms": eig_ms[top],
                        "share_of_step": eig_ms[top] / ms_per_step,
                        "note": "not an HBM- or tensor-bound kernel (see roofline_stages.eig_tridiag for the HBM-bound one); "
                                "time only", "peak_source": pk["source"]}
    else:
        # the Jacobi rotation kernels. FP32 SIMT + shared memory: neither contract roofline bounds it
        # (SURVEY 8d); reported against HBM with its algorithmic traffic (every launch reads and writes the r x r vectors).
        if eng.uses_small_path(m, n):
            L = max(m, n) + r
            jbytes = 2.0 * B * r * L * 8
            nlaunch = 1.0
        else:
            jbytes = 2.0 * B * r * r * 8
            nb = 2 if r <= 64 else ((r + 15) // 16 + ((r + 15) // 16) % 2)
            nlaunch = max(1.0, sweeps * nb)
        dom_ms = jac_ms / nlaunch if jac_ms > 0 else 0.0
        dominant = {"bound": "hbm", "kernel": "jacobi (one-sided cyclic Jacobi rotations, fp32 SIMT)",
                    "achieved": (jbytes / (dom_ms * 1e-3) / 1e9) if dom_ms > 0 else None, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": (jbytes / (dom_ms * 1e-3) / 1e9 / pk["hbm_gbs"]) if dom_ms > 0 else None,
                    "traffic": None, "avg_launch_ms": dom_ms, "launches_per_step": nlaunch,
                    "share_of_step": jac_ms / ms_per_step if ms_per_step else None,
                    "modelled_gflop_per_step": 22.0 * r ** 3 * sweeps * B / 1e9 if not eng.uses_small_path(m, n) else None,
                    "note": "latency/issue-bound fp32 rotations on L2-resident data; no HBM or tensor roofline applies, frac is informational",
                    "peak_source": pk["source"]}

    log('cpu baseline')
    # ---- CPU baseline on the host cores: oracle port on a bounded sample of THIS cube (bit-identical inputs) ----
    procs = len(os.sched_getaffinity(0))
    per_matrix_s = 2.5e-9 * m * n * r
    nsample = int(max(min(B, 20.0 / max(per_matrix_s, 1e-6)), min(B, procs)))
    idx = np.linspace(0, B - 1, nsample).astype(int)
    sample = A[torch.as_tensor(idx, device=dev)].cpu().numpy()
    cpu_rate, cpu_s = cpu_roundtrip_rate(sample, kw, procs)
    # parity spot check of the benchmarked result against the oracle on three matrices of the sample
    from oracle import visco_oracle as vo
    Uh, Sh, Vh = fac[0].cpu().numpy(), fac[1].cpu().numpy(), fac[2].cpu().numpy()
    s_err = 0.0
    for b in idx[:3]:
        k = int(rk_h[b])
        u, s, vt = vo.ref_apply_svd(sample[list(idx).index(b)], kw.get("decorrelation"), kw.get("compressionrank"))
        kk = min(k, len(s))
        s_err = max(s_err, float(np.max(np.abs(Sh[b, :kk] - s[:kk]) / s[:kk])))
    cpu_baseline = {"value": cpu_rate / 1e9, "unit": "GVis/s", "cores": procs, "kind": "port",
                    "sample": f"{nsample} of {B} matrices of this cube (D2H copies, bit-identical inputs), one process per core, "
                              f"BLAS threads=1, one SVD evaluation per matrix; {cpu_s:.1f} s",
                    "sigma_max_rel_err_vs_oracle": s_err}

    line = {
        "metric": METRIC, "value": value, "unit": "GVis/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex64 (fp32 arithmetic; Gram as 3xTF32 on tcgen05 with fp32 accumulation)", "data": "synthetic",
        "config": {"workload": desc, "shape_per_gpu": [B, m, n], **kw,
                   "l2": f"inputs {A.numel() * 8 / 1e6:.0f} MB per GPU " + ("> 126 MB L2 (no flush needed)" if A.numel() * 8 > 126e6 else "< L2"),
                   "mean_rank": kbar, "mean_sweeps": sweeps, "converged": bool(st_h[:, 3].min() == 1)},
        "compress_ms": comp_ms, "reconstruct_ms": recon_ms, "stage_ms": stage_ms,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": dominant, "roofline_stages": stages, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="kat7", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-worker", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_worker:
        return cpu_worker_main(args.cpu_worker)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun, one rank per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--workload", args.workload]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
