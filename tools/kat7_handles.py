"""Dev probe (GPU): KAT-7 cubes per second with several concurrent handles (one host thread + stream each, as bench.py's
main loop) under library options. Usage: kat7_handles.py "nh:name=value,..." ..."""
import sys
import threading
import torch
sys.path.insert(0, ".")
from visco_b200.engine import Engine, get_engine

dev = torch.device("cuda:0")
eng0 = get_engine(0)
B, m, n, k = 112, 256, 1024, 8
cubes = []
for c in range(4):
    A = torch.empty((B, m, n), dtype=torch.complex64, device=dev)
    eng0.synth_fill(A, 28, 4, bl_offset=c * 28, nbl_total=28 * 4)
    cubes.append(A)
for spec in sys.argv[1:]:
    nh, _, optstr = spec.partition(":")
    nh = int(nh)
    engines = [Engine(0) for _ in range(nh)]
    for e in engines:
        for kv in filter(None, optstr.split(",")):
            e.set_option(kv.split("=")[0], float(kv.split("=")[1]))
    streams = [torch.cuda.Stream(device=dev) for _ in range(nh)]
    outs = [torch.empty_like(cubes[0]) for _ in range(nh)]

    def run(ncubes):
        def work(hd):
            with torch.cuda.stream(streams[hd]):
                for i in range(hd, ncubes, nh):
                    U, S, Vt, ranks, stats = engines[hd].compress(cubes[i % 4], compressionrank=k)
                    engines[hd].reconstruct(U, S, Vt, ranks, out=outs[hd])
        ths = [threading.Thread(target=work, args=(hd,)) for hd in range(nh)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
    run(4 * nh)
    torch.cuda.synchronize()
    res = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        N = 240
        run(N)
        e1.record(); e1.synchronize()
        res.append(round(e0.elapsed_time(e1) / N, 4))
    print(f"{spec}: ms per cube {res}", flush=True)
    for e in engines:
        e.close()
