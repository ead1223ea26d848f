"""GPU-box probe: where does the e2e (host-input) step spend its time?"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from merlot_reserve_b200 import MerlotReserve, _lib
from merlot_reserve_b200.synthetic import load_config, synthetic_batch, synthetic_params

cfg = load_config("base")
B = 64
params = synthetic_params(cfg, 0)
model = MerlotReserve.from_config(cfg).bind(params)
batch = synthetic_batch(cfg, B, nseg=16, seed=1)
host = {k: torch.from_numpy(v).pin_memory() for k, v in batch.items()}
res = {k: v.cuda() for k, v in host.items()}
out_host = torch.empty((B, 160, 768), dtype=torch.float32).pin_memory()


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


with torch.no_grad():
    dst = torch.empty_like(res["images"])
    print("raw H2D images ms", timeit(lambda: dst.copy_(host["images"], non_blocking=True)))
    print("resident ms", timeit(lambda: model.batch_embed_video(**res)))

    def e2e():
        out = model.batch_embed_video(**host)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for chunks in (1, 2, 4, 8):
        model._H2D_CHUNKS = chunks
        print("e2e chunks", chunks, timeit(e2e))
    # host enqueue time of one resident step (CPU-bound launch path?)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.batch_embed_video(**res)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("enqueue ms", (t1 - t0) * 1e3, "total", (time.perf_counter() - t0) * 1e3)
