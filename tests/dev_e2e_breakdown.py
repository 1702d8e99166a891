"""Developer tool (GPU box): where the end-to-end step loses time against the resident step (bench.py's Case)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "selfsupervised-denoising_b200"))
import torch, bench
from ssdn.params import PipelineOutput
dev = torch.device("cuda", 0)
case = bench.Case("known", 32, dev, 0, 1, True)
for _ in range(4): case.step_resident()
case.capture()
g = case.graphed
def t(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    s = bench.timed(fn, n, dev, False) / n * 1e3
    w = (time.perf_counter() - w0) / n * 1e3
    return s, w
def replay_only(): g.graphs[0].replay()
def call_dev(): g(g.slots[0])
def h2d_only(): g.slots[0][0].copy_(case.host[0], non_blocking=True)
cur = torch.cuda.current_stream(dev)
cs = g.copy_stream
evs = [torch.cuda.Event() for _ in range(4)]
def v_record(): g.graphs[0].replay(); evs[0].record(cur)
def v_wait_done_event():
    evs[1].record(cs)                      # nothing queued on the copy stream: completes at once
    cur.wait_event(evs[1]); g.graphs[0].replay(); evs[0].record(cur)
def v_h2d_same_stream():
    g.slots[0][0].copy_(case.host[0], non_blocking=True); g.graphs[0].replay()
def v_h2d_copy_stream_nowait():
    with torch.cuda.stream(cs):
        g.slots[1][0].copy_(case.host[0], non_blocking=True)      # into the OTHER slot: no dependency at all
    g.graphs[0].replay()
variants = [("graph.replay only", replay_only), ("replay + event record", v_record), ("wait(completed event) + replay + record", v_wait_done_event),
            ("H2D on the compute stream + replay", v_h2d_same_stream), ("H2D on the copy stream (no dependency) + replay", v_h2d_copy_stream_nowait),
            ("call with device tensors", call_dev), ("resident step (bench)", case.step_resident), ("e2e step (bench)", case.step_e2e)]
# round-robin: the chip's clock drifts with temperature / the power cap, so variants are interleaved and the median reported
import statistics
res = {n: [] for n, _ in variants}
for rnd in range(7):
    for n, fn in variants:
        res[n].append(t(fn, 12)[0])
for n, _ in variants:
    print(f"{n:50s} median {statistics.median(res[n]):7.3f} ms/step   min {min(res[n]):7.3f}  max {max(res[n]):7.3f}")
s_, w_ = t(h2d_only)
print(f"H2D copy only (1.5 MB pinned): {s_:.3f} ms")
# CPU time of one call without waiting for the device
torch.cuda.synchronize(); w0 = time.perf_counter(); case.step_e2e(); w1 = time.perf_counter(); torch.cuda.synchronize()
print(f"host time to enqueue one e2e step: {(w1 - w0) * 1e3:.3f} ms")

