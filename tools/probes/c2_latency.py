"""Graph-timed forward / backward of BASELINE config 2 (64x64 image, 7x7 windows, d = 64, batch 8, bf16): kernel latency
of a launch with at most one iteration per resident CTA.  Environment knobs are read once per process."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa
label = sys.argv[1] if len(sys.argv) > 1 else ""
bf = torch.bfloat16
def graph_time(fn, reps):
    fn(); torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return best
q, k, v, g = (fa.jl_empty((64, 64, 64, 8), bf).normal_() for _ in range(4))
y, l, m = fa.windowed_fa(q, k, v, 7, 7, 3)
tf = graph_time(lambda: fa.windowed_fa(q, k, v, 7, 7, 3), 20)
tb = graph_time(lambda: fa.windowed_fa_backward(q, k, v, g, l, m, 7, 7, 3), 20)
print(f"{label:20s} C2 fwd {tf*1e3:.2f} us  bwd {tb*1e3:.2f} us", flush=True)
