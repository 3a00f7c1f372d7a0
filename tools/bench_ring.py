"""Ring attention (one long sequence sharded by tokens over the ranks) forward + backward timing.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/bench_ring.py
        [--nl 16384] [--d 128] [--heads 8] [--reps 5]
Each rank holds Nl tokens; the sequence has G*Nl.  Work per rank: 4*Nl*(G*Nl)*d*B flop forward, 10*... backward
(algorithmic, softmax excluded).  Roofline of the fused compute+exchange step (B200_PROFILING.md): the slower of
flop / measured GEMM peak and bytes over NVLink / 770 GB/s; prints both and the achieved fraction.
Timing: CUDA events on the launch stream, barrier + synchronize on both sides, max over ranks."""
import argparse, json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa

ap = argparse.ArgumentParser()
ap.add_argument("--nl", type=int, default=16384)
ap.add_argument("--d", type=int, default=128)
ap.add_argument("--heads", type=int, default=8)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
bf = torch.bfloat16
q, k, v, g = (fa.jl_empty((a.nl, a.d, a.heads), bf, dev).normal_() for _ in range(4))


def timed(fn):
    for _ in range(2):
        out = fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out = fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.reps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), out


tf, (O, l, m) = timed(lambda: fa.ring_dense_fa(q, k, v))
tb, _ = timed(lambda: fa.ring_dense_fa_backward(q, k, v, O, g, l, m))
if rank == 0:
    ntot = world * a.nl
    ff, fb = 4.0 * a.nl * ntot * a.d * a.heads, 10.0 * a.nl * ntot * a.d * a.heads
    kv_bytes = 2 * a.nl * a.d * a.heads * 2 * (world - 1)                      # K+V blocks sent per rank, forward
    acc_bytes = kv_bytes + 2 * a.nl * a.d * a.heads * 4 * world                # + fp32 dK/dV accumulators, backward
    t_flop_f, t_link_f = ff / (peaks["bf16_tflops"] * 1e9), kv_bytes / 770e6   # ms
    t_flop_b, t_link_b = fb / (peaks["bf16_tflops"] * 1e9), acc_bytes / 770e6
    print(json.dumps({"ranks": world, "N_total": ntot, "N_local": a.nl, "d": a.d, "heads": a.heads, "dtype": "bf16",
                      "fwd_ms": tf, "fwd_tflops_per_gpu": ff / tf / 1e9, "fwd_roofline_ms": max(t_flop_f, t_link_f),
                      "fwd_frac_of_roofline": max(t_flop_f, t_link_f) / tf, "fwd_nvlink_ms_if_exposed": t_link_f,
                      "bwd_ms": tb, "bwd_tflops_per_gpu": fb / tb / 1e9, "bwd_roofline_ms": max(t_flop_b, t_link_b),
                      "bwd_frac_of_roofline": max(t_flop_b, t_link_b) / tb, "bwd_nvlink_ms_if_exposed": t_link_b}), flush=True)
if world > 1:
    dist.destroy_process_group()
