"""One windowed-attention volume split into slabs over the ranks (non-overlapping windows, no exchange):
forward + backward time of the whole volume as the max over ranks (strong scaling: the volume is fixed).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/bench_slab.py
        [--size 256] [--w 5] [--stride 5] [--pad 3] [--d 64] [--reps 5]       (G = 1: plain python works too)
Every rank generates its own slab (synthetic randn): the point is the kernel time of the slab geometry, and
that the plan tiles the volume.  Timing: CUDA events, barrier + synchronize on both sides, max over ranks."""
import argparse, json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))
import fa_sm100a as fa

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--w", type=int, default=5)
ap.add_argument("--stride", type=int, default=5)
ap.add_argument("--pad", type=int, default=3)
ap.add_argument("--d", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
bf = torch.bfloat16
spatial = (a.size,) * 3
plan = fa.windowed_slab_plan(spatial, a.w, a.stride, a.pad, rank, world)
planes = plan.plane_hi - plan.plane_lo
q, k, v, g = (fa.jl_empty((a.size, a.size, planes, a.d, 1), bf, dev).normal_() for _ in range(4))


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


tf, (y, l, m) = timed(lambda: fa.windowed_fa_slab(q, k, v, a.w, plan, stride=a.stride, pad=a.pad))
tb, _ = timed(lambda: fa.windowed_fa_slab_backward(q, k, v, g, l, m, a.w, plan, stride=a.stride, pad=a.pad))
cover = torch.tensor([planes, plan.nwin], device=dev, dtype=torch.int64)
if world > 1:
    dist.all_reduce(cover)
if rank == 0:
    ntok = a.size ** 3
    nw = fa.window_counts(spatial, a.w, a.stride, a.pad)
    byf, byb = 4 * ntok * a.d * 2, 7 * ntok * a.d * 2
    print(json.dumps({"ranks": world, "volume": list(spatial), "W": a.w, "stride": a.stride, "pad": a.pad, "d": a.d, "dtype": "bf16",
                      "planes_covered": int(cover[0]), "window_planes_covered": int(cover[1]), "window_planes": nw[-1],
                      "fwd_ms": tf, "bwd_ms": tb, "fwd_alg_gbs_total": byf / tf / 1e6, "bwd_alg_gbs_total": byb / tb / 1e6,
                      "tokens_per_s_fwd": ntok / tf * 1e3, "path": fa.last_path()}), flush=True)
if world > 1:
    dist.destroy_process_group()
