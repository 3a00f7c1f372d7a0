#!/usr/bin/env python
"""bench.py -- headline benchmark of the FlashAttention.jl hot path on B200.

Workload (BASELINE.json configs[2], the config the metric's target is quoted on):
  dense_fa forward, bf16, seq N=8192, head dim d=128, batch*heads B=16*32=512 per GPU.
A "step" is one dense_fa forward over that batch.  One rank per GPU; the path shards over the
trailing batch dim with no collective, so N GPUs = N independent shards ("weak").

  python bench.py --gpus 1 --steps 10 --warmup 3            # our CUDA path
  python bench.py --impl reference --steps 2 --warmup 1     # CPU restatement of the reference
  torchrun ... bench.py --gpus N ...                        # one rank per GPU

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launch stream, barrier +
synchronize on both sides, max over ranks; inputs (3 GiB) exceed L2 (126 MB).

Besides `value` the line carries: `e2e` (the same config through fa_dense_fwd_host on page-locked HOST buffers, H2D +
kernel + D2H inside the timed region, B = 512 per GPU), `roofline`, `cpu_baseline`, and `extra` = the other legs of the
named configs (dense backward, C4 circulant, C5 / C2 windowed) plus -- for N > 1 -- the two legs that put real bytes
on NVLink / split one problem over the GPUs: `ring_dense` (ONE N*16384-token sequence sharded by tokens,
K/V blocks round the ring by ncclSend/ncclRecv, forward + backward, with a parity check against the single-GPU
kernel on rank 0) and `slab_windowed` (ONE 256^3 volume cut into slabs on window boundaries).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flashattention.jl_b200"))

N, D, B_PER_GPU = 8192, 128, 512
WORKLOAD = "dense_fa fwd bf16 N=8192 d=128 B=512(16x32 heads) per GPU [BASELINE configs[2]]"
FLOPS_PER_BATCH_ELT = 4.0 * N * N * D            # 4 N^2 d (softmax flops excluded), BASELINE.md section 3


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"bf16_tflops": j.get("bf16_tflops"), "bf16_tflops_sustained": j.get("bf16_tflops_sustained"),
                "hbm_gbs": j.get("hbm_gbs"), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.stop_flag, self.th = gpu_index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_reference_tflops(steps, warmup, batch_sample):
    """CPU restatement of the reference (oracle/fa_oracle.c, C + OpenMP; the Julia reference cannot
    run here) on a bounded sample of the workload: Float32, same N and d, `batch_sample` batch
    elements, all host cores as the reference's `@threads` would use (src/dense.jl:45)."""
    import numpy as np
    from oracle import c_oracle as co
    cores = co.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    rng = np.random.default_rng(0)
    q, k, v = (np.asfortranarray(rng.standard_normal((N, D, batch_sample), dtype=np.float32)) for _ in range(3))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        co.dense_fa(q, k, v)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return FLOPS_PER_BATCH_ELT * batch_sample / t / 1e12, t, cores


def run_reference(args):
    """`--impl reference`: the reference's own CPU path (restated) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_b = 8
    val, t, cores = cpu_reference_tflops(args.steps, args.warmup, sample_b)
    line = {
        "impl": "reference", "metric": "dense_fa forward attention TFLOP/s (4*N^2*d*B / time)", "value": val,
        "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic randn (seeded)",
        "config": {"workload": WORKLOAD, "sample": f"B={sample_b} of {B_PER_GPU} batch elements per step, Float32"},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                         "sample": f"oracle/fa_oracle.c dense_fa (C+OpenMP restatement of src/dense.jl:21-102), N={N} d={D} B={sample_b}, "
                                   f"{cores} OpenMP threads over (batch,row-block) tasks"},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "tokens_per_s": sample_b * N / t,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="batch*heads per GPU (default: the named config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the backward / circulant / windowed legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import fa_sm100a as fa

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Bn = args.batch
    warmup = max(args.warmup, 3)

    # synthetic inputs of the named shape, generated on the host in Float32 and cast
    bf = torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)

    def make(seed_shift):
        t = fa.jl_empty((N, D, Bn), bf, dev)
        base = t.permute(2, 1, 0)                     # C-contiguous (B, d, N)
        for b0 in range(0, Bn, 64):                   # chunked: bounded host memory
            nb = min(64, Bn - b0)
            base[b0:b0 + nb].copy_(torch.randn(nb, D, N, generator=g, dtype=torch.float32).to(bf))
        return t

    q, k, v = make(0), make(1), make(2)
    O = fa.jl_empty((N, D, Bn), bf, dev)
    l = fa.jl_empty((N, 1, Bn), torch.float32, dev)
    m = fa.jl_empty((N, 1, Bn), torch.float32, dev)

    def step():
        fa.dense_fa_(O, l, m, q, k, v)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    assert fa.last_path() == "tc", "headline config must run on the tcgen05 path"
    kernel_name = "tc_fwd_kernel<128,bf16> (csrc/fa_tc_fwd.cu)"
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    total_flops = FLOPS_PER_BATCH_ELT * Bn * world
    value = total_flops / (ms * 1e-3) / 1e12
    per_gpu = value / world

    # ---- the other legs of the named config and of the sharded configs (not part of `value`):
    #      dense backward at the same shape, circulant C4, windowed C5 and C2 forward/backward per GPU
    extra = None
    if not args.no_extra:
        def timeit(fn, reps):
            for _ in range(3):
                fn()
            barrier()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b_.record()
            barrier()
            t = a.elapsed_time(b_) / reps
            if world > 1:
                tt = torch.tensor([t], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t = float(tt.item())
            return t
        peaks_x = measured_peaks()
        extra = {}
        # dense backward, same tensors (dO = randn), deterministic two-kernel tcgen05 backward
        Bb = Bn                                             # the whole named batch (workspace: 4 GiB of fp16 re-encodings)
        dO = fa.jl_empty((N, D, Bb), bf, dev).normal_()
        sl = lambda t: fa.jl_array(t.permute(2, 1, 0)[:Bb].permute(2, 1, 0))
        qb, kb, vb, Ob, lb, mb = (sl(t) for t in (q, k, v, O, l, m))
        gb = dO
        tb = timeit(lambda: fa.dense_fa_backward(qb, kb, vb, Ob, gb, lb, mb), max(2, args.steps // 3))
        fl = 10.0 * N * N * D * Bb
        extra["dense_bwd"] = {"ms": tb, "tflops_per_gpu": fl / tb / 1e9, "frac_tensor_peak": fl / tb / 1e9 / peaks_x["bf16_tflops"],
                              "batch": Bb, "path": fa.last_path(), "flops": "10*N^2*d*B (7 GEMMs executed: deterministic dQ pass recomputes S, dP)"}
        del dO, qb, kb, vb, Ob, gb, lb, mb
        torch.cuda.empty_cache()
        # C4: circulant 1-D forward, N=16384, W=255, d=64, B=512 total sharded over the ranks
        Bc = 512 // world
        cq, ck, cv = (fa.jl_empty((16384, 64, Bc), bf, dev).normal_() for _ in range(3))
        cO = fa.jl_empty((16384, 64, Bc), bf, dev); cl = fa.jl_empty((16384, 1, Bc), torch.float32, dev); cm = fa.jl_empty((16384, 1, Bc), torch.float32, dev)
        tc_ = timeit(lambda: fa.circulant_fa_(cO, cl, cm, cq, ck, cv, 255), max(args.steps, 30))     # one pass, 30+ calls
        by = (4 * 16384 * 64 * 2 + 8 * 16384) * Bc
        extra["C4_circulant_fwd"] = {"ms": tc_, "batch_per_gpu": Bc, "tflops_per_gpu": 4.0 * 16384 * 255 * 64 * Bc / tc_ / 1e9,
                                     "alg_gbs_per_gpu": by / tc_ / 1e6, "frac_hbm_peak": by / tc_ / 1e6 / peaks_x["hbm_gbs"],
                                     "tokens_per_s": Bc * 16384 * world / tc_ * 1e3, "path": fa.last_path()}
        del cq, ck, cv, cO, cl, cm
        torch.cuda.empty_cache()
        # C5: windowed 3-D forward + backward, 64^3 volume, W=5 (stride 5, pad 3), d=64, B=64 total sharded
        Bw = max(1, 64 // world)
        wq, wk, wv, wg = (fa.jl_empty((64, 64, 64, 64, Bw), bf, dev).normal_() for _ in range(4))
        wy, wl, wm = fa.windowed_fa(wq, wk, wv, 5, 5, 3)
        tf_ = timeit(lambda: fa.windowed_fa(wq, wk, wv, 5, 5, 3), max(2, args.steps // 2))
        pf = fa.last_path()
        tw_ = timeit(lambda: fa.windowed_fa_backward(wq, wk, wv, wg, wl, wm, 5, 5, 3), max(2, args.steps // 2))
        ntok = 64 ** 3
        byf = (4 * ntok * 64 * 2) * Bw + 8 * 125 * 2744 * Bw
        byb = (7 * ntok * 64 * 2) * Bw + 8 * 125 * 2744 * Bw
        extra["C5_windowed3d"] = {"batch_per_gpu": Bw, "fwd_ms": tf_, "bwd_ms": tw_,
                                  "fwd_alg_gbs_per_gpu": byf / tf_ / 1e6, "bwd_alg_gbs_per_gpu": byb / tw_ / 1e6,
                                  "fwd_frac_hbm_peak": byf / tf_ / 1e6 / peaks_x["hbm_gbs"], "bwd_frac_hbm_peak": byb / tw_ / 1e6 / peaks_x["hbm_gbs"],
                                  "tokens_per_s_fwd": Bw * ntok * world / tf_ * 1e3, "path": pf + "/" + fa.last_path()}
        del wq, wk, wv, wg, wy, wl, wm
        torch.cuda.empty_cache()
        # C2: windowed 2-D forward + backward, 64x64 image, W=7 (stride 7, pad 3), d=64, B=8 (a 17 MB problem: one wave of
        # CTAs, so the calls are captured into a CUDA graph -- GPU time without the Python wrapper / launch latency)
        def graph_time(fn, reps):
            fn(); torch.cuda.synchronize()
            st_ = torch.cuda.Stream()
            with torch.cuda.stream(st_):
                fn()
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_, stream=st_):
                    for _ in range(reps):
                        fn()
            torch.cuda.synchronize()
            g_.replay(); torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g_.replay(); b_.record(); torch.cuda.synchronize()
            return a.elapsed_time(b_) / reps
        iq, ik, iv, ig = (fa.jl_empty((64, 64, 64, 8), bf, dev).normal_() for _ in range(4))
        iy, il, im = fa.windowed_fa(iq, ik, iv, 7, 7, 3)
        t2f = graph_time(lambda: fa.windowed_fa(iq, ik, iv, 7, 7, 3), 20)
        p2 = fa.last_path()
        t2b = graph_time(lambda: fa.windowed_fa_backward(iq, ik, iv, ig, il, im, 7, 7, 3), 20)
        by2f = (4 * 4096 * 64 * 2) * 8 + 8 * 49 * 100 * 8
        by2b = (7 * 4096 * 64 * 2) * 8 + 8 * 49 * 100 * 8
        extra["C2_windowed2d"] = {"batch": 8, "fwd_ms": t2f, "bwd_ms": t2b, "timing": "CUDA graph of 20 calls, per call",
                                  "fwd_alg_gbs": by2f / t2f / 1e6, "bwd_alg_gbs": by2b / t2b / 1e6,
                                  "fwd_frac_hbm_peak": by2f / t2f / 1e6 / peaks_x["hbm_gbs"], "bwd_frac_hbm_peak": by2b / t2b / 1e6 / peaks_x["hbm_gbs"],
                                  "path": p2 + "/" + fa.last_path()}
        del iq, ik, iv, ig, iy, il, im
        torch.cuda.empty_cache()

        # ---- N > 1 only: the two legs with a real multi-GPU data plane
        if world > 1:
            # (1) ring attention: ONE sequence of world*16384 tokens, d = 128, 8 heads, bf16, sharded by tokens; K/V blocks
            #     travel round the ring over NVLink (ncclSend/ncclRecv inside fa_ring_dense_fwd/_bwd), overlapped with the
            #     tcgen05 kernels.  Per-GPU rate, fraction of the same kernel on ONE block without any exchange, and parity
            #     of rank 0's slice against the single-GPU kernel on the whole (gathered) sequence.
            Nl, Hh = 16384, 8
            rq, rk, rv, rg = (fa.jl_empty((Nl, D, Hh), bf, dev).normal_() for _ in range(4))
            rO, rl, rm = fa.ring_dense_fa(rq, rk, rv)
            t_rf = timeit(lambda: fa.ring_dense_fa(rq, rk, rv), 5)
            t_rb = timeit(lambda: fa.ring_dense_fa_backward(rq, rk, rv, rO, rg, rl, rm), 3)
            bO, bl, bm = (fa.jl_empty(sh, dt, dev) for sh, dt in (((Nl, D, Hh), bf), ((Nl, 1, Hh), torch.float32), ((Nl, 1, Hh), torch.float32)))
            t_blk = timeit(lambda: fa.dense_fa_(bO, bl, bm, rq, rk, rv), 10)        # one block, no exchange
            ntot = Nl * world
            ff, fb = 4.0 * Nl * ntot * D * Hh, 10.0 * Nl * ntot * D * Hh              # per rank
            gath = [torch.empty(Hh, D, Nl, dtype=bf, device=dev) for _ in range(world)]
            full = []
            for t_ in (rq, rk, rv):                                                   # gather the shards (check only, untimed)
                dist.all_gather(gath, t_.permute(2, 1, 0).contiguous())
                full.append(fa.jl_array(torch.cat(gath, dim=2).permute(2, 1, 0)))
            ring_err = None
            if rank == 0:
                fO, fl, fm = fa.dense_fa(*full, flags=fa.FA_FLAG_OUT_F32)
                ref = fO[:Nl].float()
                ring_err = float((rO.float() - ref).abs().max() / ref.abs().max())
                lse_err = float(((rl.log() + rm) - (fl[:Nl].log() + fm[:Nl])).abs().max())
                del fO, fl, fm
            del full, gath
            kv_bytes = 2 * Nl * D * Hh * 2 * (world - 1)
            extra["ring_dense"] = {
                "sequence_tokens": ntot, "tokens_per_gpu": Nl, "d": D, "heads": Hh, "dtype": "bf16",
                "fwd_ms": t_rf, "bwd_ms": t_rb, "fwd_tflops_per_gpu": ff / t_rf / 1e9, "bwd_tflops_per_gpu": fb / t_rb / 1e9,
                "single_block_kernel_ms": t_blk, "fwd_frac_of_single_gpu_kernel": (t_blk * world) / t_rf,
                "nvlink_bytes_sent_per_gpu_fwd": kv_bytes, "nvlink_ms_if_exposed_at_770GBs": kv_bytes / 770e6,
                "ring_parity_err": ring_err, "ring_parity_lse_abs_err": lse_err if rank == 0 else None,
                "parity": "rank 0 slice vs single-GPU dense_fa (fp32 output) on the gathered sequence, max-abs / max-abs"}
            del rq, rk, rv, rg, rO, rl, rm, bO, bl, bm
            torch.cuda.empty_cache()
            # (2) ONE 256^3 windowed volume (W = 5, stride 5, pad 3, d = 64, bf16) cut into slabs on window boundaries
            #     (no exchange needed: non-overlapping windows); rank 0 also runs the whole volume for the speed-up and a
            #     bit-for-bit check of its slab.
            S3 = 256
            plan = fa.windowed_slab_plan((S3, S3, S3), 5, 5, 3, rank, world)
            planes = plan.plane_hi - plan.plane_lo
            sq, sk, sv, sg = (fa.jl_empty((S3, S3, planes, 64, 1), bf, dev).normal_() for _ in range(4))
            sy, sl_, sm_ = fa.windowed_fa_slab(sq, sk, sv, 5, plan, stride=5, pad=3)
            t_sf = timeit(lambda: fa.windowed_fa_slab(sq, sk, sv, 5, plan, stride=5, pad=3), 5)
            t_sb = timeit(lambda: fa.windowed_fa_slab_backward(sq, sk, sv, sg, sl_, sm_, 5, plan, stride=5, pad=3), 3)
            cover = torch.tensor([planes, plan.nwin], device=dev, dtype=torch.int64)
            dist.all_reduce(cover)
            one = None
            shapes = [None] * world
            dist.all_gather_object(shapes, planes)
            maxp = max(shapes)
            parts = {}
            for name, t_ in (("q", sq), ("k", sk), ("v", sv)):                        # gather the slabs (check only, untimed)
                mine = torch.zeros(1, 64, maxp, S3, S3, dtype=bf, device=dev)         # equal-sized pieces for all_gather
                mine[:, :, :planes].copy_(t_.permute(4, 3, 2, 1, 0))
                bufs = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(bufs, mine)
                parts[name] = fa.jl_array(torch.cat([b_[:, :, :shapes[r]] for r, b_ in enumerate(bufs)], dim=2).permute(4, 3, 2, 1, 0)) if rank == 0 else None
                del bufs, mine
            if rank == 0:
                fy, fl_, fm_ = fa.windowed_fa(parts["q"], parts["k"], parts["v"], 5, 5, 3)
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a_.record()
                for _ in range(3):
                    fa.windowed_fa(parts["q"], parts["k"], parts["v"], 5, 5, 3)
                b_.record(); torch.cuda.synchronize()
                one = a_.elapsed_time(b_) / 3
                same = bool(torch.equal(fy[:, :, plan.plane_lo:plan.plane_hi].contiguous(), sy.contiguous()))
                del fy, fl_, fm_
            dist.barrier()
            ntok3 = S3 ** 3
            extra["slab_windowed"] = {
                "volume": [S3, S3, S3], "W": 5, "stride": 5, "pad": 3, "d": 64, "dtype": "bf16",
                "planes_covered": int(cover[0]), "window_planes_covered": int(cover[1]),
                "fwd_ms": t_sf, "bwd_ms": t_sb, "one_gpu_fwd_ms": one, "fwd_speedup_vs_one_gpu": (one / t_sf) if one else None,
                "fwd_alg_gbs_total": 4 * ntok3 * 64 * 2 / t_sf / 1e6, "bwd_alg_gbs_total": 7 * ntok3 * 64 * 2 / t_sb / 1e6,
                "slab_bitwise_equal_to_one_gpu": same if rank == 0 else None}
            del sq, sk, sv, sg, sy, sl_, sm_, parts
            torch.cuda.empty_cache()

    # ---- end to end: HOST buffers through the public API (fa_dense_fwd_host), H2D + kernel + D2H every step, on the
    #      named config (B = 512 per GPU: 3 GiB in, 1 GiB + stats out).  The buffers are page-locked and NUMA-local to the
    #      rank's GPU (fa_host_alloc) -- with 8 ranks streaming at once the shared host side is the limiter, so the line
    #      also reports the host<->device copy rates measured the same way (all ranks at once) as the ceiling.
    e2e = None
    if not args.no_e2e:
        Be = Bn
        hq, hk, hv = (fa.jl_host_empty((N, D, Be), bf, local) for _ in range(3))
        for t_h, t_d in ((hq, q), (hk, k), (hv, v)):
            t_h.permute(2, 1, 0).copy_(t_d.permute(2, 1, 0))
        hO = fa.jl_host_empty((N, D, Be), bf, local)
        hl = fa.jl_host_empty((N, 1, Be), torch.float32, local)
        hm = fa.jl_host_empty((N, 1, Be), torch.float32, local)
        fa.dense_fa_(hO, hl, hm, hq, hk, hv)           # warm-up (staging arena, first touch)
        barrier()
        ne = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(ne):
            fa.dense_fa_(hO, hl, hm, hq, hk, hv)       # synchronises before returning
        barrier()
        te = (time.perf_counter() - t0) / ne
        # ceiling: plain pinned copies of the same bytes, every rank at once (H2D and D2H overlapped on two streams)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            for t_h, t_d in ((hq, q), (hk, k), (hv, v)):
                t_d.permute(2, 1, 0).copy_(t_h.permute(2, 1, 0), non_blocking=True)
        with torch.cuda.stream(s2):
            hO.permute(2, 1, 0).copy_(O.permute(2, 1, 0), non_blocking=True)
        barrier()
        tcopy = time.perf_counter() - t0
        if world > 1:
            tm = torch.tensor([te, tcopy], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            te, tcopy = float(tm[0].item()), float(tm[1].item())
        h2d, d2h = 3 * N * D * Be * 2, N * D * Be * 2 + 2 * N * Be * 4
        e2e = {"value": FLOPS_PER_BATCH_ELT * Be * world / te / 1e12, "unit": "TFLOP/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "batch_per_gpu": Be, "ms_per_step": te * 1e3, "steps": ne,
               "h2d_gbs_aggregate": h2d * world / te / 1e9, "d2h_gbs_aggregate": d2h * world / te / 1e9,
               "copy_only_ms": tcopy * 1e3, "copy_only_h2d_gbs_aggregate": h2d * world / tcopy / 1e9,
               "frac_of_copy_ceiling": tcopy / te,
               "host_memory": "page-locked, NUMA-local to each rank's GPU (fa_host_alloc)",
               "api": "fa_dense_fwd_host via fa_sm100a.dense_fa_ on host tensors (3-stream chunked pipeline)"}
        del hq, hk, hv, hO, hl, hm

    if rank == 0:
        peaks = measured_peaks()
        traffic, traffic_src = None, "no ncu capture of this kernel at this batch in profiles/traffic.json"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj.get(f"dense_fwd_d{D}_bf16_N{N}_B{Bn}")
            if ent:
                traffic, traffic_src = ent["dram_bytes"], f"bytes per launch, ncu --set full: {ent['source']}"
        except Exception:
            pass
        roof = {"bound": "tensor", "achieved": per_gpu, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": per_gpu / peaks["bf16_tflops"],
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel at this batch, from the committed
                # ncu --set full capture named in profiles/traffic.json (null when no capture matches the kernel + batch)
                "traffic": traffic, "traffic_unit": traffic_src,
                "algorithmic_bytes_per_launch": (4.0 * N * D * 2 + 8.0 * N) * Bn,
                "peak_source": peaks["source"] + " burst (cuBLAS bf16 GEMM)",
                "frac_of_sustained": per_gpu / peaks["bf16_tflops_sustained"],
                "frac_of_nominal_2250": per_gpu / 2250.0,
                "kernel": kernel_name, "algorithmic_flops_per_launch": FLOPS_PER_BATCH_ELT * Bn,
                "kernel_ms": ms}
        line = {
            "metric": "dense_fa forward attention TFLOP/s (4*N^2*d*B / time)", "value": value, "unit": "TFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic randn (seeded, generated Float32 on host then cast)",
            "config": {"workload": WORKLOAD, "batch_per_gpu": Bn, "l2": "inputs (3 GiB) larger than L2; no flush needed",
                       "sharding": "batch*heads across ranks, no collective"},
            "tokens_per_s": Bn * N * world / (ms * 1e-3), "tokens_per_s_batch_only": 16 * N * world * (Bn / 512) / (ms * 1e-3),
            "gpu_launches": args.steps, "clocks": clocks, "roofline": roof, "e2e": e2e, "extra": extra,
        }
        if not args.no_cpu:
            cv, ct, cores = cpu_reference_tflops(2, 1, 8)
            line["cpu_baseline"] = {"value": cv, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                                    "sample": f"oracle/fa_oracle.c dense_fa Float32 N={N} d={D} B=8 ({ct:.2f} s/step), {cores} OpenMP threads; "
                                              "CPU restatement of reference (Julia runtime unavailable)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
