#!/usr/bin/env python
"""bench.py -- HR query pixels/s of the DIINN query decoder (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c4|c2x2|c2x3|c2x4|c1] [--precision bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)
    python bench.py --impl reference ...        (the reference algorithm's CPU path, see below)

One "step" = one decode of the workload image: the HR query grid is sharded by row tiles over the N ranks
(feature map and weights replicated), every rank decodes its tile with the fused sm_100a kernels through the C ABI,
and NCCL all-gathers the tiles so every rank ends up with the assembled (B,3,H_up,W_up) image.
Default workload c3 = DIV2K-validation x4 shape (339x510 LR -> 1356x2040 HR), the configuration "ms per DIV2K x4 image"
is quoted on; total work is fixed as N grows ("scaling": "strong").

The JSON line carries: value (device-timed, inputs resident in HBM), e2e (host buffers through diinn_decode_host:
H2D of the feature map + decode + D2H of the image inside the timed region), roofline of the dominant kernel (stage B,
CUDA-event timed inside the timed region through the library's profiling hooks), cpu_baseline (the oracle's torch-CPU
port of the reference algorithm on a bounded row band of the same workload, rank 0, N=1 only), clocks, gpu_launches.

--impl reference times the reference's own algorithm on the host CPU cores: /root/reference is Python and does not exist
on the GPU box, so this arm runs the oracle's port of it (oracle/diinn_oracle.py: decoder_forward_torch_cpu, the same
un-hoisted 1.97 MFLOP/px algorithm on PyTorch CPU kernels with all host threads); each step is a bounded row band.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_STAGE_B_PER_PX = 3 * 2 * 256 * 512          # tensor-core FLOPs stage B executes per HR pixel
FLOP_STAGE_A_PER_LR_PX = 2 * 576 * 1024           # tensor-core FLOPs stage A executes per LR pixel
FLOP_REFERENCE_PER_PX = 1969152                   # reference arithmetic


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16acc", "fp32"])
    ap.add_argument("--no-extra", action="store_true", help="skip the extra measurements of the other BASELINE configs")
    ap.add_argument("--assembly", default="fused", choices=["fused", "nccl"],
                    help="N>1: fused = stage B stores every pixel into all ranks' image buffers over NVLink (peer / "
                         "NVSwitch multicast stores, no collective); nccl = local tiles + NCCL all-gather")
    return ap.parse_args()


def workload_desc(name):
    from diinn_b200 import synth
    B, H, W, H_up, W_up = synth.CONFIGS[name]
    return f"{name}: DIINN mode=3 init_q=False, B={B}, LR {H}x{W} (HxW) -> HR {H_up}x{W_up}, x{H_up / H:g}"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(bf16_burst=float(d["bf16_tflops"]), bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), source="MEASURED_PEAKS.json (of measured)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="B200_PROFILING.md fallback (of fallback)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def h2d_bytes_all_ranks(H, W, H_up, world, B, C=64):
    """Bytes diinn_decode_host uploads per step, summed over ranks: the LR rows each rank's HR row tile reads
    (nearest-exact rows of the tile +-1 for the 3x3 unfold), fp32."""
    import diinn_b200
    total = 0
    for r0, r1 in diinn_b200.row_partition(H_up, world):
        lo = max(min(int((r0 + 0.5) * H / H_up), H - 1) - 1, 0)
        hi = min(min(int((r1 - 0.5) * H / H_up), H - 1) + 2, H)
        total += B * C * (hi - lo) * W * 4
    return int(total)


def cpu_reference_band(workload, rows_px_target=200_000):
    """Bounded sample of `workload` for the CPU arm: the first HR rows of the image, ~200k pixels."""
    from diinn_b200 import synth
    B, H, W, H_up, W_up = synth.CONFIGS[workload]
    nrows = max(1, min(H_up, rows_px_target // (B * W_up)))
    return (0, nrows), B * nrows * W_up


def time_eager_gpu_port(workload, steps=5, warmup=2, bsize=30000 * 16):
    """Informative only ("the real same-box bar"): the reference's eager op sequence -- unfold, 576-channel
    nearest-exact gather, 9 convs, cat / relu / sin / mul kernels, query strips as in batched_step (diinn.py:149-160) --
    run by PyTorch on THIS GPU through the oracle's torch port (the reference itself is not on the GPU box)."""
    import torch
    from diinn_b200 import synth
    from oracle import diinn_oracle as orc
    B, H, W, H_up, W_up = synth.CONFIGS[workload]
    weights = synth.make_weights(seed=0)
    feat = synth.make_feat(1, B, H, W)
    run = lambda: orc.decoder_forward_torch_cpu(weights, feat, (H_up, W_up), bsize=bsize, device="cuda",  # noqa: E731
                                                return_tensor=True)
    for _ in range(warmup):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": ms, "px_per_s": B * H_up * W_up / ms * 1e3,
            "what": f"oracle torch port on cuda:0, fp32 (cudnn TF32 convs at PyTorch's default), bsize={bsize}, includes the "
                    "per-call H2D of the 44 MB feature map and weights; not the product path, not a parity reference"}


def time_cpu_port(workload, steps, warmup):
    import torch
    from diinn_b200 import synth
    from oracle import diinn_oracle as orc
    B, H, W, H_up, W_up = synth.CONFIGS[workload]
    weights = synth.make_weights(seed=0)
    feat = synth.make_feat(1, B, H, W)
    rows, npx = cpu_reference_band(workload)
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would cripple the CPU arm)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    for _ in range(warmup):
        orc.decoder_forward_torch_cpu(weights, feat, (H_up, W_up), rows=rows)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.decoder_forward_torch_cpu(weights, feat, (H_up, W_up), rows=rows)
        ts.append(time.perf_counter() - t0)
    return dict(px_per_s=npx / min(ts), mean_px_per_s=npx * len(ts) / sum(ts), ms_per_step=1e3 * sum(ts) / len(ts),
                cores=torch.get_num_threads(), host_cpus=os.cpu_count(),
                sample=f"HR rows [{rows[0]},{rows[1]}) of {workload} = {npx} px per step, full reference arithmetic "
                       f"(materialised 576-ch gather + 9 convs), torch {torch.__version__} CPU fp32")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_cpu_port(args.workload, max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": "HR query pixels/s", "value": r["mean_px_per_s"], "unit": "px/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": workload_desc(args.workload), "note": "oracle port of the reference algorithm (the Python "
                   "reference tree cannot travel to the GPU box); bounded row-band sample per step"},
        "cpu_baseline": {"value": r["mean_px_per_s"], "unit": "px/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"], "host_cpus": r["host_cpus"]},
        "e2e": {"value": r["mean_px_per_s"], "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import diinn_b200
    from diinn_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N bench.py --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, H, W, H_up, W_up = synth.CONFIGS[args.workload]
    weights = synth.make_weights(seed=0)
    dec = diinn_b200.load_numpy_weights(diinn_b200.FusedImplicitDecoder(mode=3, precision=args.precision), weights).to(dev)
    feat_host = torch.from_numpy(synth.make_feat(1, B, H, W)).pin_memory()
    feat = feat_host.to(dev)
    npx = B * H_up * W_up
    parts = diinn_b200.row_partition(H_up, world)
    r0, r1 = parts[rank]

    def step(assembly=args.assembly):
        if world == 1:
            return dec(feat, (H_up, W_up))
        if assembly == "fused":
            # clone=False: the assembled image stays in the symmetric buffer (overwritten by the next step)
            return diinn_b200.decode_sharded_fused(dec, feat, (H_up, W_up), clone=False)
        return diinn_b200.decode_sharded(dec, feat, (H_up, W_up), bands=1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            out = step()
        barrier()
        dec.set_profiling(args.precision != "fp32", dev)
        launches0 = dec.launch_count()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            out = step()
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        launches = dec.launch_count() - launches0
        kt = dec.kernel_times() if args.precision != "fp32" else None
        dec.set_profiling(False, dev)

        # ---- per-step distribution ("report best and median"), outside the timed region above
        n_dist = min(args.steps, 30)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_dist + 1)]
        barrier()
        evs[0].record()
        for i in range(n_dist):
            out = step()
            evs[i + 1].record()
        barrier()
        per_step = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(n_dist))

        # ---- e2e: host buffers through the C-ABI host entry (H2D feat + decode of this rank's tile + D2H tile)
        out_host = torch.empty((B, 3, r1 - r0, W_up), dtype=torch.float32).pin_memory()
        for _ in range(3):
            dec.decode_host(feat_host, (H_up, W_up), r0, r1, out_host, dev)
        barrier()
        t0 = time.perf_counter()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(args.steps):
            dec.decode_host(feat_host, (H_up, W_up), r0, r1, out_host, dev)
        e3.record()
        barrier()
        ms_e2e = e2.elapsed_time(e3)
        checksum = float(out_host.double().sum())

    ms_other = 0.0
    if world > 1:  # the other assembly path, for the record
        other = "nccl" if args.assembly == "fused" else "fused"
        with torch.no_grad():
            for _ in range(3):
                step(other)
            barrier()
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o0.record()
            for _ in range(20):
                step(other)
            o1.record()
            barrier()
            ms_other = o0.elapsed_time(o1) / 20
    times = torch.tensor([ms_total, ms_e2e, ms_other], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_other = (float(v) for v in times.cpu())

    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        # other BASELINE configs on one GPU (not the headline; parity for them lives in tests/)
        with torch.no_grad():
            for name in ("c1", "c2x2", "c2x3", "c2x4", "c4", "c5"):
                b, h, w, hu, wu = synth.CONFIGS[name]
                x = torch.from_numpy(synth.make_feat(1, b, h, w)).to(dev)
                for _ in range(2):
                    dec(x, (hu, wu))
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                n_it = 5
                for _ in range(n_it):
                    dec(x, (hu, wu))
                a1.record()
                torch.cuda.synchronize()
                ms = a0.elapsed_time(a1) / n_it
                extra[name] = {"ms": round(ms, 4), "px_per_s": b * hu * wu / ms * 1e3}
                if name == "c5":   # the sampled form: 16 patches x 2304 random query coordinates through query()
                    coord, cell = (torch.from_numpy(v).to(dev) for v in synth.make_query(3, b, 2304))
                    for _ in range(2):
                        dec.query(x, coord, cell)
                    torch.cuda.synchronize()
                    a0.record()
                    for _ in range(n_it):
                        dec.query(x, coord, cell)
                    a1.record()
                    torch.cuda.synchronize()
                    ms = a0.elapsed_time(a1) / n_it
                    extra["c5_sampled_query"] = {"ms": round(ms, 4), "px_per_s": b * 2304 / ms * 1e3}
                del x
        dec._workspace = None
        torch.cuda.empty_cache()
        # the other decoder wirings of the reference constructor on the headline shape (informative; parity in tests/)
        with torch.no_grad():
            b, h, w, hu, wu = synth.CONFIGS[args.workload]
            x = torch.from_numpy(synth.make_feat(1, b, h, w)).to(dev)
            for mode, init_q in ((1, False), (2, False), (4, False), (3, True)):
                d2 = diinn_b200.load_numpy_weights(
                    diinn_b200.FusedImplicitDecoder(mode=mode, init_q=init_q, precision=args.precision),
                    synth.make_weights(seed=0, mode=mode, init_q=init_q)).to(dev)
                for _ in range(2):
                    d2(x, (hu, wu))
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                n_it = 3
                for _ in range(n_it):
                    d2(x, (hu, wu))
                a1.record()
                torch.cuda.synchronize()
                ms = a0.elapsed_time(a1) / n_it
                extra[f"{args.workload}_mode{mode}_init_q{int(init_q)}"] = {"ms": round(ms, 4), "px_per_s": b * hu * wu / ms * 1e3}
                d2.release()
                del d2
            del x
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    ms_step = ms_total / args.steps
    value = npx / ms_step * 1e3
    line = {
        "metric": "HR query pixels/s", "value": value, "unit": "px/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {
            "workload": workload_desc(args.workload),
            "io_dtype": "fp32 feature map in, fp32 image out",
            "compute": {"bf16": "tcgen05 bf16 operands, fp32 TMEM accumulation", "fp16acc": "tcgen05; stage B fp16 operands, fp16 TMEM "
                        "accumulation", "fp32": "fp32 CUDA cores"}[args.precision],
            "sharding": (f"HR row tiles over {world} ranks, feature map and weights replicated, no data-path collective; "
                         + (f"assembly fused into stage B ({diinn_b200.sharding.last_fused_mode} over NVLink, symmetric "
                            "memory + 2 barriers)" if args.assembly == "fused" else
                            "assembly by in-place NCCL all_gather_into_tensor per channel"))
                        if world > 1 else "single GPU, whole image",
            "l2": "no explicit flush: each step writes then re-reads the 708 MB fp32 LR pre-activation tensor P "
                  "(5.6x the 126 MB L2) plus 33 MB of output, so no step finds its working set in L2",
        },
        "ms_per_div2k_x4_image": ms_step if args.workload == "c3" else None,
        "ms_per_step_best": per_step[0], "ms_per_step_median": per_step[len(per_step) // 2],
        "e2e": {"value": npx / (ms_e2e / args.steps) * 1e3, "unit": "px/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d_bytes_all_ranks(H, W, H_up, world, B),
                "d2h_bytes_per_step": int(npx * 3 * 4),
                "api": "diinn_decode_host (C ABI, pinned host buffers; every rank uploads the LR rows its row tile reads "
                       "(tile + 3x3 halo) of the replicated feature map and downloads its own row tile)", "checksum": checksum},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if kt is not None and kt["decodes"]:
        n = kt["decodes"]
        rows_px = B * (r1 - r0) * W_up
        ms_b = kt["stage_b_ms"] / n
        ach = FLOP_STAGE_B_PER_PX * rows_px / (ms_b * 1e-3) / 1e12
        prof = {}
        pj = os.path.join(ROOT, "profiles", "r1_stage_b_traffic.json")
        if os.path.exists(pj):
            with open(pj) as f:
                prof = json.load(f)
        line["roofline"] = {
            "bound": "tensor", "kernel": "stage_b_umma_kernel<2, %s, false, false>" % ("true" if args.precision == "fp16acc" else "false"), "achieved": ach, "peak": peaks["bf16_sustained"],
            "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
            "traffic": prof.get("dram_bytes_per_launch_c3") if args.workload == "c3" and world == 1 else None,
            "peak_source": peaks["source"] + "; sustained cuBLAS bf16 figure because the kernel is timed inside the step",
            "frac_of_burst_peak": ach / peaks["bf16_burst"],
            "algorithmic_flop_per_px": FLOP_STAGE_B_PER_PX, "avg_launch_ms": ms_b,
            "kernel_share_of_step": {"layout_nhwc_bf16": kt["layout_ms"] / n / ms_step,
                                     "stage_a_umma": kt["stage_a_ms"] / n / ms_step,
                                     "stage_b_umma": ms_b / ms_step},
            "whole_decode": {"executed_tensor_flop_per_px": FLOP_STAGE_B_PER_PX + FLOP_STAGE_A_PER_LR_PX * (H * W) / (H_up * W_up),
                             "tflops": (FLOP_STAGE_B_PER_PX * npx + FLOP_STAGE_A_PER_LR_PX * B * H * W) / (ms_step * 1e-3) / 1e12
                             if world == 1 else None,
                             "reference_arithmetic_tflops_equivalent": FLOP_REFERENCE_PER_PX * npx / (ms_step * 1e-3) / 1e12},
        }
    if world == 1:
        cpu = time_cpu_port(args.workload, steps=3, warmup=1)
        line["cpu_baseline"] = {"value": cpu["px_per_s"], "unit": "px/s", "cores": cpu["cores"], "kind": "port",
                                "sample": cpu["sample"] + "; best of 3 after 1 warm-up", "host_cpus": cpu["host_cpus"]}
        try:
            line["eager_gpu_reference"] = time_eager_gpu_port(args.workload)
        except Exception as e:  # informative leg only (e.g. out of memory on a busy box)
            line["eager_gpu_reference"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if world > 1:
        line["other_assembly"] = {"mode": "nccl" if args.assembly == "fused" else "fused", "ms_per_step": ms_other,
                                  "px_per_s": npx / ms_other * 1e3}
    if extra:
        line["extra_single_gpu_configs"] = extra
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
