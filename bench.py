#!/usr/bin/env python
"""bench.py -- HR query pixels/s of the DIINN query decoder (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c4|c2x2|c2x3|c2x4|c1] [--precision fp16|bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)
    python bench.py --impl reference ...        (the reference's own decoder on the host CPU cores, see below)

One "step" = one decode of the workload image: the HR query grid is sharded by row tiles over the N ranks (feature map
and weights replicated), every rank decodes its tile with the fused sm_100a kernels through the C ABI and stores it into
every rank's image buffer over NVLink inside the same kernel (or, --assembly nccl, NCCL all-gathers the tiles), so every
rank ends up with the assembled (B,3,H_up,W_up) image. Default workload c3 = DIV2K-validation x4
shape (339x510 LR -> 1356x2040 HR), the configuration "ms per DIV2K x4 image" is quoted on; total work is fixed as N
grows ("scaling": "strong").

The JSON line carries: value (device-timed, inputs resident in HBM), e2e (host buffers through diinn_decode_host: H2D of
the feature map + decode + D2H of the image inside the timed region), roofline of the dominant kernel (stage B, CUDA-event
timed inside the timed region through the library's profiling hooks), cpu_baseline (N=1, rank 0), clocks, gpu_launches.
At N>1 it also carries `assembled_bit_identical` -- every rank compares its assembled image, for BOTH assembly paths, with
a single-GPU decode of the whole image; the run FAILS on a mismatch -- the same `image_checksum` as the N=1 line, and
`extra_multi_gpu_configs.c4` (the 8K x12 configuration north_star's scaling target is quoted on).

--impl reference times the reference's own ImplicitDecoder (oracle/_ref, placed there unmodified by oracle/make_ref.py at
build time) on the host CPU cores through its public forward(x, size, bsize=30000), the way sr_module.py:160 drives it;
each step is a bounded sample of the workload (the image cropped to its first LR rows, ~2e5 HR pixels). Without
oracle/_ref it falls back to the oracle's port of the same algorithm and says so (kind "port").
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
MMA_TERMS = {"fp16": 1, "bf16": 1, "fp32": 3}     # MMAs per product (fp32 = fp16 hi+lo split: hi.hi + lo.hi + hi.lo)
COMPUTE_DESC = {
    "fp16": "tcgen05 fp16 operands (11-bit mantissa, saturating), fp32 TMEM accumulation",
    "bf16": "tcgen05 bf16 operands, fp32 TMEM accumulation",
    "fp32": "tcgen05 fp16 hi+lo split operands, three MMAs per product, fp32 TMEM accumulation (fp32-level precision)",
    "fp32_simt": "fp32 FMA on CUDA cores",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32", "fp32_simt"])
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
    for r0, r1 in diinn_b200.tile_partition(H, H_up, world):
        lo = max(min(int((r0 + 0.5) * H / H_up), H - 1) - 1, 0)
        hi = min(min(int((r1 - 0.5) * H / H_up), H - 1) + 2, H)
        total += B * C * (hi - lo) * W * 4
    return int(total)


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own decoder (oracle/_ref) on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------------------------------
def cpu_sample_shape(workload, px_target=200_000):
    """The workload image cropped to its first LR rows so that one reference forward covers ~px_target HR pixels at the
    workload's own width and scale factor: (B, h, W) LR -> (h * s, W_up) HR."""
    from diinn_b200 import synth
    B, H, W, H_up, W_up = synth.CONFIGS[workload]
    s = H_up / H
    h = max(3, min(H, int(round(px_target / (B * W_up) / s))))
    return B, h, W, int(round(h * s)), W_up


def time_cpu_reference(workload, steps, warmup):
    """-> dict(px_per_s (best), mean_px_per_s, ms_per_step, cores, host_cpus, kind, sample)"""
    import torch
    from diinn_b200 import synth
    from oracle import make_ref
    B, h, W, hu, W_up = cpu_sample_shape(workload)
    npx = B * hu * W_up
    weights = synth.make_weights(seed=0)
    feat = synth.make_feat(1, *synth.CONFIGS[workload][:3])[:, :, :h].copy()
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would cripple the CPU arm)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    Ref = make_ref.import_reference_decoder()
    if Ref is not None:
        dec = Ref(mode=3, init_q=False).eval()
        dec.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in weights.items()}, strict=True)
        x = torch.from_numpy(feat)

        def run():
            with torch.no_grad():
                return dec(x, [hu, W_up], 30000)
        kind = "reference"
        what = ("the reference's ImplicitDecoder(mode=3, init_q=False).forward(x, size, bsize=30000) from oracle/_ref "
                "(unmodified diinn.py), as sr_module.py:160 drives it")
    else:
        from oracle import diinn_oracle as orc

        def run():
            return orc.decoder_forward_torch_cpu(weights, feat, (hu, W_up), bsize=30000)
        kind = "port"
        what = "oracle/_ref absent: the oracle's torch-CPU port of the same un-hoisted algorithm (decoder_forward_torch_cpu)"
    for _ in range(warmup):
        run()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    return dict(px_per_s=npx / min(ts), mean_px_per_s=npx * len(ts) / sum(ts), ms_per_step=1e3 * sum(ts) / len(ts),
                cores=torch.get_num_threads(), host_cpus=os.cpu_count(), kind=kind,
                sample=f"{workload} cropped to its first {h} LR rows -> {hu}x{W_up} HR = {npx} px per step at the workload's "
                       f"width and scale; {what}; torch {torch.__version__} CPU fp32")


def time_eager_gpu_port(workload, steps=5, warmup=2, bsize=30000 * 16):
    """Informative only ("the real same-box bar"): the reference's eager op sequence -- unfold, 576-channel
    nearest-exact gather, 9 convs, cat / relu / sin / mul kernels, query strips as in batched_step (diinn.py:149-160) --
    run by PyTorch on THIS GPU (the reference module from oracle/_ref when present, else the oracle's torch port)."""
    import torch
    from diinn_b200 import synth
    from oracle import make_ref
    B, H, W, H_up, W_up = synth.CONFIGS[workload]
    weights = synth.make_weights(seed=0)
    feat = synth.make_feat(1, B, H, W)
    Ref = make_ref.import_reference_decoder()
    if Ref is not None:
        dec = Ref(mode=3, init_q=False).eval().cuda()
        dec.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in weights.items()}, strict=True)
        x = torch.from_numpy(feat).cuda()

        def run():
            with torch.no_grad():
                return dec(x, [H_up, W_up], bsize)
        what = "reference ImplicitDecoder (oracle/_ref) on cuda:0"
    else:
        from oracle import diinn_oracle as orc
        run = lambda: orc.decoder_forward_torch_cpu(weights, feat, (H_up, W_up), bsize=bsize, device="cuda",  # noqa: E731
                                                    return_tensor=True)
        what = "oracle torch port on cuda:0 (includes the per-call H2D of the feature map and weights)"
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
            "what": f"{what}, fp32 (cudnn TF32 convs at PyTorch's default), bsize={bsize}; not the product path"}


L2_NOTE = ("no explicit flush: each step writes then re-reads the LR pre-activation tensor P (c3: 354 MB as fp16 on the 16-bit "
           "operand paths, 708 MB as fp32 on the fp32 path: 2.8x / 5.6x the 126 MB L2) plus 33 MB of output, so no step finds its "
           "working set in L2")


def bench_config(workload):
    """the `config` object -- identical for both arms (the reference arm runs on OUR arm's config) and for every N; what is
    specific to an arm or to N (compute path, sharding) lives in the line's `path` object"""
    return {"workload": workload_desc(workload), "io_dtype": "fp32 feature map in, fp32 image out", "l2": L2_NOTE}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_cpu_reference(args.workload, max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": "HR query pixels/s", "value": r["mean_px_per_s"], "unit": "px/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": bench_config(args.workload),
        "path": {"note": "CPU arm: each step is a bounded sample of the workload (see cpu_baseline.sample); px/s is the metric "
                         "and the per-pixel work is uniform over the image"},
        "cpu_baseline": {"value": r["mean_px_per_s"], "unit": "px/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"], "host_cpus": r["host_cpus"]},
        "e2e": {"value": r["mean_px_per_s"], "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def timed(fn, n, barrier):
    """device time of n calls of fn (ms per call), events on the current stream, barrier + sync both sides"""
    import torch
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / n


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

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

    def make_decoder(precision, mode=3, init_q=False, w=None):
        d = diinn_b200.FusedImplicitDecoder(mode=mode, init_q=init_q, precision=precision)
        return diinn_b200.load_numpy_weights(d, w if w is not None else weights).to(dev).requires_grad_(False)

    dec = make_decoder(args.precision)
    tensor_path = args.precision != "fp32_simt"
    feat_host = torch.from_numpy(synth.make_feat(1, B, H, W)).pin_memory()
    feat = feat_host.to(dev)
    npx = B * H_up * W_up
    parts = diinn_b200.tile_partition(H, H_up, world)
    r0, r1 = parts[rank]

    def step(assembly=args.assembly, d=dec, x=feat, size=(H_up, W_up)):
        if world == 1:
            return d(x, size)
        if assembly == "fused":
            # clone=False: the assembled image stays in the symmetric buffer (valid until the second next call)
            return diinn_b200.decode_sharded_fused(d, x, size, clone=False)
        return diinn_b200.decode_sharded(d, x, size, bands=1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    identical = None
    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            out = step()
        barrier()
        # ---- correctness of what is being timed: full-image checksum, and at N>1 the assembled image of BOTH assembly
        # paths against a single-GPU decode of the whole image, on every rank (bit-identical or the run fails)
        single = dec(feat, (H_up, W_up))
        image_checksum = float(single.double().sum())
        if world > 1:
            ok = True
            for mode in ("fused", "nccl"):
                got = step(mode)
                torch.cuda.synchronize()
                ok = ok and bool(torch.equal(got, single))
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            identical = bool(int(flag))
            if not identical:
                raise SystemExit(f"rank {rank}: the assembled image differs from the single-GPU decode (fused / nccl)")
            image_checksum = float(step().double().sum())   # of the assembled image the timed path produces
        del single

        # ---- N > 1: a sharded step is launch-bound (0.25 ms of GPU work at N = 8 against ~0.15 ms of Python + driver calls
        # per rank, and a hiccup on ANY rank stalls all of them at the step's barrier), so the timed loop replays a CUDA
        # graph of two consecutive steps (one per symmetric image buffer). The library is capturable by contract (no
        # allocation, no host sync inside decode); if the capture fails the loop runs eagerly and the line says so.
        graph, graph_note, launches_per_step = None, None, None
        if world > 1 and args.steps % 2 == 0 and os.environ.get("DIINN_BENCH_NO_GRAPH") is None:
            try:
                gs = torch.cuda.Stream(device=dev)
                gs.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(gs):
                    step()
                    step()
                torch.cuda.current_stream(dev).wait_stream(gs)
                barrier()
                l0 = dec.launch_count()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=gs, capture_error_mode="thread_local"):
                    step()
                    out = step()
                launches_per_step = (dec.launch_count() - l0) // 2
                barrier()
                graph.replay()                       # and the replayed image is still the single-GPU image
                torch.cuda.synchronize()
                okg = torch.tensor([1 if float(out.double().sum()) == image_checksum else 0], device=dev)
                dist.all_reduce(okg, op=dist.ReduceOp.MIN)
                if not int(okg):
                    raise RuntimeError("graph replay produced a different image")
                graph_note = "timed loop = steps/2 replays of a CUDA graph holding two consecutive steps"
            except Exception as e:  # noqa: BLE001  (any capture problem: fall back to the eager loop)
                graph, graph_note = None, f"eager loop (graph capture failed: {type(e).__name__}: {e})"[:200]
                torch.cuda.synchronize()
            flag = torch.tensor([1 if graph is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)   # all ranks replay, or none does
            if not int(flag):
                graph = None
                graph_note = graph_note if graph_note and graph_note.startswith("eager") else "eager loop (another rank could not capture)"

        dec.set_profiling(tensor_path and graph is None, dev)
        launches0 = dec.launch_count()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        if graph is not None:
            for _ in range(args.steps // 2):
                graph.replay()
        else:
            for _ in range(args.steps):
                out = step()
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        launches = launches_per_step * args.steps if graph is not None else dec.launch_count() - launches0
        kt = dec.kernel_times() if (tensor_path and graph is None) else None
        dec.set_profiling(False, dev)

        # ---- per-step distribution ("report best and median"), outside the timed region above; eager, and
        # with the library's per-kernel events when the timed region replayed a graph (events cannot be read out of a graph)
        n_dist = min(args.steps, 30)
        if graph is not None:
            dec.set_profiling(tensor_path, dev)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_dist + 1)]
        barrier()
        evs[0].record()
        for i in range(n_dist):
            out = step()
            evs[i + 1].record()
        barrier()
        per_step = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(n_dist))
        if graph is not None:
            kt = dec.kernel_times() if tensor_path else None
            dec.set_profiling(False, dev)

        # ---- e2e: host buffers through the C-ABI host entry (H2D feat + decode of this rank's tile + D2H tile)
        out_host = torch.empty((B, 3, r1 - r0, W_up), dtype=torch.float32).pin_memory()
        for _ in range(3):
            dec.decode_host(feat_host, (H_up, W_up), r0, r1, out_host, dev)
        ms_e2e = timed(lambda: dec.decode_host(feat_host, (H_up, W_up), r0, r1, out_host, dev), args.steps, barrier) * args.steps
        e2e_sum = torch.tensor([float(out_host.double().sum())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_sum)   # sum over the ranks' row tiles = checksum of the whole image
        e2e_checksum = float(e2e_sum)

    ms_other = 0.0
    multi = {}
    if world > 1:
        other = "nccl" if args.assembly == "fused" else "fused"
        with torch.no_grad():
            for _ in range(3):
                step(other)
            ms_other = timed(lambda: step(other), 20, barrier)
            if not args.no_extra and args.workload != "c4":
                # the 8K x12 configuration (north_star's scaling target): compute + assembly, both paths, and the rank's own
                # tile alone (compute only), max over ranks
                b4, h4, w4, hu4, wu4 = synth.CONFIGS["c4"]
                x4 = torch.from_numpy(synth.make_feat(1, b4, h4, w4)).to(dev)
                a4, c4 = diinn_b200.tile_partition(h4, hu4, world)[rank]
                legs = {"fused": lambda: step("fused", dec, x4, (hu4, wu4)), "nccl": lambda: step("nccl", dec, x4, (hu4, wu4)),
                        "compute_only": lambda: dec.forward_rows(x4, (hu4, wu4), a4, c4)}
                single4 = dec(x4, (hu4, wu4))
                ok4 = bool(torch.equal(legs["fused"](), single4)) and bool(torch.equal(legs["nccl"](), single4))
                del single4
                res = {}
                for name, fn in legs.items():
                    for _ in range(2):
                        fn()
                    res[name] = timed(fn, 10, barrier)
                t4 = torch.tensor([res["fused"], res["nccl"], res["compute_only"], 0.0 if ok4 else 1.0], dtype=torch.float64, device=dev)
                dist.all_reduce(t4, op=dist.ReduceOp.MAX)
                f4, n4, c4ms, bad4 = (float(v) for v in t4.cpu())
                px4 = b4 * hu4 * wu4
                multi["c4"] = {"workload": workload_desc("c4"), "ms_fused": f4, "px_per_s_fused": px4 / f4 * 1e3, "ms_nccl": n4,
                               "px_per_s_nccl": px4 / n4 * 1e3, "ms_compute_only_max_rank": c4ms,
                               "assembled_bit_identical": bad4 == 0.0, "steps": 10,
                               "note": "device-timed, max over ranks; compute_only = each rank's own 1/N row tile without "
                                       "assembly"}
                if bad4 != 0.0:
                    raise SystemExit(f"rank {rank}: c4 assembled image differs from the single-GPU decode")
                del x4
                torch.cuda.empty_cache()
    times = torch.tensor([ms_total, ms_e2e, ms_other], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_other = (float(v) for v in times.cpu())

    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        def time_decoder(d, name, n_it=5, query=False):
            b, h, w, hu, wu = synth.CONFIGS[name]
            x = torch.from_numpy(synth.make_feat(1, b, h, w)).to(dev)
            for _ in range(2):
                d(x, (hu, wu))
            torch.cuda.synchronize()
            ms = timed(lambda: d(x, (hu, wu)), n_it, torch.cuda.synchronize)
            res = {"ms": round(ms, 4), "px_per_s": b * hu * wu / ms * 1e3}
            if query:   # c5's sampled form: 16 patches x 2304 random query coordinates through query()
                coord, cell = (torch.from_numpy(v).to(dev) for v in synth.make_query(3, b, 2304))
                for _ in range(2):
                    d.query(x, coord, cell)
                ms = timed(lambda: d.query(x, coord, cell), n_it, torch.cuda.synchronize)
                res = (res, {"ms": round(ms, 4), "px_per_s": b * 2304 / ms * 1e3})
            return res

        with torch.no_grad():
            # the other BASELINE configs on one GPU in the run's precision (not the headline; parity for them lives in tests/)
            for name in ("c1", "c2x2", "c2x3", "c2x4", "c4", "c5"):
                if name == "c5":
                    extra["c5"], extra["c5_sampled_query"] = time_decoder(dec, name, query=True)
                else:
                    extra[name] = time_decoder(dec, name)
            dec._workspace = None
            torch.cuda.empty_cache()
            # configs 1-2 are stated in fp32 (and bf16): the fp32-PRECISION tensor path (fp16 hi+lo split) and bf16 operands
            for prec, names in (("fp32", ("c1", "c2x2", "c2x3", "c2x4", "c3")), ("bf16", ("c2x4", "c3")),
                                ("fp16", ("c2x4", "c3"))):
                if prec == args.precision:
                    continue
                d2 = make_decoder(prec)
                for name in names:
                    extra[f"{name}_{prec}"] = time_decoder(d2, name, n_it=3 if prec == "fp32" else 5)
                d2.release()
                del d2
                torch.cuda.empty_cache()
            if "c2x4_fp32" in extra:
                base = extra.get("c2x4_fp16", extra["c2x4"]) if args.precision != "fp32" else None
                if base:
                    extra["c2x4_fp32"]["throughput_vs_16bit_path"] = extra["c2x4_fp32"]["px_per_s"] / base["px_per_s"]
            # the other decoder wirings of the reference constructor on the headline shape (informative; parity in tests/)
            for mode, init_q in ((1, False), (2, False), (4, False), (3, True)):
                d2 = make_decoder(args.precision, mode, init_q, synth.make_weights(seed=0, mode=mode, init_q=init_q))
                extra[f"{args.workload}_mode{mode}_init_q{int(init_q)}"] = time_decoder(d2, args.workload, n_it=3)
                d2.release()
                del d2
            # LIIF-proper decoding (the reference's own LIIF imnet on the same kernels): LIIF.forward from the encoder output
            # on, local ensemble (4 evaluations per query) and plain, on a 256x256 -> 1024x1024 grid (informative)
            for ens in (True, False):
                lq = diinn_b200.load_liif_imnet(diinn_b200.FusedLIIFQuery(local_ensemble=ens, precision=args.precision if
                                                args.precision != "fp32_simt" else "fp32"),
                                                {"imnet." + k: v for k, v in synth.make_liif_weights(1).items()}).to(dev)
                b, h, w, hu, wu = synth.CONFIGS["c2x4"]
                xl = torch.from_numpy(synth.make_feat(1, b, h, w)).to(dev)
                coord, cell = lq.make_coord_and_cell(xl, (hu, wu))
                for _ in range(2):
                    lq.query_rgb(xl, coord, cell)
                ms = timed(lambda: lq.query_rgb(xl, coord, cell), 5, torch.cuda.synchronize)
                extra["liif_c2x4_ensemble" if ens else "liif_c2x4_plain"] = {
                    "ms": round(ms, 4), "queries_per_s": b * hu * wu / ms * 1e3, "imnet_evals_per_s": b * hu * wu * (4 if ens else 1) / ms * 1e3}
                lq.release()
                del lq, xl, coord, cell
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    ms_step = ms_total / args.steps
    value = npx / ms_step * 1e3
    terms = MMA_TERMS.get(args.precision, 1)
    line = {
        "metric": "HR query pixels/s", "value": value, "unit": "px/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": bench_config(args.workload),
        "path": {
            "loop": graph_note or "eager loop, one decode call per step",
            "compute": COMPUTE_DESC[args.precision],
            "sharding": (f"HR row tiles over {world} ranks, feature map and weights replicated, no data-path collective; "
                         + (f"assembly fused into stage B ({diinn_b200.sharding.last_fused_mode} over NVLink, two alternating "
                            "symmetric-memory image buffers + 1 barrier per step)" if args.assembly == "fused" else
                            "assembly by in-place NCCL all_gather_into_tensor per channel"))
                        if world > 1 else "single GPU, whole image",
        },
        "ms_per_div2k_x4_image": ms_step if args.workload == "c3" else None,
        "ms_per_step_best": per_step[0], "ms_per_step_median": per_step[len(per_step) // 2],
        "per_step_distribution": ("separate eager region of min(steps, 30) steps, one event per step"
                                  + (" and the library's per-kernel events (the timed region replays a CUDA graph)" if graph is not None else "")),
        "image_checksum": image_checksum,
        "e2e": {"value": npx / (ms_e2e / args.steps) * 1e3, "unit": "px/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d_bytes_all_ranks(H, W, H_up, world, B),
                "d2h_bytes_per_step": int(npx * 3 * 4),
                "api": "diinn_decode_host (C ABI, pinned host buffers; every rank uploads the LR rows its row tile reads "
                       "(tile + 3x3 halo) of the replicated feature map and downloads its own row tile; row bands pipeline "
                       "upload / decode / download)", "checksum": e2e_checksum},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if identical is not None:
        line["assembled_bit_identical"] = identical
    if kt is not None and kt["decodes"]:
        n = kt["decodes"]
        rows_px = B * (r1 - r0) * W_up
        ms_b = kt["stage_b_ms"] / n
        ach = terms * FLOP_STAGE_B_PER_PX * rows_px / (ms_b * 1e-3) / 1e12
        prof = {}
        pj = os.path.join(ROOT, "profiles", "r2_stage_b_traffic.json")
        if os.path.exists(pj):
            with open(pj) as f:
                prof = json.load(f)
        # a timed region of tens of milliseconds runs at boost clocks (burst regime); after ~50-100 ms the 1 kW power cap
        # pulls the clocks down and the sustained cuBLAS figure is the comparable one (both fractions are reported)
        burst = ms_total < 100.0
        peak = peaks["bf16_burst"] if burst else peaks["bf16_sustained"]
        fmt = {"fp16": 1, "bf16": 0, "fp32": 2}[args.precision]
        line["roofline"] = {
            "bound": "tensor",
            "kernel": f"stage_b_umma_kernel<CG=2, FMT={fmt}, kDump=0, kPix=0, kSel={int(terms == 1 and args.workload != 'c2x2' and args.workload != 'c2x3')}, kLiif=0, kTab={int(terms == 1 and args.workload in ('c1', 'c2x2', 'c2x3', 'c2x4', 'c3', 'c5'))}>",
            "achieved": ach, "peak": peak,
            "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": prof.get("dram_bytes_per_launch_c3") if args.workload == "c3" and world == 1 and terms == 1 else None,
            "traffic_source": prof.get("source"),
            "peak_source": peaks["source"] + ("; BURST cuBLAS bf16 figure: the timed region lasts %.0f ms" % ms_total if burst else
                                              "; sustained cuBLAS bf16 figure: the timed region lasts %.1f s" % (ms_total / 1e3)),
            "frac_of_burst_peak": ach / peaks["bf16_burst"], "frac_of_sustained_peak": ach / peaks["bf16_sustained"],
            "algorithmic_flop_per_px": FLOP_STAGE_B_PER_PX, "executed_tensor_flop_per_px": terms * FLOP_STAGE_B_PER_PX,
            "select_mma_flop_per_px": (6 * 2 * 256 * 16 if terms == 1 and args.workload in ("c3", "c2x4") else None),
            "avg_launch_ms": ms_b,
            "kernel_share_of_step": {"layout_nhwc": kt["layout_ms"] / n / ms_step,
                                     "stage_a_umma": kt["stage_a_ms"] / n / ms_step,
                                     "stage_b_umma": ms_b / ms_step},
            "whole_decode": {"executed_tensor_flop_per_px": terms * (FLOP_STAGE_B_PER_PX + FLOP_STAGE_A_PER_LR_PX * (H * W) / (H_up * W_up)),
                             "tflops": terms * (FLOP_STAGE_B_PER_PX * npx + FLOP_STAGE_A_PER_LR_PX * B * H * W) / (ms_step * 1e-3) / 1e12
                             if world == 1 else None,
                             "reference_arithmetic_tflops_equivalent": FLOP_REFERENCE_PER_PX * npx / (ms_step * 1e-3) / 1e12},
        }
    if world == 1:
        cpu = time_cpu_reference(args.workload, steps=3, warmup=1)
        line["cpu_baseline"] = {"value": cpu["px_per_s"], "unit": "px/s", "cores": cpu["cores"], "kind": cpu["kind"],
                                "sample": cpu["sample"] + "; best of 3 after 1 warm-up", "host_cpus": cpu["host_cpus"]}
        try:
            line["eager_gpu_reference"] = time_eager_gpu_port(args.workload)
        except Exception as e:  # informative leg only (e.g. out of memory on a busy box)
            line["eager_gpu_reference"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if world > 1:
        line["other_assembly"] = {"mode": "nccl" if args.assembly == "fused" else "fused", "ms_per_step": ms_other,
                                  "px_per_s": npx / ms_other * 1e3}
        if multi:
            line["extra_multi_gpu_configs"] = multi
    if extra:
        line["extra_single_gpu_configs"] = extra
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
