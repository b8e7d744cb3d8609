#!/usr/bin/env python
"""bench.py -- LPSR plate crops/sec on synthetic 3x64x192 crops (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference arithmetic on the box's host cores (torch CPU port)

A step = one LPSR forward over one batch of B synthetic crops per GPU (weak scaling: per-GPU batch fixed, the path
shards by batch with no data-path collective).  `value` = crops/s with inputs resident in HBM; `e2e` = the
same metric through the C-ABI host-buffer call (lpsr_forward_host: pinned H2D + forward + D2H inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "lpsr_plate_crops_per_sec"
UNIT = "crops/s"
DENSE_FLOP_PER_PIXEL = 297_104          # dense-conv FLOPs per pixel (roofline numerator, whole forward)
# MACs per pixel of the layers the tensor-core kernel executes (K x N per layer, SURVEY 8a GEMM view)
UMMA_MAC_PER_PIXEL = {
    # AutoEncoder (reference arithmetic: depthwise 5x5 + pointwise 1x1 per DConv, per full-resolution pixel)
    "ae.conv_in": 27 * 12, "ae.enc0": 25 * 12 + 12 * 12, "ae.enc1": (25 * 48 + 48 * 12) / 4, "ae.dec0": (25 * 48 + 48 * 48) / 16,
    "ae.dec1": (25 * 12 + 12 * 48) / 4,
    "ae.conv_out": 108 * 3, "rdn.shallowF1": 147 * 32, "rdn.shallowF2": 288 * 32, "rdn.gff0": 128 * 32, "rdn.gff1": 288 * 32, "final_conv": 288 * 1,
    "rdb0": (288 + 432 + 576 + 720) * 16 + 96 * 32, "rdb2": (288 + 432 + 576 + 720) * 16 + 96 * 32,
    "csar1.conv_in": 2 * 288 * 32, "csar3.conv_in": 2 * 288 * 32,
    "csar1.tail": 32 * 64 + 64 * 32 + 64 * 32, "csar3.tail": 32 * 64 + 64 * 32 + 64 * 32,   # only when the tail runs on tensor cores
}
CSAR_TAIL_ELEMS_PER_PIXEL = 96          # read x_in + read x + write out, 32 ch each
# Layer-by-layer activation traffic of the reference graph (elements per full-resolution pixel: channels read + written by every layer,
# dense concatenation in place, residual reads included; SURVEY 8d quotes ~3.66 kB/px in bf16).  AutoEncoder 102 (conv_in 3+12, enc0
# 12+12, enc1 (48+12)/4, dec0 (48+48)/16, dec1 (12+48)/4+12, conv_out 12+3), shallowF1 3+32, shallowF2 32+32, each RDB 48+64+80+96+160,
# each CSAR 64+64+32(pool)+96, gff 128+32 and 32+32+32, final conv 32+2.  This is the HBM roofline of an implementation that runs the
# graph one layer at a time; fused layers (RDB tail, CSAR tail, pooled conv_in.2) move less than this.
LAYERWISE_ELEMS_PER_PIXEL = 102 + 35 + 64 + 2 * 448 + 2 * 256 + 160 + 96 + 34


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic():
    """DRAM bytes per forward of the tensor-core conv launches / CSAR tail launches from the committed `ncu --set full`
    capture (profiles/r1_ncu_traffic.json, written by tools/ncu_traffic.py); scaled linearly with the batch."""
    p = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else None


def preprocess_bench(model, dev, n=1024, with_cpu=True):
    """SURVEY 8f row n1 (the step in front of the hot path): `preprocess_for_sr` (inference/run.py:80-96) batched on the device.
    Times the whole public call (packing into pinned memory, one H2D copy, one kernel) over n synthetic BGR crops of ragged sizes,
    and -- as the reported baseline -- the reference's own per-plate recipe (cv2 + Pillow + torchvision) on a bounded sample."""
    import lpsr_b200
    rng = np.random.default_rng(0)
    plates = [rng.integers(0, 256, (int(rng.integers(16, 72)), int(rng.integers(60, 320)), 3), dtype=np.uint8) for _ in range(n)]
    for _ in range(2):
        lpsr_b200.preprocess_for_sr_batch(model, plates, device=dev)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        x = lpsr_b200.preprocess_for_sr_batch(model, plates, device=dev)
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / reps
    in_bytes = sum(p.size for p in plates)
    out = {"value": n / dt, "unit": "crops/s", "ms_per_batch": 1e3 * dt, "batch": n, "h2d_bytes_per_batch": in_bytes,
           "out_bytes_per_batch": int(x.numel() * 4), "api": "lpsr_b200.preprocess_for_sr_batch -> lpsr_preprocess_resize (C ABI)",
           "parity": "bit-exact with cv2.cvtColor + PIL.Image.resize(BICUBIC) + ToTensor (tests/golden/preprocess_cases.npz)"}
    if with_cpu:
        try:
            import cv2
            import torchvision.transforms as T
            from PIL import Image
            tf = T.Compose([T.ToTensor()])
            sample = plates[:128]
            t0 = time.perf_counter()
            for p in sample:
                tf(Image.fromarray(cv2.cvtColor(p, cv2.COLOR_BGR2RGB)).resize((192, 32), Image.BICUBIC)).unsqueeze(0)
            out["cpu_reference"] = {"value": len(sample) / (time.perf_counter() - t0), "unit": "crops/s", "cores": 1, "kind": "reference",
                                    "sample": f"{len(sample)} crops, the reference's per-plate recipe (cv2 + Pillow + torchvision), host only"}
        except Exception as exc:   # the libraries are optional on the GPU box
            out["cpu_reference"] = {"unavailable": repr(exc)}
    return out


def load_shipped_weights():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_best_model.npz")))


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference arithmetic (torch CPU ops, fp32) on the box's host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(args, steps, warmup, sample_b, budget_s=None):
    from oracle import lpsr_torch_port as port   # bench.py's cpu_baseline / reference leg may execute oracle/

    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(cores)
    W = port.to_torch_weights(load_shipped_weights())
    x = torch.rand(sample_b, 3, args.height, args.width, generator=torch.Generator().manual_seed(0))
    for _ in range(warmup):
        port.lpsr_forward(x, W)
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        port.lpsr_forward(x, W)
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_start > budget_s:
            break
    total = sum(times)
    return {"value": sample_b * len(times) / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} steps x {sample_b} crops of 3x{args.height}x{args.width} fp32, torch {torch.__version__} "
                      f"CPU ops in the reference's op order (oracle/lpsr_torch_port.py), {cores} threads",
            "ms_per_step": 1e3 * total / len(times), "steps": len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_b = 16
    r = cpu_reference_run(args, args.steps, args.warmup, sample_b)
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": workload_config(args, sample_b, "fp32", note="CPU sample of the same workload"),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, batch, precision, note=None):
    cfg = {"workload": f"LPSR forward {precision}, batch {batch} of 3x{args.height}x{args.width} synthetic crops per GPU "
                       f"(BASELINE.json configs[2])", "per_gpu_batch": batch, "global_batch": batch * args.gpus,
           "crop": [3, args.height, args.width], "precision": precision, "weights": "reference best_model.pth (fixture)",
           "parallelism": f"batch-sharded x{args.gpus}, replicated weights, no per-layer collective",
           "l2": "inputs (151 MB/step at B=1024) and activations (GBs) exceed the 126 MB L2; no explicit flush"}
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    import lpsr_b200
    from lpsr_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, H, Wd = args.batch, args.height, args.width

    model = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision=args.precision)
    model.load_live_weights(load_shipped_weights())
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(1000 + rank)
    x_host = torch.rand(B, 3, H, Wd, generator=g).pin_memory()
    x = x_host.to(dev)
    launches = model.launch_count(B, H, Wd)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput -------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        y = model(x)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        y = model(x)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---------------- per-kernel device times (CUDA events between launches, same stream) --------------------------
    lib = capi.load_library()
    h = model._handle(dev)
    ws = model._workspace(h, local, B, H, Wd)
    cap = 64
    fam_ms, n_prof = {}, min(args.steps, 5)
    names = []
    for _ in range(n_prof):
        ms = (C.c_float * cap)()
        nm = C.create_string_buffer(cap * 64)
        n = lib.lpsr_forward_profiled(h, x.data_ptr(), y.data_ptr(), B, H, Wd, model._aligned_ptr(ws),
                                      ws.numel() - (model._aligned_ptr(ws) - ws.data_ptr()),
                                      torch.cuda.current_stream(dev).cuda_stream, ms, nm, 64, cap)
        capi.check(min(n, 0), h, "lpsr_forward_profiled")
        names = [nm.raw[i * 64:(i + 1) * 64].split(b"\0")[0].decode() for i in range(n)]
        for i, name in enumerate(names):
            fam_ms[name] = fam_ms.get(name, 0.0) + ms[i] / n_prof
    pk = peaks()
    P = ((H + 3) // 4 * 4) * ((Wd + 3) // 4 * 4)
    per_kernel = {}
    for name, t in fam_ms.items():
        kind = name.split(":")[1]
        per_kernel.setdefault(kind, [0.0, 0])
        per_kernel[kind][0] += t
        per_kernel[kind][1] += names.count(name)
    is_umma = lambda n_: n_.split(":")[1].startswith("umma_conv") or n_.endswith(":csar_tail_umma")
    umma_ms = sum(t for n_, t in fam_ms.items() if is_umma(n_))
    umma_tags = {n_.split(":")[0] for n_ in names if is_umma(n_)}
    umma_flops = 2.0 * B * P * sum(v for k, v in UMMA_MAC_PER_PIXEL.items() if k in umma_tags)
    n_umma = sum(1 for n_ in names if is_umma(n_))
    esz = 4 if args.precision == "fp32" else 2
    # CSAR tail = everything after conv_in: pooling, channel gate, spatial MLP, gating, conv_out, residual
    tail_ms = sum(t for n_, t in fam_ms.items() if ".tail:" in n_)
    n_tail = sum(1 for n_ in names if ".tail:" in n_)
    tail_bytes = 2.0 * B * P * CSAR_TAIL_ELEMS_PER_PIXEL * esz
    total_prof_ms = sum(fam_ms.values())
    tr = ncu_traffic()
    scale = (B * P) / float(tr["batch"] * tr["pixels_per_crop"]) if tr else 0.0
    if n_umma:
        ach = umma_flops / (umma_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": f"umma_conv_kernel (tcgen05 implicit-GEMM conv, {n_umma} launches/forward)",
                    "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    "peak_kind": f"bf16 dense sustained, of {pk['src']}",
                    "traffic": (tr["umma_dram_bytes_per_forward"] * scale) if tr else None,
                    "traffic_note": "sum of dram__bytes_read+write over the family's launches of one forward (ncu --set full, profiles/), scaled to this batch",
                    "algorithmic_flops_per_forward": umma_flops, "kernel_ms_per_forward": umma_ms,
                    "share_of_step": umma_ms / total_prof_ms}
    else:   # fp32 mode: FFMA direct convolution dominates; still reported against the bf16 tensor peak
        d_ms = sum(t for n_, t in fam_ms.items() if n_.endswith(":conv_direct"))
        fl = 2.0 * B * P * (DENSE_FLOP_PER_PIXEL / 2 - 12288)
        ach = fl / (d_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "conv_direct_kernel (FFMA, fp32 parity mode)", "achieved": ach, "peak": pk["tf_sustained"],
                    "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"], "peak_kind": f"bf16 dense sustained, of {pk['src']}",
                    "traffic": None, "share_of_step": d_ms / total_prof_ms}
    roofline_csar = {"bound": "hbm", "kernel": f"CSAR tail (pool + channel gate + spatial MLP + gating + conv_out + residual; {n_tail} launches/forward)", "achieved": tail_bytes / (tail_ms * 1e-3) / 1e9,
                     "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": tail_bytes / (tail_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                     "peak_kind": f"copy bandwidth, of {pk['src']}", "traffic": (tr["tail_dram_bytes_per_forward"] * scale) if tr else None,
                     "algorithmic_bytes_per_forward": tail_bytes, "kernel_ms_per_forward": tail_ms}
    conv_frac_whole = value / world * P * DENSE_FLOP_PER_PIXEL / 1e12 / pk["tf_sustained"]
    lw_bytes = float(B) * P * LAYERWISE_ELEMS_PER_PIXEL * esz
    roofline_layerwise = {"bound": "hbm", "kernel": "whole forward, layer-by-layer activation traffic of the reference graph",
                          "achieved": lw_bytes / (ms_step * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                          "frac": lw_bytes / (ms_step * 1e-3) / 1e9 / pk["hbm_gbs"], "peak_kind": f"copy bandwidth, of {pk['src']}",
                          "algorithmic_bytes_per_forward": lw_bytes,
                          "note": "fused kernels move fewer bytes than this; > 1.0 would mean the layer-by-layer HBM bound is beaten"}

    # ---------------- end to end through the C-ABI host-buffer call --------------------------------------------------
    y_host = torch.empty((B, 1, (H + 3) // 4 * 4, (Wd + 3) // 4 * 4), dtype=torch.float32).pin_memory()
    for _ in range(2):
        model.forward_host(x_host, out=y_host, device=local)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.forward_host(x_host, out=y_host, device=local)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = {"value": world * B * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
           "d2h_bytes_per_step": y_host.numel() * 4, "ms_per_step": 1e3 * e2e_s / args.steps,
           "api": "LPSR.forward_host -> lpsr_forward_host (C ABI, pinned host buffers)"}
    checksum = float(y_host.double().mean())

    pre_stats = preprocess_bench(model, dev, with_cpu=not args.no_cpu_baseline) if rank == 0 and world == 1 else None

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_reference_run(args, steps=1000, warmup=1, sample_b=16, budget_s=args.cpu_budget)
        cpu_base = {k: cpu_base[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32"}[args.precision], "data": "synthetic",
                "config": workload_config(args, B, args.precision), "clocks": clocks, "e2e": e2e, "gpu_launches": launches * args.steps,
                "launches_per_step": launches, "roofline": roofline, "roofline_csar": roofline_csar, "roofline_layerwise_hbm": roofline_layerwise,
                "conv_roofline_frac_whole_forward": conv_frac_whole,
                "kernel_ms_per_forward": {k: round(v[0], 4) for k, v in per_kernel.items()},
                "layer_ms_per_forward": {k: round(v, 4) for k, v in fam_ms.items()},
                "cpu_baseline": cpu_base, "preprocess": pre_stats, "output_mean": checksum, "umma": os.environ.get("LPSR_UMMA", "1")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("LPSR_BENCH_PRECISION", "bf16"), choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--batch", type=int, default=1024, help="crops per GPU per step")
    ap.add_argument("--height", type=int, default=64)
    ap.add_argument("--width", type=int, default=192)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        import __graft_entry__ as ge
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            ge.build()
        run_ours(args)


if __name__ == "__main__":
    main()
