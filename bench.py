#!/usr/bin/env python
"""bench.py -- LPSR plate crops/sec on synthetic 3x64x192 crops (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference arithmetic on the box's host cores (torch CPU port)

A step = one LPSR forward over one batch of B synthetic crops per GPU PLUS (N > 1) the one collective of the path, the NCCL
all-gather of the fp32 outputs over NVLink, inside the timed region.  Default = weak scaling (per-GPU batch
fixed at 1024); `--scaling strong` runs BASELINE configs[2] as written (global batch 1024, 1024/N crops per GPU); the
default line also carries the strong-scaling measurement of the same run under `strong_scaling`.  `value` = crops/s with
inputs resident in HBM; `e2e` = the same metric through the C-ABI host-buffer call (lpsr_forward_host: pinned H2D +
forward + D2H inside the timed region).  The default 16-bit mode is fp16: it is the one that meets the 1e-2 bound on
smooth crops (tests/test_gpu_parity.py); bf16 is kept as a switch.  Prints ONE JSON line on rank 0.
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
        self.nvml = None
        self.samples = []          # (sm MHz, max sm MHz, power W, reason bitmask) through NVML: every 10 ms, the first one immediately
        self._stop = threading.Event()

    def _nvml_loop(self, pynvml, handle):
        while True:
            try:
                self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)),
                                     float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)),
                                     pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0,
                                     int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle))))
            except Exception:
                pass
            if self._stop.wait(0.01):
                return

    def start(self):
        # NVML in-process (the same counters nvidia-smi prints) so that even a 30 ms timed region (strong scaling at 8 GPUs) is sampled;
        # nvidia-smi -lms as the fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remaps ordinals: resolve by PCI bus id of the torch device
            try:
                bus = torch.cuda.get_device_properties(self.idx).pci_bus_id
                dom = torch.cuda.get_device_properties(self.idx).pci_domain_id
                dev = torch.cuda.get_device_properties(self.idx).pci_device_id
                handle = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0".encode())
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._nvml_loop, args=(pynvml, handle), daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
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
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=1)
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = [s[0] for s in self.samples]
            reasons = sorted(k for k, b in bits.items() if any(s[3] & b for s in self.samples))
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max((s[1] for s in self.samples), default=None),
                    "power_w_max": max((s[2] for s in self.samples), default=None), "samples": len(sm), "reasons": reasons,
                    "source": "NVML, 10 ms interval during the timed region"}
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
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


def ncu_traffic():
    """DRAM bytes per forward of the tensor-core conv launches / CSAR tail launches from the committed `ncu --set full`
    capture (profiles/r1_ncu_traffic.json, written by tools/ncu_traffic.py); scaled linearly with the batch."""
    for name in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            d = json.load(open(p))
            d["source"] = f"profiles/{name}"
            return d
    return None


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


def workload_config(args, batch, precision, note=None, scaling="weak"):
    gb = batch * args.gpus
    cfg = {"workload": f"LPSR forward {precision}, batch {batch} of 3x{args.height}x{args.width} synthetic crops per GPU "
                       f"(BASELINE.json configs[2]{', global batch 1024 sharded 1024/N' if scaling == 'strong' else ''})",
           "per_gpu_batch": batch, "global_batch": gb,
           "crop": [3, args.height, args.width], "precision": precision, "weights": "reference best_model.pth (fixture)",
           "parallelism": f"batch-sharded x{args.gpus}, replicated weights, no per-layer collective; the one NCCL all-gather of the "
                          f"fp32 outputs is inside the timed region (side stream, overlapping the next forward)",
           "l2": "inputs (151 MB/step at B=1024) and activations (GBs) exceed the 126 MB L2; no explicit flush"}
    if note:
        cfg["note"] = note
    return cfg


def pin_to_gpu_numa(local_rank: int):
    """Pin this process (and the pinned buffers it is about to first-touch) to the CPUs of its GPU's NUMA node: with 8 ranks each
    moving ~200 MB per step between pinned host memory and its GPU, cross-socket traffic is what limited e2e scaling (round 1: 0.88)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return {"numa_node": node, "pinned": False}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "pinned": True, "cpus": len(cpus)}
        return {"numa_node": node, "pinned": False}
    except Exception as exc:   # best effort: single-socket boxes, containers without /sys
        return {"pinned": False, "why": repr(exc)[:80]}


def timed_loop(dev, steps, step_fn, finish_fn, barrier, max_over_ranks):
    """EXACTLY `steps` calls of step_fn bracketed by barrier + synchronize; device time by CUDA events, max over ranks."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step_fn()
    finish_fn()
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1))


def context_baselines(args, dev, model_factory, rank0_weights):
    """Reported context, not targets: the reference arithmetic run by PyTorch eager + cuDNN on this same B200
    (TF32 on, PyTorch's default, and off), this library at BASELINE configs[3] (3x128x384) and at the call-site shape B=1 32x192
    (stream launches and CUDA-graph replay), and the fp32 mode at configs[1]."""
    out = {}
    from oracle import lpsr_torch_port as port   # context baseline = the reference arithmetic through ATen/cuDNN on the GPU

    def ev_time(fn, n, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    try:
        Wg = {k: v.to(dev) for k, v in port.to_torch_weights(rank0_weights).items()}
        xb = torch.rand(256, 3, 64, 192, device=dev)
        x1 = torch.rand(1, 3, 32, 192, device=dev)
        eager = {}
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ms = ev_time(lambda: port.lpsr_forward(xb, Wg), 5)
            ms1 = ev_time(lambda: port.lpsr_forward(x1, Wg), 20)
            eager["tf32_on" if tf32 else "tf32_off"] = {"crops_per_s_b256_64x192": 256 / (ms * 1e-3), "ms_b256": ms, "latency_us_b1_32x192": 1e3 * ms1}
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        eager["what"] = ("reference graph (oracle/lpsr_torch_port.py: the same ATen ops in the same order as my_models/lpsr.py) in PyTorch "
                         f"eager + cuDNN, fp32 NCHW, torch {torch.__version__}, same GPU")
        out["eager_cudnn"] = eager
        del Wg, xb
    except Exception as exc:
        out["eager_cudnn"] = {"unavailable": repr(exc)[:200]}
    pk = peaks()
    try:   # BASELINE configs[3]: 3x128x384 (activation-bandwidth stress), B = 256 (equal pixels to 1024 x 64x192)
        m = model_factory(args.precision)
        x4 = torch.rand(256, 3, 128, 384, device=dev)
        ms = ev_time(lambda: m(x4), 5)
        cps = 256 / (ms * 1e-3)
        out["config4_128x384_b256"] = {"crops_per_s": cps, "ms_per_step": ms, "precision": args.precision,
                                       "conv_roofline_frac_burst": cps * 128 * 384 * DENSE_FLOP_PER_PIXEL / 1e12 / pk["tf_burst"],
                                       "launches": m.launch_count(256, 128, 384)}
        del x4
        # call-site shape: batch 1 of 3x32x192 (inference/run.py:200-202)
        x1 = torch.rand(1, 3, 32, 192, device=dev)
        lat = {"stream_launches_us": 1e3 * ev_time(lambda: m(x1), 50)}
        sgraph = torch.cuda.Stream(dev)
        with torch.cuda.stream(sgraph):
            m(x1)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=sgraph):
            m(x1)
        lat["graph_replay_us"] = 1e3 * ev_time(g.replay, 50)
        lat["launches"] = m.launch_count(1, 32, 192)
        out["latency_b1_32x192"] = lat
        del g
    except Exception as exc:
        out["config4_or_latency_error"] = repr(exc)[:200]
    try:   # BASELINE configs[4]: headless inference/run.py:153-214 on synthetic 1080p frames with injected boxes, LPSR stage timed
        from lpsr_b200 import pipeline as pl
        frames, boxes = pl.synthetic_clip(32, 3, seed=0)
        m = model_factory(args.precision)
        pl.enhance_frames(m, frames[:2], boxes[:2])                       # warm-up
        per_frame, per_clip = pl.StageTimes(), pl.StageTimes()
        for f, b in zip(frames, boxes):                                   # the reference's granularity: one frame (<= 3 plates) at a time
            pl.enhance_frames(m, [f], [b], times=per_frame)
        for _ in range(3):                                                # the whole clip in one call (96 plates)
            pl.enhance_frames(m, frames, boxes, times=per_clip)
        x1 = torch.rand(1, 3, 32, 192)
        t0 = time.perf_counter()
        for _ in range(20):                                               # the reference's own per-plate device round trip (run.py:201-202)
            m(x1.to(dev)).squeeze(0).cpu()
        per_plate_sync_us = 1e6 * (time.perf_counter() - t0) / 20
        out["config5_pipeline"] = {
            "what": "lpsr_b200.pipeline.enhance_frames on 32 synthetic 1080p frames x 3 injected plate boxes (crop 32x192); CUDA-event time "
                    "of pre-processing + forward + uint8 conversion",
            "per_frame_call": {"plates_per_call": 3, "lpsr_stage_us_per_plate": 1e3 * per_frame.lpsr_ms / per_frame.plates,
                               "forward_us_per_plate": 1e3 * per_frame.forward_ms / per_frame.plates,
                               "d2h_us_per_plate": 1e3 * per_frame.d2h_ms / per_frame.plates,
                               "host_us_per_plate": 1e3 * per_frame.host_ms / per_frame.plates},
            "per_clip_call": {"plates_per_call": 96, "lpsr_stage_us_per_plate": 1e3 * per_clip.lpsr_ms / per_clip.plates,
                              "forward_us_per_plate": 1e3 * per_clip.forward_ms / per_clip.plates,
                              "d2h_us_per_plate": 1e3 * per_clip.d2h_ms / per_clip.plates,
                              "host_us_per_plate": 1e3 * per_clip.host_ms / per_clip.plates},
            "reference_style_per_plate_roundtrip_us": per_plate_sync_us,
            "precision": args.precision}
    except Exception as exc:
        out["config5_error"] = repr(exc)[:200]
    try:   # BASELINE configs[1]: fp32 mode, B = 256
        m32 = model_factory("fp32")
        xb = torch.rand(256, 3, 64, 192, device=dev)
        ms = ev_time(lambda: m32(xb), 3, warm=2)
        out["config2_fp32_b256"] = {"crops_per_s": 256 / (ms * 1e-3), "ms_per_step": ms,
                                    "conv_roofline_frac_burst": 256 / (ms * 1e-3) * 64 * 192 * DENSE_FLOP_PER_PIXEL / 1e12 / pk["tf_burst"]}
    except Exception as exc:
        out["config2_error"] = repr(exc)[:200]
    return out


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    import lpsr_b200
    from lpsr_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = pin_to_gpu_numa(local) if world > 1 or args.pin_numa else {"pinned": False, "why": "single process"}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    strong = args.scaling == "strong"
    B_weak = args.batch
    B_strong = max(1, args.batch // world)
    B, H, Wd = (B_strong if strong else B_weak), args.height, args.width
    Hp, Wp = (H + 3) // 4 * 4, (Wd + 3) // 4 * 4
    weights = load_shipped_weights()

    def model_factory(precision):
        m = lpsr_b200.LPSR(3, 32, 16, 4, 4, None, precision=precision)
        m.load_live_weights(weights)
        return m.to(dev).eval()

    model = model_factory(args.precision)
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

    main_stream = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev) if world > 1 else None

    def make_step(xs):
        """forward of this rank's shard + (N > 1) the all-gather of the outputs on a side stream (overlaps the next forward)."""
        bs = xs.shape[0]
        full = torch.empty((world * bs, 1, Hp, Wp), dtype=torch.float32, device=dev) if world > 1 else None
        state = {"y": None}
        # The forward's 28 launches are captured ONCE into a CUDA graph (the caller-side pattern for a launch-bound inner loop; bit-identical
        # to stream launches: tests/test_gpu_parity.py::test_cuda_graph_capture_call_site_batch1) and replayed per step; --no-graph: stream launches.
        graphs = []
        if not args.no_graph:
            try:
                cs = torch.cuda.Stream(dev)
                cs.wait_stream(main_stream)
                with torch.cuda.stream(cs):
                    for _ in range(2):
                        model(xs)
                main_stream.wait_stream(cs)
                # N > 1: two graphs with their own output tensors, so that the gather of step k (side stream) can overlap forward k + 1
                for _ in range(2 if world > 1 else 1):
                    g_ = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g_, stream=cs):
                        y_static = model(xs)
                    graphs.append((g_, y_static))
            except Exception as e:   # capture is an optimisation of the harness, never a requirement
                print(f"[bench] CUDA graph capture failed ({e}); timing stream launches", file=sys.stderr)
                graphs = []
        state["graph"] = bool(graphs)
        gather_done = [None] * len(graphs)
        turn = [0]

        def step():
            b = turn[0]
            if graphs:
                turn[0] = (b + 1) % len(graphs)
                if gather_done[b] is not None:
                    main_stream.wait_event(gather_done[b])      # the gather that read this graph's output two steps ago
                graphs[b][0].replay()
                y = graphs[b][1]
            else:
                y = model(xs)
            if world > 1:
                side.wait_stream(main_stream)
                with torch.cuda.stream(side):
                    dist.all_gather_into_tensor(full, y)
                    if graphs:
                        gather_done[b] = side.record_event()
                if not graphs:
                    y.record_stream(side)
            state["y"] = y

        def finish():
            if world > 1:
                main_stream.wait_stream(side)

        return step, finish, state, full

    # ---------------- device-resident throughput: forward (+ gather) ---------------------------------------------------
    step, finish, state, full = make_step(x)
    for _ in range(max(args.warmup, 3)):
        step()
    finish()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total = timed_loop(dev, args.steps, step, finish, barrier, max_over_ranks)
    clocks = sampler.stop()
    y = state["y"]
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    gather = None
    other_scaling = None
    if world > 1:
        # the collective alone (same buffers): bytes and time reported separately
        yg = y
        def gstep():
            dist.all_gather_into_tensor(full, yg)
        for _ in range(3):
            gstep()
        g_ms = timed_loop(dev, args.steps, gstep, lambda: None, barrier, max_over_ranks) / args.steps
        nb = yg.numel() * 4
        gather = {"collective": "ncclAllGather (torch.distributed.all_gather_into_tensor) of the fp32 [B/G,1,H,W] outputs",
                  "bytes_sent_per_rank": nb, "bytes_received_per_rank": nb * (world - 1), "ms_alone": g_ms,
                  "algbw_GBps": nb * world / (g_ms * 1e-3) / 1e9, "in_timed_region": True, "overlap": "side stream, under the next forward"}
        # the other scaling mode, same process: strong = BASELINE configs[2] as written (global 1024, 1024/N per GPU)
        Bo = B_weak if strong else B_strong
        xo = torch.rand(Bo, 3, H, Wd, generator=torch.Generator().manual_seed(2000 + rank)).to(dev)
        ostep, ofinish, _, _ = make_step(xo)
        for _ in range(3):
            ostep()
        ofinish()
        o_ms = timed_loop(dev, args.steps, ostep, ofinish, barrier, max_over_ranks)
        other_scaling = {"scaling": "weak" if strong else "strong", "per_gpu_batch": Bo, "global_batch": Bo * world,
                         "value": world * Bo * args.steps / (o_ms * 1e-3), "unit": UNIT, "ms_per_step": o_ms / args.steps,
                         "launches_per_step": model.launch_count(Bo, H, Wd), "gather_in_timed_region": True}
        del xo

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
    P = Hp * Wp
    per_kernel = {}
    for name, t in fam_ms.items():
        kind = name.split(":")[1]
        per_kernel.setdefault(kind, [0.0, 0])
        per_kernel[kind][0] += t
        per_kernel[kind][1] += names.count(name)
    # the dominant kernel family: the tcgen05 implicit-GEMM convolutions (umma_conv_kernel instantiations + the fused RDB chain);
    # the CSAR tail kernel (also tcgen05, HBM bound) is reported separately under roofline_csar
    is_conv = lambda n_: n_.split(":")[1].startswith(("umma_conv", "rowconv", "rdb_chain"))
    is_tail = lambda n_: ".tail:" in n_
    conv_ms = sum(t for n_, t in fam_ms.items() if is_conv(n_))
    conv_tags = {n_.split(":")[0] for n_ in names if is_conv(n_)}
    if any(n_.endswith(":rowconv_chain") for n_ in names):
        conv_tags.add("final_conv")      # gff.1 and the final conv are one launch (csrc/rowchain.cuh), tagged rdn.gff1
    # SURVEY 8(d) numerator: DENSE-conv FLOPs only (depthwise and linear arithmetic excluded); the AutoEncoder's pointwise 1x1s
    # are the dense part of its DConvs
    dense_mac = dict(UMMA_MAC_PER_PIXEL)
    dense_mac.update({"ae.enc0": 12 * 12, "ae.enc1": 48 * 12 / 4, "ae.dec0": 48 * 48 / 16, "ae.dec1": 12 * 48 / 4})
    conv_flops = 2.0 * B * P * sum(v for k, v in dense_mac.items() if k in conv_tags and not k.endswith(".tail"))
    n_conv = sum(1 for n_ in names if is_conv(n_))
    esz = 4 if args.precision == "fp32" else 2
    tail_ms = sum(t for n_, t in fam_ms.items() if is_tail(n_))
    n_tail = sum(1 for n_ in names if is_tail(n_))
    tail_bytes = 2.0 * B * P * CSAR_TAIL_ELEMS_PER_PIXEL * esz
    total_prof_ms = sum(fam_ms.values())
    tr = ncu_traffic()
    scale = (B * P) / float(tr["batch"] * tr["pixels_per_crop"]) if tr else 0.0
    if n_conv:
        ach = conv_flops / (conv_ms * 1e-3) / 1e12
        kinds = sorted({n_.split(":")[1] for n_ in names if is_conv(n_)})
        roofline = {"bound": "tensor", "kernel": f"tcgen05 implicit-GEMM convolutions ({', '.join(kinds)}; {n_conv} launches/forward)",
                    "achieved": ach, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": ach / pk["tf_burst"],
                    "peak_kind": f"bf16 dense BURST (timed region {ms_total:.0f} ms), of {pk['src']}",
                    "frac_vs_sustained": ach / pk["tf_sustained"], "peak_sustained": pk["tf_sustained"],
                    "traffic": (tr["umma_dram_bytes_per_forward"] * scale) if tr else None,
                    "traffic_note": (f"sum of dram__bytes_read+write over the family's launches of one forward (ncu --set full, {tr['source']}), "
                                     "scaled to this batch; not measured in this run") if tr else None,
                    "algorithmic_flops_per_forward": conv_flops,
                    "algorithmic_note": "SURVEY 8(d) dense-conv FLOPs of the layers this family executes (CSAR tail 1x1s excluded: roofline_csar)",
                    "kernel_ms_per_forward": conv_ms, "share_of_step": conv_ms / total_prof_ms}
    else:   # fp32 mode: FFMA direct convolution dominates; still reported against the bf16 tensor peak
        d_ms = sum(t for n_, t in fam_ms.items() if n_.endswith(":conv_direct"))
        fl = 2.0 * B * P * (DENSE_FLOP_PER_PIXEL / 2 - 12288)
        ach = fl / (d_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "conv_direct_kernel (FFMA, fp32 parity mode)", "achieved": ach, "peak": pk["tf_burst"],
                    "unit": "TFLOP/s", "frac": ach / pk["tf_burst"], "peak_kind": f"bf16 dense burst, of {pk['src']}",
                    "frac_vs_sustained": ach / pk["tf_sustained"], "traffic": None, "share_of_step": d_ms / total_prof_ms}
    roofline_csar = {"bound": "hbm", "kernel": f"CSAR tail (csar_tail_umma_kernel + channel_gate_kernel: pool finalize + channel gate + spatial MLP + gating + conv_out + residual; {n_tail} launches/forward)",
                     "achieved": tail_bytes / (tail_ms * 1e-3) / 1e9 if tail_ms else None,
                     "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": (tail_bytes / (tail_ms * 1e-3) / 1e9 / pk["hbm_gbs"]) if tail_ms else None,
                     "peak_kind": f"copy bandwidth, of {pk['src']}", "traffic": (tr["tail_dram_bytes_per_forward"] * scale) if tr else None,
                     "algorithmic_bytes_per_forward": tail_bytes, "kernel_ms_per_forward": tail_ms}
    whole = value / world * P * DENSE_FLOP_PER_PIXEL / 1e12
    conv_frac_whole = {"achieved_tflops": whole, "frac_burst": whole / pk["tf_burst"], "frac_sustained": whole / pk["tf_sustained"],
                       "note": "whole-step rate x 297,104 dense FLOP/px (SURVEY 8d) per GPU; the north-star target is 0.50 of burst"}
    lw_bytes = float(B) * P * LAYERWISE_ELEMS_PER_PIXEL * esz
    roofline_layerwise = {"bound": "hbm", "kernel": "whole forward, layer-by-layer activation traffic of the reference graph",
                          "achieved": lw_bytes / (ms_step * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                          "frac": lw_bytes / (ms_step * 1e-3) / 1e9 / pk["hbm_gbs"], "peak_kind": f"copy bandwidth, of {pk['src']}",
                          "algorithmic_bytes_per_forward": lw_bytes,
                          "note": "fused kernels move fewer bytes than this; > 1.0 means the layer-by-layer HBM bound is beaten"}

    # ---------------- end to end through the C-ABI host-buffer call --------------------------------------------------
    y_host = torch.empty((B, 1, Hp, Wp), dtype=torch.float32).pin_memory()
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
           "api": "LPSR.forward_host -> lpsr_forward_host (C ABI, pinned host buffers)", "numa": numa}
    checksum = float(y_host.double().mean())

    extras = rank == 0 and world == 1 and not args.no_extras
    pre_stats = preprocess_bench(model, dev, with_cpu=not args.no_cpu_baseline) if extras else None
    ctx_base = context_baselines(args, dev, model_factory, weights) if extras else None

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_reference_run(args, steps=1000, warmup=1, sample_b=16, budget_s=args.cpu_budget)
        cpu_base = {k: cpu_base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu_base["note"] = "B=16 fp32 on the host cores vs this arm's batch: crops/s is a per-crop rate, so the two are comparable"

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32"}[args.precision], "data": "synthetic",
                "config": workload_config(args, B, args.precision, scaling=args.scaling), "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches * args.steps, "launches_per_step": launches,
                "launch_mode": "CUDA graph replay of the forward's launches (captured once)" if state.get("graph") else "stream launches (programmatic dependent launch)", "roofline": roofline, "roofline_csar": roofline_csar,
                "roofline_layerwise_hbm": roofline_layerwise, "conv_roofline_whole_forward": conv_frac_whole,
                "kernel_ms_per_forward": {k: round(v[0], 4) for k, v in per_kernel.items()},
                "layer_ms_per_forward": {k: round(v, 4) for k, v in fam_ms.items()},
                "gather": gather, ("weak_scaling" if strong else "strong_scaling"): other_scaling,
                "cpu_baseline": cpu_base, "context_baselines": ctx_base, "preprocess": pre_stats, "output_mean": checksum,
                "umma": os.environ.get("LPSR_UMMA", "1")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # fp16 is the 16-bit mode that meets the 1e-2 parity bound on smooth crops with the trained checkpoint (tests/test_gpu_parity.py)
    ap.add_argument("--precision", default=os.environ.get("LPSR_BENCH_PRECISION", "fp16"), choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch crops per GPU; strong: --batch crops in total, batch/N per GPU (BASELINE configs[2] as written)")
    ap.add_argument("--no-graph", action="store_true", help="time stream launches instead of replaying the captured CUDA graph of the forward")
    ap.add_argument("--no-extras", action="store_true", help="skip context baselines / config 4 / B=1 latency / preprocess lines")
    ap.add_argument("--pin-numa", action="store_true", help="pin to the GPU's NUMA node even at N=1")
    ap.add_argument("--batch", type=int, default=1024, help="crops per GPU per step (weak) or in total (strong)")
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
