#!/usr/bin/env python
"""bench.py -- BASELINE.json headline benchmark: wavedec2 db4 level 4 on a batch of 4096x4096 float32.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one multi-level forward transform of one batch of synthetic images (per GPU).
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

WAVELET, LEVEL, MODE, H, W = "db4", 4, "reflect", 4096, 4096
METRIC = "Msamples/s, wavedec2 db4 L4 4096x4096 fp32 (forward)"


def coeff_sizes(n: int, filt_len: int, levels: int):
    out = []
    for _ in range(levels):
        n = (n + filt_len - 1) // 2
        out.append(n)
    return out


def algorithmic_bytes_per_image() -> int:
    """Input read once + every returned coefficient written once: 134 612 360 B."""
    sz = coeff_sizes(H, 8, LEVEL)
    coeffs = sz[-1] ** 2 + 3 * sum(s * s for s in sz)
    return 4 * (H * W + coeffs)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference_throughput(sample_batch: int, reps: int):
    """The oracle port (the reference's own torch-CPU operator sequence) on the host cores."""
    from oracle import ptwt_port as P

    ncpu = os.cpu_count() or 1
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(sample_batch, H, W, generator=g, dtype=torch.float32)
    best, best_cores, times = float("inf"), ncpu, []
    # torch's CPU convolution does not always scale to every hardware thread: give the reference its
    # best thread count among {all, half (physical cores), 32}
    for cores in sorted({ncpu, max(ncpu // 2, 1), min(32, ncpu)}, reverse=True):
        torch.set_num_threads(cores)
        P.wavedec2(x[:1], WAVELET, mode=MODE, level=LEVEL)  # warm-up (oneDNN primitive creation)
        for _ in range(reps):
            t0 = time.perf_counter()
            P.wavedec2(x, WAVELET, mode=MODE, level=LEVEL)
            dt = time.perf_counter() - t0
            times.append(dt)
            if dt < best:
                best, best_cores = dt, cores
    return sample_batch * H * W / best / 1e6, best_cores, times


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_batch
    steps, warmup = max(args.steps, 1), args.warmup
    from oracle import ptwt_port as P

    ncpu = os.cpu_count() or 1
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(sample, H, W, generator=g, dtype=torch.float32)
    # pick the thread count the reference's operators run fastest with (see cpu_reference_throughput)
    cores, best = ncpu, float("inf")
    for cand in sorted({ncpu, max(ncpu // 2, 1), min(32, ncpu)}, reverse=True):
        torch.set_num_threads(cand)
        P.wavedec2(x[:1], WAVELET, mode=MODE, level=LEVEL)
        t0 = time.perf_counter()
        P.wavedec2(x, WAVELET, mode=MODE, level=LEVEL)
        dt = time.perf_counter() - t0
        if dt < best:
            best, cores = dt, cand
    torch.set_num_threads(cores)
    for _ in range(max(warmup, 1)):
        P.wavedec2(x, WAVELET, mode=MODE, level=LEVEL)
    t0 = time.perf_counter()
    for _ in range(steps):
        P.wavedec2(x, WAVELET, mode=MODE, level=LEVEL)
    dt = (time.perf_counter() - t0) / steps
    val = sample * H * W / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (torch.randn, seed 1234)",
        "config": {"workload": f"wavedec2 {WAVELET} level={LEVEL} mode={MODE}, {sample}x{H}x{W} float32 per step "
                               f"(bounded sample of the 64-image batch) on the host CPU"},
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} images of {H}x{W} per step, oracle/ptwt_port.py "
                                   "(F.pad + F.conv2d(stride=2) + split, the reference's operator sequence)"},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step (BASELINE config: 64)")
    ap.add_argument("--cpu-batch", type=int, default=8, help="images in the CPU-baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-incumbent", action="store_true", help="skip timing the reference algorithm on the GPU")
    ap.add_argument("--incumbent-batch", type=int, default=16)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import pytorch_wavelet_toolbox_b200 as wt
    from pytorch_wavelet_toolbox_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    B = args.batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(B, H, W, generator=g, device=dev, dtype=torch.float32)

    def step():
        return wt.wavedec2(x, WAVELET, mode=MODE, level=LEVEL)

    for _ in range(max(args.warmup, 3)):
        out = step()
    del out
    torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    _native.launch_count_reset()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        ev[i][0].record()
        out = step()
        ev[i][1].record()
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = _native.launch_count()
    total_ms = t_start.elapsed_time(t_end)
    step_ms = sorted(a.elapsed_time(b) for a, b in ev)
    if dist is not None:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * B * H * W / (ms_per_step * 1e-3) / 1e6

    # parity spot check of what was just timed (one image, against the oracle) -- outside the timed region
    max_err = None
    if rank == 0:
        from oracle import ptwt_port as P

        want = P.wavedec2(x[:1].cpu(), WAVELET, mode=MODE, level=LEVEL)
        flat_w = [want[0]] + [b for lv in want[1:] for b in lv]
        flat_g = [out[0]] + [b for lv in out[1:] for b in lv]
        scale = max(float(t.abs().max()) for t in flat_w)
        max_err = max(float((a[:1].cpu() - b).abs().max()) for a, b in zip(flat_g, flat_w)) / scale
    del out

    # roofline: the dominant kernel is the level-1 launch of fwd2d_strip_f32_kernel (72 % of the step);
    # it is timed alone with CUDA events (a level=1 transform is exactly that one launch), achieved =
    # its algorithmic bytes 4*(H*W + 4*Mh*Mw) per image / its average duration.  The whole step
    # (4 launches, approximation bands re-read between levels) is reported as step_*.
    peak, peak_src = measured_peak_gbs()
    alg = algorithmic_bytes_per_image() * B
    med_ms = step_ms[len(step_ms) // 2]
    step_achieved = alg / (med_ms * 1e-3) / 1e9
    m1 = coeff_sizes(H, 8, 1)[0]
    alg_k = 4 * (H * W + 4 * m1 * m1) * B
    for _ in range(3):
        wt.wavedec2(x, WAVELET, mode=MODE, level=1)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    _native.launch_count_reset()
    k0.record()
    for _ in range(20):
        wt.wavedec2(x, WAVELET, mode=MODE, level=1)
    k1.record()
    torch.cuda.synchronize(dev)
    k_launches = _native.launch_count()
    k_ms = k0.elapsed_time(k1) / 20
    achieved = alg_k / (k_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("dominant_kernel_dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None

    # inverse transform of the same coefficients (reported separately)
    inverse = None
    if rank == 0 or True:
        coeffs = step()
        for _ in range(3):
            rec = wt.waverec2(coeffs, WAVELET)
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        i0.record()
        for _ in range(10):
            rec = wt.waverec2(coeffs, WAVELET)
        i1.record()
        torch.cuda.synchronize(dev)
        inv_ms = i0.elapsed_time(i1) / 10
        rt_err = float((rec[:2] - x[:2]).abs().max())
        inverse = {"ms_per_step": inv_ms, "value": B * H * W / (inv_ms * 1e-3) / 1e6, "unit": "Msamples/s (per GPU)",
                   "step_frac": alg / (inv_ms * 1e-3) / 1e9 / peak, "round_trip_max_abs_err": rt_err}
        del coeffs, rec

    # end to end through the public API with HOST (pinned) buffers: H2D + transform + D2H every step
    e2e = None
    if not args.no_e2e:
        xh = torch.empty((B, H, W), dtype=torch.float32, pin_memory=True)
        xh.copy_(x)
        for _ in range(2):
            oh = wt.wavedec2(xh, WAVELET, mode=MODE, level=LEVEL)
        d2h = sum(t.numel() for t in [oh[0]] + [b for lv in oh[1:] for b in lv]) * 4
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            oh = wt.wavedec2(xh, WAVELET, mode=MODE, level=LEVEL)
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / args.e2e_steps
        if dist is not None:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * B * H * W / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": B * H * W * 4,
               "d2h_bytes_per_step": d2h, "ms_per_step": dt * 1e3}
        del oh, xh

    # the incumbent on this GPU: the reference's own algorithm (F.pad -> F.conv2d(stride 2) with the four
    # outer-product filters, per level) executed by torch/cuDNN on the same device -- what ptwt does today
    # when it is handed CUDA tensors.  Informational: a sample of the batch, device-resident, CUDA events.
    incumbent = None
    if rank == 0 and not args.no_incumbent:
        try:
            sys.path.insert(0, str(Path(__file__).resolve().parent))
            from oracle import ptwt_port as P

            nb = min(args.incumbent_batch, B)
            xs = x[:nb]
            for _ in range(2):
                ref_c = P.wavedec2(xs, WAVELET, mode=MODE, level=LEVEL)
            torch.cuda.synchronize(dev)
            j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            j0.record()
            for _ in range(3):
                ref_c = P.wavedec2(xs, WAVELET, mode=MODE, level=LEVEL)
            j1.record()
            torch.cuda.synchronize(dev)
            inc_ms = j0.elapsed_time(j1) / 3
            inc_alg = nb * algorithmic_bytes_per_image()
            incumbent = {"value": nb * H * W / (inc_ms * 1e-3) / 1e6, "unit": "Msamples/s (1 GPU)", "ms": inc_ms,
                         "sample": f"{nb} images of {H}x{W}, device resident",
                         "step_frac": inc_alg / (inc_ms * 1e-3) / 1e9 / peak,
                         "what": "reference algorithm (F.pad + F.conv2d stride 2, torch/cuDNN) on the same B200"}
            del ref_c, xs
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001
            incumbent = {"unavailable": f"{type(ex).__name__}: {str(ex)[:160]}"}

    cpu = None
    if rank == 0 and not args.no_cpu:
        v, cores, times = cpu_reference_throughput(args.cpu_batch, 3)
        cpu = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_batch} images of {H}x{W}, best of 3 ({', '.join(f'{t:.3f}s' for t in times)}), "
                         "oracle/ptwt_port.py = the reference's F.pad + F.conv2d(stride=2) sequence on all host threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (torch.randn on device, seed 1234+rank)",
            "config": {"workload": f"wavedec2 {WAVELET} level={LEVEL} mode={MODE}, batch {B} x {H}x{W} float32 per GPU "
                                   "(BASELINE.json configs[1])",
                       "l2": "inputs (4.29 GB) and outputs (4.32 GB) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"batch-sharded x{world}, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "fwd2d_strip_f32_kernel<8,64,TMA> (level-1 launch)",
                         "kernel_algorithmic_bytes_per_launch": alg_k, "kernel_ms_per_launch": k_ms,
                         "kernel_launches_timed": int(k_launches),
                         "step_achieved": step_achieved, "step_frac": step_achieved / peak,
                         "algorithmic_bytes_per_step": alg, "median_step_ms": med_ms, "min_step_ms": step_ms[0]},
            "cpu_baseline": cpu, "e2e": e2e, "inverse": inverse, "incumbent_gpu": incumbent, "gpu_launches": int(launches), "clocks": clocks,
            "parity": {"max_rel_err_vs_oracle": max_err, "tolerance": 1e-5},
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
