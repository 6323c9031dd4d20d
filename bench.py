#!/usr/bin/env python
"""
bench.py -- BASELINE.json's metric: 100-iteration ADMM reconstructions / second on synthetic
3x1080x1920 (padded 2160x3840) DiffuserCam-shape measurements, fp32.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our engine
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (oracle port)

A "step" is one complete reconstruction of one measurement per GPU:
set_data -> reset -> 100 x _update -> crop/clamp image.  With N > 1 (launched by torchrun, one
rank per GPU) the batch axis is sharded: every rank reconstructs its own measurement per step
(weak scaling); the PSF is broadcast once over NCCL at init and there is no collective inside
the iteration.  Rank 0 prints ONE JSON line.

value     : whole-job recon/s with the measurement already resident in HBM (device timing,
            CUDA events, max over ranks).
e2e       : same metric through the host-buffer call (pinned host measurement -> H2D ->
            recon -> D2H image) with the copies inside the timed region.
roofline  : our dominant kernel (the longest launch of the iteration: the fused spatial-update + row-FFT
            kernel B2), algorithmic bytes / CUDA-event duration measured live via lpc_profile(),
            against MEASURED_PEAKS.json.
cpu_baseline : the numpy oracle port of the reference, timed on this box's host cores on a
            bounded sample (a few iterations at full size, scaled to 100).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ADMM-100 reconstructions/sec @ 3x1080x1920"
UNIT = "recon/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="admm", choices=["admm", "fista"])
    ap.add_argument("--n-iter", type=int, default=None)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--channels", type=int, default=3)
    ap.add_argument("--batch-per-gpu", type=int, default=1)
    ap.add_argument("--depth", type=int, default=1, help="PSF depth D > 1: BASELINE configs[4], D independent 2-D problems "
                    "against one measurement (ADMM(allow_depth=True); the reference has no 3-D ADMM, SURVEY 8a note)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--engine", default=None, choices=[None, "v1"], help="v1 = K1/K2 + cuFFT (A/B against the fused engine)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(smax) if smax else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------- reference arm
def run_reference(args, out):
    """The reference's own CPU implementation of the path: the numpy oracle port (the reference is
    pure Python and cannot travel to the GPU box), all host threads given to pocketfft.
    One step = ONE full-size iteration (a bounded sample); value scales it to n_iter."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lensless_oracle as orc

    n_iter = args.n_iter or (100 if args.algo == "admm" else 300)
    cores = os.cpu_count() or 1
    h, w, C = args.height, args.width, args.channels
    rng = np.random.default_rng(0)
    psf = (rng.random((1, h, w, C)) ** 8).astype(np.float32)
    psf /= np.linalg.norm(psf.ravel())
    data = rng.random((1, h, w, C)).astype(np.float32)
    solver = orc.ALGOS[args.algo](psf, dtype="float32", workers=cores)
    solver.set_data(data)
    solver.reset()
    for _ in range(args.warmup):
        solver.update()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solver.update()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = 1.0 / (dt * n_iter)
    sample = f"{args.steps} timed full-size iterations after {args.warmup} warm-up, scaled x{n_iter} (per-iteration cost is constant)"
    line = {
        "impl": "reference",
        "metric": METRIC if args.algo == "admm" else f"FISTA-{n_iter} reconstructions/sec @ 3x1080x1920",
        "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * n_iter * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.algo.upper()} n_iter={n_iter} fp32, single {C}x{h}x{w} measurement (BASELINE configs[1]), numpy+scipy.fft oracle port of lensless/recon"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.emit(json.dumps(line))


def cpu_baseline(args, n_iter):
    """numpy oracle AS SHIPPED by the reference (scipy.fft with 1 thread), bounded sample."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lensless_oracle as orc

    h, w, C = args.height, args.width, args.channels
    rng = np.random.default_rng(0)
    psf = (rng.random((1, h, w, C)) ** 8).astype(np.float32)
    psf /= np.linalg.norm(psf.ravel())
    data = rng.random((1, h, w, C)).astype(np.float32)
    solver = orc.ALGOS[args.algo](psf, dtype="float32", workers=None)
    solver.set_data(data)
    solver.reset()
    solver.update()
    k = 3
    t0 = time.perf_counter()
    for _ in range(k):
        solver.update()
    dt = (time.perf_counter() - t0) / k
    return {
        "value": 1.0 / (dt * n_iter), "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"{k} timed full-size iterations after 1 warm-up, scaled x{n_iter}; numpy + scipy.fft (pocketfft, 1 thread as the reference ships it); host has {os.cpu_count()} cores",
        "s_per_iter": dt,
    }


def torch_baseline_gpu(args, n_iter, device, psf, meas):
    """The reference's TORCH path on the SAME GPU ('the real bar for the fused
    kernels'): oracle/lensless_oracle_torch.py restates lensless/recon/{admm,gd}.py's torch branch operation for
    operation (pinned against the reference's torch outputs in tests/); timed with explicit synchronisation on a
    bounded sample and scaled to n_iter."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lensless_oracle_torch as ot

    solver = ot.ALGOS[args.algo](psf)
    solver.set_data(meas[:1])
    solver.reset()
    for _ in range(3):
        solver.update()
    k = 10
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        solver.update()
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / k
    del solver
    torch.cuda.empty_cache()
    return {
        "value": 1.0 / (ms * 1e-3 * n_iter), "unit": UNIT, "kind": "port (torch ops of the reference's torch branch, same GPU)",
        "ms_per_iter": ms, "sample": f"{k} timed full-size iterations after 3 warm-up, CUDA events + synchronize, scaled x{n_iter}",
    }


# ----------------------------------------------------------------------------------------- stdout hygiene
class OneLineStdout:
    """Rank 0 must print exactly ONE JSON line.  Libraries write banners to fd 1 (NCCL prints its version there
    when NCCL_DEBUG is set on the box), so everything but the final line is sent to stderr."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


# ----------------------------------------------------------------------------------------- our arm
def main():
    args = parse()
    out = OneLineStdout()
    if args.impl == "reference":
        return run_reference(args, out)

    import ctypes

    import torch
    import torch.distributed as dist

    import lenslesspicam_b200 as L
    from lenslesspicam_b200 import _lib
    from lenslesspicam_b200 import dist as lpd

    world, rank, local = lpd.env_world()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    lpd.init(backend="nccl", device=device)

    n_iter = args.n_iter or (100 if args.algo == "admm" else 300)
    h, w, C, Bg = args.height, args.width, args.channels, args.batch_per_gpu

    # ---- synthetic inputs: PSF on rank 0, broadcast once over NCCL; per-rank scenes
    D = args.depth
    gen = torch.Generator().manual_seed(0)
    psf = torch.rand((D, h, w, C), generator=gen) ** 8
    psf = (psf / psf.norm()).to(device)
    if world > 1 and rank != 0:
        psf.zero_()
    lpd.broadcast_psf(psf, src=0)  # the only collective of the job: init-time, NCCL over NVLink/NVSwitch
    gen = torch.Generator().manual_seed(1000 + rank)
    scene = (torch.rand((Bg, 1, h, w, C), generator=gen) ** 2).to(device)
    if D == 1:
        sim = L.RealFFTConvolve2D(psf, pad=True, norm="backward")
        meas = sim.convolve(scene).clamp_(min=0)
        meas = meas / meas.amax(dim=(1, 2, 3, 4), keepdim=True)
        del sim
    else:
        meas = scene  # the arithmetic is data-independent except clamps / thresholds: any measurement times the same
    del scene
    host_in = meas.cpu().pin_memory()
    host_out = torch.empty((Bg, D, h, w, C), dtype=torch.float32).pin_memory()

    cls = L.ADMM if args.algo == "admm" else L.FISTA
    stream = torch.cuda.Stream(device=device)
    with torch.cuda.stream(stream):
        extra = {"allow_depth": True} if (D > 1 and args.algo == "admm") else {}
        recon = cls(psf, n_iter=n_iter, no_graph=args.no_graph, engine=args.engine, **extra)

        def step_device():
            recon.set_data(meas)
            return recon.apply(disp_iter=None, plot=False)

        def step_host():
            recon.apply_host(host_in, host_out, n_iter=n_iter)
            stream.synchronize()  # the caller reads the image: a real end-to-end step ends on the host

        def barrier():
            lpd.barrier(device)

        def timed(fn, K):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(K):
                fn()
            e1.record(stream)
            barrier()
            return lpd.max_over_ranks(e0.elapsed_time(e1), device=device)

        for _ in range(max(args.warmup, 3)):
            step_device()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms_dev = timed(step_device, args.steps)
        for _ in range(2):
            step_host()
        ms_e2e_sync = timed(step_host, args.steps)  # one measurement at a time, host synchronised after each
        # the streaming call: the same per-measurement work (H2D, reset, n_iter iterations, crop, D2H), copies on side
        # streams against double-buffered staging so they overlap the neighbouring measurements' iterations
        host_ins = [host_in] * args.steps
        host_outs = [host_out] + [torch.empty_like(host_out).pin_memory() for _ in range(min(args.steps, 4) - 1)]
        host_outs = [host_outs[k % len(host_outs)] for k in range(args.steps)]
        recon.apply_host_stream(host_ins[:2], host_outs[:2], n_iter=n_iter)
        ms_e2e = timed(lambda: recon.apply_host_stream(host_ins, host_outs, n_iter=n_iter), 1)
        clocks = sampler.stop() if rank == 0 else None

        # ---- per-phase device times of one iteration, CUDA events on the launching stream
        recon.set_data(meas)
        recon.apply(n_iter=2, disp_iter=None)
        phase_ms = (ctypes.c_double * 8)()
        nph = ctypes.c_int()
        lib = _lib.load()
        _lib.check(lib.lpc_profile(recon._plan, 10, stream.cuda_stream, phase_ms, ctypes.byref(nph)), recon._plan)
        algo_code = _lib.ALGO_ADMM if args.algo == "admm" else _lib.ALGO_FISTA
        phases = {}
        for i in range(nph.value):
            name = lib.lpc_phase_name(recon._plan, i).decode()
            phases[name] = phases.get(name, 0.0) + phase_ms[i]
        bytes_iter = recon.bytes_per_iter()

    if rank != 0:
        lpd.finalize()
        return

    total = world * Bg
    ms_step = ms_dev / args.steps
    value = total * args.steps / (ms_dev / 1e3)
    e2e_value = total * args.steps / (ms_e2e / 1e3)
    peak, peak_src = measured_peaks()

    _, Hp, Wp, _ = recon._padded_shape
    P = 4.0 * Bg * D * Hp * Wp * C
    Sp = 8.0 * Bg * D * Hp * (Wp // 2 + 1) * C
    d = 4.0 * Bg * D * h * w * C
    # algorithmic bytes of OUR kernels, per launch
    kernel_bytes = {
        "admm_spatial(K1)": 11.0 * P + d,        # v1: reads V,HV,xi',rho',eta'0,eta'1 (+y); writes xi',rho',eta'0,eta'1,a
        "admm_spectral(K2)": 5.5 * Sp,           # v1: reads A,Xi,H,R; writes Vhat,HVhat
        "rows_ifft(B1)": 2.0 * Sp + 2.0 * P,             # v2: reads Tin(2Sp); writes V,HV
        "rows_spatial_fft(B2)": 10.0 * P + 2.0 * Sp + d,  # v2: reads V,HV,xi',rho',eta'(2P) (+y); writes xi',rho',eta'(2P),Tout(2Sp)
        "admm_cols(A)": 5.0 * Sp,                # v2: reads Tout(2Sp),H; writes Tin(2Sp)  (R is evaluated on the fly)
        "update": 5.0 * d, "residual": 3.0 * d, "spec_mul_H": 3.0 * Sp, "spec_mul_Hadj": 3.0 * Sp,
    }
    # fused GD engine: G-layout buffer (window rows only) and the filter table, per launch
    Tg = 8.0 * Bg * D * (-(-(Wp // 2 + 1) // 4) * 4) * h * C
    Fg = 8.0 * D * (-(-(Wp // 2 + 1) // 4) * 4) * Hp * C
    kernel_bytes.update({
        "gd_cols(H)": 2.0 * Tg + Fg, "gd_cols(Hadj)": 2.0 * Tg + Fg,
        "gd_rows(residual)": 2.0 * Tg + d, "gd_rows(update)": 2.0 * Tg + 4.0 * d,
    })
    own = {n: t for n, t in phases.items() if n in kernel_bytes}
    kname = max(own, key=own.get)  # our dominant kernel = the one with the longest launch
    kbytes = kernel_bytes[kname]
    engine = lib.lpc_engine_version(recon._plan)
    if args.algo == "admm":
        own_launches = (3 if engine == 2 else 2) * n_iter + 2  # kernels / iteration, pack + unpack per recon
    else:
        own_launches = 4 * n_iter + 3
    k_ms = phases.get(kname, 0.0)
    achieved = kbytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    iter_ms = sum(phases.values())
    traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    try:
        if (h, w, C, Bg) == (1080, 1920, 3, 1):
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kname)
    except Exception:
        traffic = None
    roofline = {
        "bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": kbytes, "launch_ms": k_ms,
        "phases_ms_per_iter": phases,
        "phases_GBs": {n: kernel_bytes[n] / (t * 1e-3) / 1e9 for n, t in own.items() if t > 0},
        "engine": "v2 fused own-FFT (3 kernels / iteration)" if engine == 2 else "v1 K1/K2 + cuFFT",
        "iteration": {
            "algorithmic_bytes": bytes_iter, "ms_ungraphed": iter_ms, "ms_in_step": ms_step / n_iter / Bg,
            "achieved_GBs": bytes_iter / (ms_step / n_iter * 1e-3) / 1e9,
            "frac_of_peak": bytes_iter / (ms_step / n_iter * 1e-3) / 1e9 / peak,
        },
    }
    line = {
        "metric": (METRIC if (args.algo == "admm" and D == 1 and n_iter == 100 and (h, w, C) == (1080, 1920, 3)) else
                   f"{args.algo.upper()}-{n_iter} reconstructions/sec @ {str(D) + 'x' if D > 1 else ''}{C}x{h}x{w}"),
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": (f"{args.algo.upper()} n_iter={n_iter} fp32, single {C}x{h}x{w} measurement per GPU per step (BASELINE configs[1]), padded {Hp}x{Wp}"
                         if D == 1 else
                         f"{args.algo.upper()} n_iter={n_iter} fp32, {D}-depth PSF stack x {C}x{h}x{w}, one measurement per GPU per step (BASELINE configs[4]: "
                         f"{D} independent 2-D problems against the same measurement -- the reference has no 3-D ADMM), padded {Hp}x{Wp}"),
            "batch_per_gpu": Bg, "parallelism": f"batch-sharded x{world}, PSF broadcast once (NCCL), no collective in the iteration",
            "l2": "per-recon working set ~1.4 GB >> 126 MB L2 (inputs larger than L2; no explicit flush)",
            "cuda_graph": not args.no_graph,
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(4 * Bg * h * w * C), "d2h_bytes_per_step": int(d), "ms_per_step": ms_e2e / args.steps,
                "call": "apply_host_stream: K measurements, every one copied H2D and read back D2H inside the timed region, copies overlapped with the neighbours' iterations",
                "sync_per_step": {"value": total * args.steps / (ms_e2e_sync / 1e3), "ms_per_step": ms_e2e_sync / args.steps,
                                  "call": "apply_host + stream.synchronize() per measurement"}},
        "gpu_launches": own_launches * args.steps,  # our kernels inside the `value` timed region (cuFFT kernels not counted)
        "roofline": roofline,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, n_iter)
        if D == 1:
            try:
                line["torch_baseline_gpu"] = torch_baseline_gpu(args, n_iter, device, psf, meas)
            except Exception as e:  # never lose the bench line to the extra baseline
                line["torch_baseline_gpu"] = {"unavailable": repr(e)[:200]}
    out.emit(json.dumps(line))
    lpd.finalize()


if __name__ == "__main__":
    main()
