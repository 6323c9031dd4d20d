#!/usr/bin/env python
"""
bench.py -- BASELINE.json's metric: 100-iteration ADMM reconstructions / second on synthetic
3x1080x1920 (padded 2160x3840) DiffuserCam-shape measurements, fp32.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our engine
    python bench.py --impl reference [--steps K] [--warmup W]      # the UNMODIFIED reference's CPU path (baseline/_ref)

A "step" of our arm is one complete reconstruction of one measurement per GPU:
set_data -> reset -> 100 x _update -> crop/clamp image.  With N > 1 (launched by torchrun, one
rank per GPU) the batch axis is sharded: every rank reconstructs its own measurement per step
(weak scaling); the PSF is broadcast once over NCCL at init and there is no collective inside
the iteration.  Rank 0 prints ONE JSON line.

value     : whole-job recon/s with the measurement already resident in HBM (device timing,
            CUDA events, max over ranks).
e2e       : same metric through the host-buffer call (pinned host measurement -> H2D ->
            recon -> D2H image) with the copies inside the timed region.
roofline  : our dominant kernel (the longest launch of the iteration), algorithmic bytes / CUDA-event
            duration measured live via lpc_profile(), against MEASURED_PEAKS.json.
configs   : the other BASELINE.json configurations, each timed the same way in the same run:
            fista300 (configs[2]), bundled_ds8_admm5 / joss_ds4_gray (configs[0] and the JOSS paper's table),
            depth16_admm50 (configs[4]), batch256 (configs[3]: 256 host-resident measurements through
            lenslesspicam_b200.dist.reconstruct_sharded, time-to-solution -> strong scaling over N).
parity    : N = 1 only -- ours vs the reference's torch path run on the same GPU in fp64 (truth) and fp32
            (trajectory) for ADMM-100 and FISTA-300 at the headline shape (checker, not measured).
cpu_baseline / torch_baseline_gpu : the unmodified reference (baseline/_ref) on this box's host cores (numpy path as
            shipped, bounded sample) and on the same GPU (torch path, explicit synchronisation); the oracle port
            stands in only when baseline/_ref is absent.

The reference arm (--impl reference): one step = ONE full-size iteration (`_update`, recon.py:575-576) of the
unmodified reference's numpy path with every host thread given to scipy.fft; `ms_per_step` is that iteration's
real duration and `value` scales it to n_iter (per-iteration cost is constant).
"""
import argparse
import contextlib
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
    ap.add_argument("--configs", default="default", help="'default' = every BASELINE config next to the headline when the "
                    "headline shape is benchmarked, 'none', or a comma list of fista300,bundled,depth16,raw,batch256")
    ap.add_argument("--batch256-n", type=int, default=256, help="measurements of the sharded-batch config (BASELINE configs[3])")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip cpu_baseline / torch_baseline_gpu / parity")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--engine", default=None, choices=[None, "v1"], help="v1 = K1/K2 + cuFFT (A/B against the fused engine)")
    return ap.parse_args()


def next_fast_len(n):
    """smallest 5-smooth integer >= n (scipy.fftpack.next_fast_len, rfft_convolve.py:16,112)"""
    best = None
    p5 = 1
    while p5 < 2 * n:
        p35 = p5
        while p35 < 2 * n:
            c = p35
            while c < n:
                c *= 2
            best = c if best is None or c < best else best
            p35 *= 3
        p5 *= 5
    return best if n > 6 else max(n, 0)


def workload(algo, n_iter, C, h, w, D=1):
    """config.workload -- the SAME string for both arms (the driver compares them)"""
    Hp, Wp = next_fast_len(2 * h - 1), next_fast_len(2 * w - 1)
    if D == 1:
        return (f"{algo.upper()} n_iter={n_iter} fp32, single {C}x{h}x{w} measurement per GPU per step "
                f"(BASELINE configs[{1 if algo == 'admm' else 2}]), padded {Hp}x{Wp}")
    return (f"{algo.upper()} n_iter={n_iter} fp32, {D}-depth PSF stack x {C}x{h}x{w}, one measurement per GPU per step (BASELINE "
            f"configs[4]: {D} independent 2-D problems against the same measurement -- the reference has no 3-D ADMM), padded {Hp}x{Wp}")


def metric_name(algo, n_iter, C, h, w, D=1):
    if algo == "admm" and D == 1 and n_iter == 100 and (h, w, C) == (1080, 1920, 3):
        return METRIC
    return f"{algo.upper()}-{n_iter} reconstructions/sec @ {str(D) + 'x' if D > 1 else ''}{C}x{h}x{w}"


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


def load_reference():
    """the UNMODIFIED reference from baseline/_ref (None when it was not installed)"""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import ref_loader

        return ref_loader.load()
    except Exception:
        return None


def synthetic_numpy(h, w, C, seed=0):
    import numpy as np

    rng = np.random.default_rng(seed)
    psf = (rng.random((1, h, w, C)) ** 8).astype(np.float32)
    psf /= np.linalg.norm(psf.ravel())
    data = rng.random((1, h, w, C)).astype(np.float32)
    return psf, data


# ----------------------------------------------------------------------------------------- reference arm
def run_reference(args, out):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 works; the others exit 0
    n_iter = args.n_iter or (100 if args.algo == "admm" else 300)
    cores = os.cpu_count() or 1
    h, w, C = args.height, args.width, args.channels
    psf, data = synthetic_numpy(h, w, C)
    ref = load_reference()
    if ref is not None:
        from scipy import fft as sfft

        kind = "reference"
        cls = ref.ADMM if args.algo == "admm" else ref.FISTA
        solver = cls(psf, dtype="float32")
        solver.set_data(data)
        ctx = sfft.set_workers(cores)  # scipy.fft's own global knob: the reference code itself is untouched
        step = lambda i: solver._update(i)  # noqa: E731  (the loop body of apply(), recon.py:575-576)
        what = "unmodified reference (baseline/_ref) numpy path, scipy.fft.set_workers(all cores)"
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import lensless_oracle as orc

        kind = "port"
        solver = orc.ALGOS[args.algo](psf, dtype="float32", workers=cores)
        solver.set_data(data)
        solver.reset()
        ctx = contextlib.nullcontext()
        step = lambda i: solver.update()  # noqa: E731
        what = "numpy + scipy.fft oracle port (baseline/_ref absent), all cores"
    with ctx:
        for i in range(args.warmup):
            step(i)
        t0 = time.perf_counter()
        for i in range(args.steps):
            step(args.warmup + i)
        dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = 1.0 / (dt * n_iter)
    sample = (f"one step = ONE full-size iteration ({what}); {args.steps} timed after {args.warmup} warm-up; "
              f"value = 1 / (ms_per_step x {n_iter} iterations): the per-iteration cost is constant")
    extra = {}
    if ref is not None and args.algo in ("admm", "fista"):
        try:  # the reference's torch-CPU path, every intra-op thread (the faster of its two CPU paths on most hosts)
            import torch

            torch.set_num_threads(cores)
            cls = ref.ADMM if args.algo == "admm" else ref.FISTA
            ts = cls(torch.from_numpy(psf), dtype="float32")
            ts.set_data(torch.from_numpy(data))
            ts._update(0)
            t0 = time.perf_counter()
            for i in range(3):
                ts._update(1 + i)
            dtt = (time.perf_counter() - t0) / 3
            extra["torch_cpu"] = {"value": 1.0 / (dtt * n_iter), "unit": UNIT, "s_per_iter": dtt, "threads": cores,
                                  "sample": "3 timed full-size iterations after 1 warm-up, unmodified reference torch-CPU path"}
        except Exception as e:
            extra["torch_cpu"] = {"unavailable": repr(e)[:200]}
    line = {
        "impl": "reference",
        "metric": metric_name(args.algo, n_iter, C, h, w),
        "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(args.algo, n_iter, C, h, w), "step": "one iteration of the workload (bounded sample)",
                   "iterations_per_reconstruction": n_iter},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "s_per_iter": dt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        **extra,
    }
    out.emit(json.dumps(line))


def cpu_baseline(args, n_iter):
    """The reference's numpy path AS SHIPPED (scipy.fft with its default single thread), bounded sample."""
    h, w, C = args.height, args.width, args.channels
    psf, data = synthetic_numpy(h, w, C)
    ref = load_reference()
    if ref is not None:
        cls = ref.ADMM if args.algo == "admm" else ref.FISTA
        solver = cls(psf, dtype="float32")
        solver.set_data(data)
        step, kind = (lambda i: solver._update(i)), "reference"
        what = "unmodified reference (baseline/_ref), numpy + scipy.fft as shipped (1 FFT thread)"
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import lensless_oracle as orc

        solver = orc.ALGOS[args.algo](psf, dtype="float32", workers=None)
        solver.set_data(data)
        solver.reset()
        step, kind = (lambda i: solver.update()), "port"
        what = "numpy + scipy.fft oracle port (1 FFT thread)"
    step(0)
    k = 3
    t0 = time.perf_counter()
    for i in range(k):
        step(1 + i)
    dt = (time.perf_counter() - t0) / k
    return {
        "value": 1.0 / (dt * n_iter), "unit": UNIT, "cores": 1, "kind": kind,
        "sample": f"{k} timed full-size iterations after 1 warm-up, scaled x{n_iter}; {what}; host has {os.cpu_count()} cores "
                  "(the all-cores figure is the --impl reference arm)",
        "s_per_iter": dt,
    }


def reference_gpu_solver(algo, psf, dtype):
    """the reference's torch path on the GPU: the unmodified class from baseline/_ref, else the restatement"""
    import torch

    ref = load_reference()
    tdt = torch.float32 if dtype == "float32" else torch.float64
    if ref is not None:
        cls = ref.ADMM if algo == "admm" else ref.FISTA
        return cls(psf.to(tdt), dtype=dtype), "reference"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lensless_oracle_torch as ot

    class _Wrap:
        def __init__(self):
            self.s = ot.ALGOS[algo](psf.to(tdt))

        def set_data(self, d):
            self.s.set_data(d)

        def apply(self, n_iter, **kw):
            return self.s.apply(n_iter)

    return _Wrap(), "port"


def torch_baseline_gpu(args, n_iter, device, psf, meas):
    """The reference's TORCH path on the SAME GPU ('the real bar for the fused
    kernels'), as profile/admm.py:93-108 runs it but WITH torch.cuda.synchronize(): full apply(n_iter) calls."""
    import torch

    solver, kind = reference_gpu_solver(args.algo, psf, "float32")
    solver.set_data(meas[0])
    solver.apply(n_iter=3, disp_iter=None, plot=False)
    torch.cuda.synchronize(device)
    k = 2
    t0 = time.perf_counter()
    for _ in range(k):
        solver.apply(n_iter=n_iter, disp_iter=None, plot=False)
    torch.cuda.synchronize(device)
    s = (time.perf_counter() - t0) / k
    del solver
    torch.cuda.empty_cache()
    return {
        "value": 1.0 / s, "unit": UNIT, "kind": kind + " (lensless torch path, same GPU)", "ms_per_recon": s * 1e3,
        "ms_per_iter": s * 1e3 / n_iter,
        "sample": f"{k} full apply(n_iter={n_iter}) calls after a 3-iteration warm-up, wall clock between torch.cuda.synchronize()",
    }


def parity_block(L, device, psf, meas, configs):
    """ours vs the reference's torch path on this GPU in fp64 (truth) and fp32 (trajectory), rel-to-max errors"""
    import torch

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max())

    out = {}
    for algo, n in configs:
        try:
            cls = L.ADMM if algo == "admm" else L.FISTA
            rec = cls(psf)
            rec.set_data(meas[:1])
            mine = rec.apply(n_iter=n, disp_iter=None, plot=False)
            del rec
            res = {}
            for dt in ("float64", "float32"):
                solver, kind = reference_gpu_solver(algo, psf, dt)
                solver.set_data(meas[0].to(torch.float64 if dt == "float64" else torch.float32))
                res[dt] = solver.apply(n_iter=n, disp_iter=None, plot=False).clone()
                del solver
                torch.cuda.empty_cache()
            out[f"{algo}-{n}"] = {
                "ours_vs_fp64": rel(mine, res["float64"]), "ref32_vs_fp64": rel(res["float32"], res["float64"]),
                "ours_vs_ref32": rel(mine, res["float32"]), "truth": kind + " torch path, fp64, same GPU", "metric": "max|a-b|/max|b|",
            }
            del mine, res
        except Exception as e:
            out[f"{algo}-{n}"] = {"unavailable": repr(e)[:200]}
        torch.cuda.empty_cache()
    return out


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

    import numpy as np
    import torch

    import lenslesspicam_b200 as L
    from lenslesspicam_b200 import _lib
    from lenslesspicam_b200 import dist as lpd

    world, rank, local = lpd.env_world()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    lpd.init(backend="nccl", device=device)
    lib = _lib.load()
    peak, peak_src = measured_peaks()

    n_iter = args.n_iter or (100 if args.algo == "admm" else 300)
    h, w, C, Bg = args.height, args.width, args.channels, args.batch_per_gpu
    D = args.depth
    headline = (args.algo == "admm" and D == 1 and n_iter == 100 and (h, w, C, Bg) == (1080, 1920, 3, 1) and args.engine is None)
    if args.configs == "default":
        want = {"fista300", "bundled", "depth16", "raw", "batch256"} if headline else set()
    elif args.configs == "none":
        want = set()
    else:
        want = {c.strip() for c in args.configs.split(",") if c.strip()}

    stream = torch.cuda.Stream(device=device)

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

    def make_problem(D_, h_, w_, C_, B_, seed_psf=0):
        """synthetic inputs: PSF on rank 0, broadcast once over NCCL; per-rank scenes"""
        gen = torch.Generator().manual_seed(seed_psf)
        psf_ = torch.rand((D_, h_, w_, C_), generator=gen) ** 8
        psf_ = (psf_ / psf_.norm()).to(device)
        if world > 1 and rank != 0:
            psf_.zero_()
        lpd.broadcast_psf(psf_, src=0)  # the only collective of the job: init-time, NCCL over NVLink/NVSwitch
        gen = torch.Generator().manual_seed(1000 + rank)
        scene = (torch.rand((B_, 1, h_, w_, C_), generator=gen) ** 2).to(device)
        if D_ == 1:
            sim = L.RealFFTConvolve2D(psf_, pad=True, norm="backward")
            meas_ = sim.convolve(scene).clamp_(min=0)
            meas_ = meas_ / meas_.amax(dim=(1, 2, 3, 4), keepdim=True)
            del sim
        else:
            meas_ = scene  # the arithmetic is data-independent except clamps / thresholds: any measurement times the same
        return psf_, meas_

    def bench_config(algo, psf_, meas_, n_it, steps, warm=3, **extra):
        """device-resident recon/s of one configuration, same timing rules as the headline"""
        cls = L.ADMM if algo == "admm" else L.FISTA
        rec = cls(psf_, n_iter=n_it, **extra)

        def step():
            rec.set_data(meas_)
            return rec.apply(disp_iter=None, plot=False)

        for _ in range(warm):
            step()
        ms = timed(step, steps)
        bpi = rec.bytes_per_iter()
        eng = rec.engine_version
        B_ = int(meas_.shape[0])
        del rec
        torch.cuda.empty_cache()
        ms_step = ms / steps
        gbs = bpi * n_it / (ms_step * 1e-3) / 1e9
        return {"value": world * B_ * 1e3 / ms_step, "unit": UNIT, "ms_per_step": ms_step, "steps": steps, "warmup": warm,
                "n_iter": n_it, "engine_version": eng,
                "roofline": {"bound": "hbm", "algorithmic_bytes_per_iter": bpi, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak}}

    with torch.cuda.stream(stream):
        psf, meas = make_problem(D, h, w, C, Bg)
        host_in = meas.cpu().pin_memory()
        host_out = torch.empty((Bg, D, h, w, C), dtype=torch.float32).pin_memory()

        cls = L.ADMM if args.algo == "admm" else L.FISTA
        extra = {"allow_depth": True} if (D > 1 and args.algo == "admm") else {}
        recon = cls(psf, n_iter=n_iter, no_graph=args.no_graph, engine=args.engine, **extra)

        def step_device():
            recon.set_data(meas)
            return recon.apply(disp_iter=None, plot=False)

        def step_host():
            recon.apply_host(host_in, host_out, n_iter=n_iter)
            stream.synchronize()  # the caller reads the image: a real end-to-end step ends on the host

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
        _lib.check(lib.lpc_profile(recon._plan, 10, stream.cuda_stream, phase_ms, ctypes.byref(nph)), recon._plan)
        phases = {}
        for i in range(nph.value):
            name = lib.lpc_phase_name(recon._plan, i).decode()
            phases[name] = phases.get(name, 0.0) + phase_ms[i]
        bytes_iter = recon.bytes_per_iter()
        # non-iteration share of a reconstruction: pack (set_data), reset memsets, crop/clamp (get_image), graph launch
        ov = {}
        for name, fn in (("set_data(pack)", lambda: recon._send_data()), ("reset", recon.reset), ("form_image(crop+clamp)", recon._form_image)):
            fn()
            ov[name] = timed(fn, 5) / 5
        engine = lib.lpc_engine_version(recon._plan)
        _, Hp, Wp, _ = recon._padded_shape

        # ---- the other BASELINE configurations, same run
        cfgs = {}
        if "fista300" in want:
            cfgs["fista300"] = bench_config("fista", psf, meas, 300, 3)
            cfgs["fista300"]["workload"] = workload("fista", 300, C, h, w)
        if "bundled" in want:
            for key, fname in (("bundled_ds8", "bundled_ds8.npz"), ("joss_ds4_gray", "bundled_ds4_gray.npz")):
                fp = os.path.join(ROOT, "tests", "golden", fname)
                if not os.path.exists(fp):
                    continue
                g = np.load(fp)
                bp, bd = torch.from_numpy(g["psf"]).to(device), torch.from_numpy(g["data"]).to(device)[None]
                shape = "x".join(str(v) for v in (bp.shape[3], bp.shape[1], bp.shape[2]))
                for algo_, n_ in (("admm", 5), ("fista", 300)):
                    r = bench_config(algo_, bp, bd, n_, 10)
                    r["workload"] = (f"{algo_.upper()} n_iter={n_} fp32 on the reference's bundled tape PSF / thumbs-up measurement, {shape} "
                                     f"({'load_data(downsample=8), BASELINE configs[0]' if key == 'bundled_ds8' else 'downsample=4 gray: profile/admm.py:12-16, the JOSS paper table'})")
                    cfgs[f"{key}_{algo_}{n_}"] = r
                del bp, bd
        if "depth16" in want:
            dp, dm = make_problem(16, h, w, 1, 1, seed_psf=7)
            cfgs["depth16_admm50"] = bench_config("admm", dp, dm, 50, 3, allow_depth=True)
            cfgs["depth16_admm50"]["workload"] = workload("admm", 50, 1, h, w, D=16)
            del dp, dm
        if "raw" in want:
            # from a raw 12-bit sensor frame in pinned host memory to the image, per step: H2D of the
            # uint16 frame (3040x4056x3, the bundled data's size), GPU loader (BGR->RGB, background subtraction, clip,
            # antialiased downsample x4 -> 760x1014, max-normalise), ADMM-100 on the fused engine, D2H of the image
            from lenslesspicam_b200 import io as lio

            Hr, Wr = 3040, 4056
            gen = torch.Generator().manual_seed(5)
            raw_psf = (torch.rand((Hr, Wr, 3), generator=gen) ** 12 * 3900 + 64).to(torch.int16)
            raw_meas = (torch.rand((Hr, Wr, 3), generator=gen) * 2000 + 64).to(torch.int16).pin_memory()
            hh, ww = Hr // 4, Wr // 4
            rpsf = torch.empty((1, hh, ww, 3), dtype=torch.float32, device=device)
            rbg = torch.zeros((3,), dtype=torch.float32, device=device)
            d_raw = raw_psf.to(device)
            _lib.check(lib.lpc_io_load_psf(local, 16, d_raw.data_ptr(), Hr, Wr, 3, lio.IO_BGR, 5, 25, hh, ww, rpsf.data_ptr(), rbg.data_ptr(), stream.cuda_stream))
            rrec = L.ADMM(rpsf, n_iter=100)
            rdata = torch.empty((1, 1, hh, ww, 3), dtype=torch.float32, device=device)
            rout = torch.empty((1, hh, ww, 3), dtype=torch.float32).pin_memory()
            d_meas = torch.empty((Hr, Wr, 3), dtype=torch.int16, device=device)

            def step_raw():
                d_meas.copy_(raw_meas, non_blocking=True)
                _lib.check(lib.lpc_io_load_image(local, 16, d_meas.data_ptr(), Hr, Wr, 3, lio.IO_BGR | lio.IO_NORMALIZE_FIRST, rbg.data_ptr(),
                                                 hh, ww, 1, rdata.data_ptr(), stream.cuda_stream))
                rrec.set_data(rdata)
                rout.copy_(rrec.apply(disp_iter=None, plot=False), non_blocking=True)
                stream.synchronize()

            for _ in range(3):
                step_raw()
            ms_raw = timed(step_raw, 5) / 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5):
                _lib.check(lib.lpc_io_load_image(local, 16, d_meas.data_ptr(), Hr, Wr, 3, lio.IO_BGR | lio.IO_NORMALIZE_FIRST, rbg.data_ptr(),
                                                 hh, ww, 1, rdata.data_ptr(), stream.cuda_stream))
            e1.record(stream)
            stream.synchronize()
            cfgs["raw_frame_to_image"] = {
                "value": world * 1e3 / ms_raw, "unit": UNIT, "ms_per_step": ms_raw, "loader_ms": e0.elapsed_time(e1) / 5,
                "h2d_bytes_per_step": int(Hr * Wr * 3 * 2), "d2h_bytes_per_step": int(hh * ww * 3 * 4), "engine_version": rrec.engine_version,
                "workload": f"raw uint16 {Hr}x{Wr}x3 frame (pinned host) -> GPU loader (load_data arithmetic, downsample 4) -> ADMM-100 @ 3x{hh}x{ww} -> image in host memory",
            }
            del rrec, d_raw, d_meas, rdata, rpsf
        del recon
        torch.cuda.empty_cache()
        if "batch256" in want:
            # BASELINE configs[3]: N host-resident measurements, batch axis sharded over the ranks, every measurement
            # copied H2D and its image D2H inside the timed region; time-to-solution = max over ranks
            NB = args.batch256_n
            base = meas[0:1].cpu()
            mine_idx = lpd.shard_batch(NB, world, rank)
            n_mine = len(mine_idx)
            host_batch = torch.empty((n_mine, 1, h, w, C), dtype=torch.float32).pin_memory()  # this rank's shard only
            for k, i in enumerate(mine_idx):  # distinct measurements (a cheap roll of a synthetic one)
                host_batch[k] = torch.roll(base[0], shifts=(7 * i) % h, dims=1)
            solver = L.ADMM(psf, n_iter=100)
            lpd.reconstruct_sharded(host_batch[:2], psf, n_iter=100, solver=solver, total=2 * world)  # warm-up: plans, graphs
            out_buf = torch.empty((n_mine, 1, h, w, C), dtype=torch.float32).pin_memory()  # the caller's result buffer
            barrier()
            t0 = time.perf_counter()
            local_out, mine = lpd.reconstruct_sharded(host_batch, psf, n_iter=100, solver=solver, out=out_buf, total=NB)
            torch.cuda.synchronize(device)
            dt_ms = lpd.max_over_ranks((time.perf_counter() - t0) * 1e3, device=device)
            bpi = 100 * solver.bytes_per_iter()  # per image (plans of batch 1)
            cfgs["batch256"] = {
                "value": NB * 1e3 / dt_ms, "unit": UNIT, "time_to_solution_ms": dt_ms, "measurements": NB, "per_gpu": len(mine),
                "scaling": "strong", "h2d_bytes": int(NB * 4 * h * w * C), "d2h_bytes": int(NB * 4 * h * w * C),
                "workload": f"ADMM n_iter=100, batch={NB} synthetic {C}x{h}x{w} host-resident measurements sharded across {world} GPU(s) "
                            "(BASELINE configs[3]) through lenslesspicam_b200.dist.reconstruct_sharded; host wall clock around the call, max over ranks",
                "roofline": {"bound": "hbm", "achieved": NB * bpi / (dt_ms * 1e-3) / 1e9 / world, "peak": peak, "unit": "GB/s per GPU",
                             "frac": NB * bpi / (dt_ms * 1e-3) / 1e9 / world / peak},
            }
            del solver, host_batch, local_out, out_buf
            torch.cuda.empty_cache()

    if rank != 0:
        lpd.finalize()
        return

    total = world * Bg
    ms_step = ms_dev / args.steps
    value = total * args.steps / (ms_dev / 1e3)
    e2e_value = total * args.steps / (ms_e2e / 1e3)

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
        "phases_frac": {n: kernel_bytes[n] / (t * 1e-3) / 1e9 / peak for n, t in own.items() if t > 0},
        "engine": "v2 fused own-FFT (3 kernels / iteration)" if engine == 2 else "v1 K1/K2 + cuFFT",
        "iteration": {
            "algorithmic_bytes": bytes_iter, "ms_ungraphed": iter_ms, "ms_in_step": ms_step / n_iter / Bg,
            "achieved_GBs": bytes_iter / (ms_step / n_iter * 1e-3) / 1e9,
            "frac_of_peak": bytes_iter / (ms_step / n_iter * 1e-3) / 1e9 / peak,
        },
        "per_recon_overhead_ms": {**ov, "iterations(graph)": ms_step - sum(ov.values()), "step": ms_step},
    }
    line = {
        "metric": metric_name(args.algo, n_iter, C, h, w, D),
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": workload(args.algo, n_iter, C, h, w, D),
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
        "configs": cfgs,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, n_iter)
        if D == 1:
            try:
                line["torch_baseline_gpu"] = torch_baseline_gpu(args, n_iter, device, psf, meas)
            except Exception as e:  # never lose the bench line to the extra baseline
                line["torch_baseline_gpu"] = {"unavailable": repr(e)[:200]}
            if headline:
                line["parity"] = parity_block(L, device, psf, meas, (("admm", 100), ("fista", 300)))
    out.emit(json.dumps(line))
    lpd.finalize()


if __name__ == "__main__":
    main()
