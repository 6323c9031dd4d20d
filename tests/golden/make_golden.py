"""
Generate the golden fixtures in this directory by EXECUTING THE REFERENCE
(``/root/reference``, LenslessPiCam v1.0.8) in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference cannot travel to the GPU box, so its outputs are committed here
together with this script.  ``matplotlib`` and ``rawpy`` are not installed in
the image; they are only imported at module top level by the reference
(recon.py:166-167, utils/io.py:93) and never called on this path, so empty stub
modules are created in a temp dir.

Fixtures (all float32 inputs, seeded):

* ``small_even_rgb``   24x32x3  -> padded 48x64     ADMM {1,5,30}, GD/Nesterov/FISTA {1,5,30}
* ``small_odd_rgb``    23x38x3  -> padded 45x75 (odd x odd)  same
* ``small_gray``       20x27x1  -> padded 40x54
* ``small_mixed``      14x32x3  -> padded 27x64 (odd x even)
* ``bundled_ds16``     the reference's bundled tape PSF + thumbs-up measurement
                       through ``load_data(downsample=16)`` -> 190x253x3, ADMM-5/FISTA-20 outputs
* ``convolver``        RealFFTConvolve2D pad/crop/convolve/deconvolve on the
                       reference test's shapes (PSF 5x47x29x3, data 2x5x47x29x3 (the reference test uses batch 12); test_convolver.py:11-59)
* ``batch``            UnrolledADMM.forward(batch) == per-sample ADMM
* ``unrolled``         UnrolledADMM / UnrolledFISTA .forward(batch) with per-iteration parameters
                       (``python tests/golden/make_golden.py --only-unrolled`` regenerates just this one)

Each npz stores inputs and, per algorithm/iteration count, the reference's
torch-fp32 output (``t32``: the parity target), its numpy-fp32 output (``n32``)
and torch-fp64 (``t64``: truth for error envelopes).
"""

import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _import_reference():
    stub = tempfile.mkdtemp(prefix="lpc_stubs_")
    for mod in ("matplotlib", "rawpy"):
        os.makedirs(os.path.join(stub, mod), exist_ok=True)
        open(os.path.join(stub, mod, "__init__.py"), "w").close()
    for sub in ("pyplot", "cm"):
        open(os.path.join(stub, "matplotlib", sub + ".py"), "w").close()
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    import lensless  # noqa

    return lensless


def synth(h, w, C, seed, D=1):
    rng = np.random.default_rng(seed)
    psf = rng.random((D, h, w, C)) ** 8
    psf /= np.linalg.norm(psf.ravel())
    data = rng.random((1, h, w, C))
    # a smooth-ish blob so TV / non-negativity are both active
    yy, xx = np.mgrid[0:h, 0:w]
    blob = np.exp(-(((yy - h / 2) / (h / 4)) ** 2 + ((xx - w / 3) / (w / 5)) ** 2))
    data = 0.3 * data + 0.7 * blob[None, :, :, None]
    data /= data.max()
    return psf.astype(np.float32), data.astype(np.float32)


def run_ref(lensless, torch, name, psf, data, n_iters, backend, **kw):
    cls = getattr(lensless, name)
    if backend == "n32":
        p, d, dt = psf.copy(), data.copy(), "float32"
    elif backend == "t32":
        p, d, dt = torch.from_numpy(psf.copy()), torch.from_numpy(data.copy()), "float32"
    else:
        p, d, dt = (
            torch.from_numpy(psf.astype(np.float64)),
            torch.from_numpy(data.astype(np.float64)),
            "float64",
        )
    out = {}
    for n in n_iters:
        rec = cls(p, dtype=dt, **kw)
        rec.set_data(d)
        res = rec.apply(n_iter=n, disp_iter=None, plot=False)
        res = res.numpy() if hasattr(res, "numpy") else np.asarray(res)
        out[n] = np.array(res, copy=True)
    return out


ALGOS = {"admm": "ADMM", "gd": "GradientDescent", "nesterov": "NesterovGradientDescent", "fista": "FISTA"}


def make_case(lensless, torch, fname, psf, data, iters):
    store = {"psf": psf, "data": data}
    for key, cls in ALGOS.items():
        for backend in ("t32", "n32", "t64"):
            if backend == "n32" and key == "fista":
                continue  # numpy FISTA upcasts to float64 under NumPy 2 (gd.py:238-239)
            res = run_ref(lensless, torch, cls, psf, data, iters[key], backend)
            for n, arr in res.items():
                store[f"{key}_{backend}_{n}"] = arr
    np.savez_compressed(os.path.join(HERE, fname), **store)
    print("wrote", fname, {k: v.shape for k, v in store.items() if k in ("psf", "data")})


def make_unrolled(lensless, torch):
    """``unrolled.npz``: UnrolledADMM / UnrolledFISTA .forward(batch) with DIFFERENT parameters per iteration
    (unrolled_admm.py:80-99,181-234; unrolled_fista.py:60-83,102-106), torch fp32 and fp64, no_grad."""
    psf, _ = synth(24, 32, 3, 9)
    batch = np.stack([synth(24, 32, 3, 200 + b)[1] for b in range(3)])  # [3,1,24,32,3]
    n = 6
    rng = np.random.default_rng(99)
    mu1 = 1e-6 * (0.5 + rng.random(n) * 2)
    mu2 = 1e-5 * (0.5 + rng.random(n) * 2)
    mu3 = 4e-5 * (0.5 + rng.random(n) * 2)
    tau = 1e-4 * (0.5 + rng.random(n) * 2)
    store = {"psf": psf, "batch": batch, "mu1": mu1, "mu2": mu2, "mu3": mu3, "tau": tau}
    with torch.no_grad():
        for tag, tdt, sdt in (("t32", torch.float32, "float32"), ("t64", torch.float64, "float64")):
            un = lensless.UnrolledADMM(torch.from_numpy(psf).to(tdt), dtype=sdt, n_iter=n)
            un._mu1_p = torch.nn.Parameter(torch.from_numpy(mu1).to(tdt))
            un._mu2_p = torch.nn.Parameter(torch.from_numpy(mu2).to(tdt))
            un._mu3_p = torch.nn.Parameter(torch.from_numpy(mu3).to(tdt))
            un._tau_p = torch.nn.Parameter(torch.from_numpy(tau).to(tdt))
            store[f"unrolled_admm_{tag}"] = un.forward(torch.from_numpy(batch).to(tdt)).numpy().copy()
        # FISTA: perturbed step sizes per iteration and channel, a non-default momentum schedule
        un = lensless.UnrolledFISTA(torch.from_numpy(psf), n_iter=n, tk=1.5)
        alpha0 = un._alpha_p.detach().numpy().astype(np.float64)  # [n, C] = 1.8 / max|H*H|
        alpha = alpha0 * (0.6 + 0.5 * rng.random(alpha0.shape))
        tk = un._tk_p.detach().numpy().astype(np.float64) * (0.9 + 0.2 * rng.random(n + 1))
        store["alpha"] = alpha
        store["tk"] = tk
        for tag, tdt, sdt in (("t32", torch.float32, "float32"), ("t64", torch.float64, "float64")):
            un = lensless.UnrolledFISTA(torch.from_numpy(psf).to(tdt), dtype=sdt, n_iter=n, tk=1.5)
            un._alpha_p = torch.nn.Parameter(torch.from_numpy(alpha).to(tdt))
            un._tk_p = torch.nn.Parameter(torch.from_numpy(tk).to(tdt))
            store[f"unrolled_fista_{tag}"] = un.forward(torch.from_numpy(batch).to(tdt)).numpy().copy()
        # forward(batch, psfs): one PSF per sample; UnrolledFISTA keeps the CONSTRUCTOR PSF's initial value and step
        # sizes (trainable_recon.py:347-350 swaps the convolver only; unrolled_fista.py:55-73)
        psfs = np.stack([psf * (0.5 + rng.random(psf.shape)).astype(np.float32) for _ in range(3)])
        store["psfs"] = psfs
        un = lensless.UnrolledFISTA(torch.from_numpy(psf), n_iter=n, tk=1.5)
        un._alpha_p = torch.nn.Parameter(torch.from_numpy(alpha).to(torch.float32))
        un._tk_p = torch.nn.Parameter(torch.from_numpy(tk).to(torch.float32))
        store["unrolled_fista_psfs_t32"] = un.forward(torch.from_numpy(batch), psfs=torch.from_numpy(psfs)).numpy().copy()
        un = lensless.UnrolledADMM(torch.from_numpy(psf), n_iter=n)
        un._mu1_p = torch.nn.Parameter(torch.from_numpy(mu1).to(torch.float32))
        un._mu2_p = torch.nn.Parameter(torch.from_numpy(mu2).to(torch.float32))
        un._mu3_p = torch.nn.Parameter(torch.from_numpy(mu3).to(torch.float32))
        un._tau_p = torch.nn.Parameter(torch.from_numpy(tau).to(torch.float32))
        store["unrolled_admm_psfs_t32"] = un.forward(torch.from_numpy(batch), psfs=torch.from_numpy(psfs)).numpy().copy()
        # defaults (constant parameters): must equal ADMM / a FISTA without the first-iteration aliasing
        un = lensless.UnrolledFISTA(torch.from_numpy(psf), n_iter=n)
        store["unrolled_fista_default_t32"] = un.forward(torch.from_numpy(batch)).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "unrolled.npz"), **store)
    print("wrote unrolled.npz", {k: v.shape for k, v in store.items()})


def make_bundled_large(lensless, torch):
    """BASELINE configs[0] at the sizes the reference's own scripts use:

    * ``bundled_ds8.npz``      ``load_data(downsample=8)`` RGB -> 3x380x507 (padded 768x1024): ADMM-5 / ADMM-100 /
                               FISTA-300 of the reference's torch-fp32 path, plus ``reconstruction_error`` values
    * ``bundled_ds4_gray.npz`` ``load_data(downsample=4, gray=True)`` -> 1x760x1014 (padded 1536x2048), the settings of
                               profile/admm.py:12-16 and profile/gradient_descent.py:12-16 (the JOSS paper's table)

    fp64 truths are NOT stored (size): the GPU tests recompute them with oracle/lensless_oracle_torch.py in fp64,
    which tests/test_oracle_golden.py pins against the fp64 goldens of the small cases."""
    from lensless.utils.io import load_data

    cwd = os.getcwd()
    os.chdir(REF)
    try:
        sets = {}
        for name, kw in (("bundled_ds8", dict(downsample=8)), ("bundled_ds4_gray", dict(downsample=4, gray=True))):
            psf, data = load_data(psf_fp="data/psf/tape_rgb.png", data_fp="data/raw_data/thumbs_up_rgb.png",
                                  plot=False, dtype="float32", normalize=True, **kw)
            sets[name] = (np.asarray(psf, dtype=np.float32), np.asarray(data, dtype=np.float32))
    finally:
        os.chdir(cwd)
    for name, (psf, data) in sets.items():
        store = {"psf": psf, "data": data}
        for n in (5, 100):
            store[f"admm_t32_{n}"] = run_ref(lensless, torch, "ADMM", psf, data, (n,), "t32")[n]
        store["fista_t32_300"] = run_ref(lensless, torch, "FISTA", psf, data, (300,), "t32")[300]
        # reconstruction_error (recon.py:607-653) of the reference on its own ADMM-5 state, both normalisations
        rec = lensless.ADMM(torch.from_numpy(psf.copy()))
        rec.set_data(torch.from_numpy(data.copy()))
        rec.apply(n_iter=5, disp_iter=None, plot=False)
        store["admm5_recon_error_norm"] = rec.reconstruction_error(normalize=True).numpy().copy()
        store["admm5_recon_error_raw"] = rec.reconstruction_error(normalize=False).numpy().copy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)
        print("wrote", name + ".npz", psf.shape, data.shape,
              {k: float(np.abs(v).max()) for k, v in store.items() if k.startswith(("admm", "fista"))})


def standin_denoiser(x, noise_level):
    """deterministic stand-in for DruNet in the plug-and-play goldens / tests: a 5-point periodic blur blended in by
    ``noise_level``, clamped at zero; shape-preserving on [..., H, W, C] tensors"""
    import torch

    blur = (x + torch.roll(x, 1, -3) + torch.roll(x, -1, -3) + torch.roll(x, 1, -2) + torch.roll(x, -1, -2)) / 5.0
    return torch.clamp((1.0 - noise_level) * x + noise_level * blur, min=0.0)


def make_pnp(lensless, torch):
    """``pnp.npz``: the reference's plug-and-play branches (admm.py:126-133,235-243,266-275,302-329; gd.py:89-92,136-140)
    with a stand-in denoiser.  The reference constructor only accepts the name "DruNet" (weights from the hub), so the
    objects are built without a denoiser and then put into exactly the state the constructor's PnP branch creates
    (admm.py:126-135: _denoiser, _denoiser_noise_level, _denoiser_use_dual, _proj = denoiser, _PsiT = identity, reset();
    gd.py:89-92: _proj = denoiser); every iteration after that is the reference's own code."""
    psf, data = synth(24, 32, 3, 31)
    nl = 0.3
    rng = np.random.default_rng(32)
    init = (rng.random((1, 1, 48, 64, 3)) * 0.1).astype(np.float32)  # padded-shape warm start: with use_dual the
    # reference's rk has no data term (admm.py:266-275), a zero start would stay zero
    store = {"psf": psf, "data": data, "noise_level": np.float64(nl), "init": init}
    for use_dual in (True, False):
        for tag, conv in (("t32", np.float32), ("t64", np.float64)):
            rec = lensless.ADMM(torch.from_numpy(psf.astype(conv)), dtype="float32" if conv is np.float32 else "float64",
                                initial_est=torch.from_numpy(init.astype(conv).copy()) if use_dual else None)
            rec._denoiser = standin_denoiser
            rec._denoiser_noise_level = nl
            rec._denoiser_use_dual = use_dual
            rec._proj = rec._denoiser
            rec._PsiT = lambda x: x
            rec.reset()
            rec.set_data(torch.from_numpy(data.astype(conv)))
            store[f"admm_dual{int(use_dual)}_{tag}_8"] = rec.apply(n_iter=8, disp_iter=None, plot=False).numpy().copy()
    for key, cls in (("gd", lensless.GradientDescent), ("nesterov", lensless.NesterovGradientDescent), ("fista", lensless.FISTA)):
        for tag, conv in (("t32", np.float32), ("t64", np.float64)):
            rec = cls(torch.from_numpy(psf.astype(conv)), dtype="float32" if conv is np.float32 else "float64")
            rec._denoiser = standin_denoiser
            rec._denoiser_noise_level = nl
            rec._proj = rec._denoiser
            rec.set_data(torch.from_numpy(data.astype(conv)))
            store[f"{key}_{tag}_8"] = rec.apply(n_iter=8, disp_iter=None, plot=False).numpy().copy()
    # a custom projection (gd.py:84-86): clamp into [0, 0.05]
    rec = lensless.FISTA(torch.from_numpy(psf), proj=lambda x: torch.clamp(x, 0.0, 0.05))
    rec.set_data(torch.from_numpy(data))
    store["fista_boxproj_t32_8"] = rec.apply(n_iter=8, disp_iter=None, plot=False).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "pnp.npz"), **store)
    print("wrote pnp.npz", {k: float(np.abs(v).max()) for k, v in store.items() if "_t32_" in k})


def make_loader(lensless, torch):
    """``loader.npz`` + small raw frames (``raw_psf.png``, ``raw_data.png``: synthetic 12-bit BGR uint16; ``raw8_*.png``:
    8-bit; ``real_*_crop.png``: a 456x608 crop of the reference's bundled tape PSF / thumbs-up measurement): outputs of the
    reference's own ``load_data`` / ``load_psf`` / ``load_image`` (io.py:20-600) on them, for the GPU loader's parity."""
    import cv2
    from lensless.utils.io import load_data, load_image, load_psf

    rng = np.random.default_rng(123)
    H, W = 304, 404
    yy, xx = np.mgrid[0:H, 0:W]
    psf = np.full((H, W, 3), 64.0) + rng.normal(0, 3, (H, W, 3))
    for _ in range(60):  # caustic-like bright spots
        cy, cx, a = rng.integers(30, H - 30), rng.integers(30, W - 30), rng.uniform(300, 3500)
        spot = a * np.exp(-(((yy - cy) / rng.uniform(1.5, 4)) ** 2 + ((xx - cx) / rng.uniform(1.5, 4)) ** 2))
        psf += spot[:, :, None] * rng.uniform(0.6, 1.0, 3)
    data = np.full((H, W, 3), 64.0) + rng.normal(0, 3, (H, W, 3))
    for _ in range(8):
        cy, cx, a = rng.integers(40, H - 40), rng.integers(40, W - 40), rng.uniform(500, 2500)
        blob = a * np.exp(-(((yy - cy) / rng.uniform(20, 60)) ** 2 + ((xx - cx) / rng.uniform(20, 60)) ** 2))
        data += blob[:, :, None] * rng.uniform(0.5, 1.0, 3)
    files = {}
    for name, arr in (("raw_psf", psf), ("raw_data", data)):
        fp = os.path.join(HERE, name + ".png")
        cv2.imwrite(fp, np.clip(arr, 0, 4095).astype(np.uint16))
        files[name] = fp
        fp8 = os.path.join(HERE, name.replace("raw_", "raw8_") + ".png")
        cv2.imwrite(fp8, np.clip(arr / 16, 0, 255).astype(np.uint8))
        files[name.replace("raw_", "raw8_")] = fp8
    for name, src in (("real_psf_crop", "data/psf/tape_rgb.png"), ("real_data_crop", "data/raw_data/thumbs_up_rgb.png")):
        img = cv2.imread(os.path.join(REF, src), cv2.IMREAD_UNCHANGED)
        fp = os.path.join(HERE, name + ".png")
        cv2.imwrite(fp, np.ascontiguousarray(img[1292:1748, 1724:2332]))
        files[name] = fp
    store = {}
    cases = {
        "ds4": dict(downsample=4, normalize=True),
        "ds4_nonorm": dict(downsample=4),
        "ds2_gray": dict(downsample=2, gray=True, normalize=True),
        "ds4_flip": dict(downsample=4, flip=True, normalize=True),
        "ds4_fliplr": dict(downsample=4, flip_lr=True, normalize=True),
        "ds3_single": dict(downsample=3, single_psf=True, normalize=True),
        "ds4_nobg": dict(downsample=4, bg_pix=None, normalize=True),
        "shape": dict(shape=(1, 100, 150, 3), downsample=2, normalize=True),  # (the reference needs a downsample next to shape: io.py:360)
        "ds1p6": dict(downsample=1.6, normalize=True),
    }
    for tag, kw in cases.items():
        if tag == "ds3_single":  # the reference then repeats the 1-channel PSF across the RGB data's channels
            psf_o, data_o = load_data(files["raw_psf"], files["raw_data"], plot=False, dtype="float32", **kw)
        else:
            psf_o, data_o = load_data(files["raw_psf"], files["raw_data"], plot=False, dtype="float32", **kw)
        store[f"{tag}_psf"], store[f"{tag}_data"] = np.asarray(psf_o, np.float32), np.asarray(data_o, np.float32)
    psf_o, data_o, bg_o = load_data(files["raw8_psf"], files["raw8_data"], plot=False, dtype="float32", downsample=4, normalize=True, return_bg=True)
    store["u8_ds4_psf"], store["u8_ds4_data"], store["u8_ds4_bg"] = np.asarray(psf_o, np.float32), np.asarray(data_o, np.float32), np.asarray(bg_o, np.float32)
    psf_o, data_o, bg_o = load_data(files["real_psf_crop"], files["real_data_crop"], plot=False, dtype="float32", downsample=4, normalize=True, return_bg=True)
    store["real_ds4_psf"], store["real_ds4_data"], store["real_ds4_bg"] = np.asarray(psf_o, np.float32), np.asarray(data_o, np.float32), np.asarray(bg_o, np.float32)
    p, b = load_psf(files["raw_psf"], downsample=4, return_bg=True)
    store["load_psf_ds4"], store["load_psf_ds4_bg"] = p, b
    store["load_image_ds4"] = load_image(files["raw_data"], downsample=4, return_float=True, as_4d=True, bg=b)
    store["load_image_full_nonorm"] = load_image(files["raw_data"], return_float=True, normalize=False)
    np.savez_compressed(os.path.join(HERE, "loader.npz"), **store)
    print("wrote loader.npz", {k: v.shape for k, v in store.items()})


def make_recon_error(lensless, torch):
    """``recon_error.npz``: reconstruction_error (recon.py:607-653) of ADMM (pad=False branch) and FISTA (pad=True
    branch) on a small case, normalised and raw, default arguments and explicit (prediction, lensless, psfs)."""
    psf, data = synth(24, 32, 3, 12)
    psf2, _ = synth(24, 32, 3, 13)
    store = {"psf": psf, "data": data, "psf2": psf2}
    for key, cls in (("admm", lensless.ADMM), ("fista", lensless.FISTA)):
        rec = cls(torch.from_numpy(psf.copy()))
        rec.set_data(torch.from_numpy(data.copy()))
        rec.apply(n_iter=10, disp_iter=None, plot=False)
        store[f"{key}_err_norm"] = rec.reconstruction_error(normalize=True).numpy().copy()
        store[f"{key}_err_raw"] = rec.reconstruction_error(normalize=False).numpy().copy()
        pred = rec.get_image_estimate().clone()
        store[f"{key}_pred"] = pred.numpy().copy()
        store[f"{key}_err_psf2"] = rec.reconstruction_error(prediction=pred, lensless=torch.from_numpy(data[None].copy()),
                                                            psfs=torch.from_numpy(psf2.copy())).numpy().copy()
    # set_image_estimate (recon.py:415-442) mid-run and right after a reset, gradient-descent family
    rng = np.random.default_rng(77)
    est = rng.random((1, 1, 24, 32, 3)).astype(np.float32) * 0.5
    store["est"] = est
    for key, cls in (("gd", lensless.GradientDescent), ("nesterov", lensless.NesterovGradientDescent), ("fista", lensless.FISTA)):
        rec = cls(torch.from_numpy(psf.copy()))
        rec.set_data(torch.from_numpy(data.copy()))
        rec.apply(n_iter=3, disp_iter=None, plot=False)
        rec.set_image_estimate(torch.from_numpy(est.copy()))
        store[f"{key}_setest_mid"] = rec.apply(n_iter=4, disp_iter=None, plot=False, reset=False).numpy().copy()
        rec.reset()
        rec.set_image_estimate(torch.from_numpy(est.copy()))
        store[f"{key}_setest_fresh"] = rec.apply(n_iter=4, disp_iter=None, plot=False, reset=False).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "recon_error.npz"), **store)
    print("wrote recon_error.npz", {k: v for k, v in store.items() if "err" in k})


def main():
    lensless = _import_reference()
    import torch

    if "--only-unrolled" in sys.argv:
        torch.manual_seed(0)
        return make_unrolled(lensless, torch)
    if "--only-bundled-large" in sys.argv:
        torch.set_num_threads(8)
        return make_bundled_large(lensless, torch)
    if "--only-recon-error" in sys.argv:
        return make_recon_error(lensless, torch)
    if "--only-pnp" in sys.argv:
        return make_pnp(lensless, torch)
    if "--only-loader" in sys.argv:
        return make_loader(lensless, torch)

    torch.manual_seed(0)
    torch.set_num_threads(8)
    from lensless.recon.rfft_convolve import RealFFTConvolve2D

    it = {"admm": (1, 5, 30), "gd": (1, 5, 30), "nesterov": (1, 5, 30), "fista": (1, 5, 30)}
    make_case(lensless, torch, "small_even_rgb.npz", *synth(24, 32, 3, 1), it)
    make_case(lensless, torch, "small_odd_rgb.npz", *synth(23, 38, 3, 2), it)
    make_case(lensless, torch, "small_gray.npz", *synth(20, 27, 1, 3), it)
    make_case(lensless, torch, "small_mixed.npz", *synth(14, 32, 3, 4), it)

    # ---- ADMM with non-default parameters and a warm start (padded-shape initial_est, SURVEY 8b viii)
    psf, data = synth(24, 32, 3, 5)
    rng = np.random.default_rng(55)
    init = rng.random((1, 48, 64, 3)).astype(np.float32) * 0.1
    store = {"psf": psf, "data": data, "init": init}
    kw = dict(mu1=3e-6, mu2=2e-5, mu3=1e-4, tau=3e-4)
    for backend in ("t32", "t64"):
        store[f"admm_params_{backend}_10"] = run_ref(lensless, torch, "ADMM", psf, data, (10,), backend, **kw)[10]
    for backend, conv in (("t32", np.float32), ("t64", np.float64)):
        rec = lensless.ADMM(torch.from_numpy(psf.astype(conv)), initial_est=torch.from_numpy(init.astype(conv).copy()))
        rec.set_data(torch.from_numpy(data.astype(conv)))
        store[f"admm_init_{backend}_10"] = rec.apply(n_iter=10, disp_iter=None, plot=False).numpy().copy()
    # continuation: apply(5) then apply(5, reset=False) == apply(10)
    rec = lensless.ADMM(torch.from_numpy(psf))
    rec.set_data(torch.from_numpy(data))
    rec.apply(n_iter=5, disp_iter=None, plot=False)
    store["admm_cont_t32_5p5"] = rec.apply(n_iter=5, disp_iter=None, plot=False, reset=False).numpy().copy()
    # background subtraction (recon.py:553-555)
    bg = (rng.random((1, 24, 32, 3)) * 0.2).astype(np.float32)
    rec = lensless.ADMM(torch.from_numpy(psf))
    rec.set_data(torch.from_numpy(data.copy()))
    store["bg"] = bg
    store["admm_bg_t32_5"] = rec.apply(n_iter=5, disp_iter=None, plot=False, background=torch.from_numpy(bg)).numpy().copy()
    # FISTA / GD with non-default tk / lip_fact
    rec = lensless.FISTA(torch.from_numpy(psf), tk=2.0, lip_fact=1.2)
    rec.set_data(torch.from_numpy(data))
    store["fista_params_t32_10"] = rec.apply(n_iter=10, disp_iter=None, plot=False).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "variants.npz"), **store)
    print("wrote variants.npz")

    # ---- 3-D (depth) gradient descent: D independent planes sharing alpha / x0 (SURVEY 8a note)
    psf3, _ = synth(16, 20, 3, 6, D=4)
    _, data3 = synth(16, 20, 3, 7)
    store = {"psf": psf3, "data": data3}
    for key in ("gd", "fista"):
        store[f"{key}_t32_10"] = run_ref(lensless, torch, ALGOS[key], psf3, data3, (10,), "t32")[10]
        store[f"{key}_t64_10"] = run_ref(lensless, torch, ALGOS[key], psf3, data3, (10,), "t64")[10]
    np.savez_compressed(os.path.join(HERE, "depth_gd.npz"), **store)
    print("wrote depth_gd.npz")

    # ---- convolver on the reference test's shapes (test_convolver.py)
    rng = np.random.default_rng(7)
    psf = rng.random((5, 47, 29, 3)).astype(np.float32)
    data = rng.random((2, 5, 47, 29, 3)).astype(np.float32)
    data_d1 = rng.random((2, 1, 47, 29, 3)).astype(np.float32)
    store = {"psf": psf, "data": data, "data_d1": data_d1}
    for norm in ("ortho", "backward"):
        conv = RealFFTConvolve2D(torch.from_numpy(psf), pad=True, norm=norm)
        store[f"padded_shape"] = np.array(conv._padded_shape)
        store[f"start"] = np.array(conv._start_idx)
        store[f"conv_{norm}"] = conv.convolve(torch.from_numpy(data)).numpy()
        store[f"deconv_{norm}"] = conv.deconvolve(torch.from_numpy(data)).numpy()
        store[f"conv_d1_{norm}"] = conv.convolve(torch.from_numpy(data_d1)).numpy()
        if norm == "ortho":
            store["H_ortho"] = conv._H.numpy()
    conv = RealFFTConvolve2D(torch.from_numpy(psf), pad=False, norm="backward")
    padded = conv._pad(torch.from_numpy(data))
    store["conv_nopad"] = conv.convolve(padded).numpy()
    store["deconv_nopad"] = conv.deconvolve(padded).numpy()
    np.savez_compressed(os.path.join(HERE, "convolver.npz"), **store)
    print("wrote convolver.npz")

    # ---- batch: UnrolledADMM.forward(batch) (the only batched implementation)
    psf, _ = synth(24, 32, 3, 8)
    batch = np.stack([synth(24, 32, 3, 100 + b)[1] for b in range(4)])  # [4,1,24,32,3]
    store = {"psf": psf, "batch": batch}
    with torch.no_grad():
        un = lensless.UnrolledADMM(torch.from_numpy(psf), n_iter=10)
        store["unrolled_admm_t32_10"] = un.forward(torch.from_numpy(batch)).numpy().copy()
    per = []
    for b in range(4):
        rec = lensless.ADMM(torch.from_numpy(psf))
        rec.set_data(torch.from_numpy(batch[b]))
        per.append(rec.apply(n_iter=10, disp_iter=None, plot=False).numpy().copy())
    store["admm_t32_10"] = np.stack(per)
    np.savez_compressed(os.path.join(HERE, "batch.npz"), **store)
    print("wrote batch.npz")

    make_unrolled(lensless, torch)

    # ---- bundled data through the reference's own loader (config 1 inputs)
    from lensless.utils.io import load_data

    cwd = os.getcwd()
    os.chdir(REF)
    try:
        psf, data = load_data(
            psf_fp="data/psf/tape_rgb.png",
            data_fp="data/raw_data/thumbs_up_rgb.png",
            downsample=16,
            plot=False,
            dtype="float32",
            normalize=True,
        )
    finally:
        os.chdir(cwd)
    psf = np.asarray(psf, dtype=np.float32)
    data = np.asarray(data, dtype=np.float32)
    store = {"psf": psf, "data": data}
    for backend in ("t32", "n32", "t64"):
        store[f"admm_{backend}_5"] = run_ref(lensless, torch, "ADMM", psf, data, (5,), backend)[5]
    for backend in ("t32", "t64"):
        store[f"admm_{backend}_100"] = run_ref(lensless, torch, "ADMM", psf, data, (100,), backend)[100]
        store[f"fista_{backend}_20"] = run_ref(lensless, torch, "FISTA", psf, data, (20,), backend)[20]
    np.savez_compressed(os.path.join(HERE, "bundled_ds16.npz"), **store)
    print("wrote bundled_ds16.npz", psf.shape, data.shape)
    make_bundled_large(lensless, torch)
    make_recon_error(lensless, torch)
    make_pnp(lensless, torch)
    make_loader(lensless, torch)
    make_loader(lensless, torch)


if __name__ == "__main__":
    main()
