"""
CPU oracle for the LenslessPiCam iterative-deconvolution hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the baseline legs of ``bench.py``
(``cpu_baseline``, ``--impl reference``, and ``torch_baseline_gpu``, which times
the sibling ``lensless_oracle_torch.py``) may import it.  The product path
(``lenslesspicam_b200``) never imports, links or executes anything here.

It restates, in plain numpy + ``scipy.fft`` (pocketfft -- the same FFT the
reference's numpy path calls), the algorithm of these reference files
(``/root/reference``, LCAV/LenslessPiCam v1.0.8, commit dcb4aba):

* ``lensless/recon/rfft_convolve.py``  -- RealFFTConvolve2D (pad / crop / H / convolve / deconvolve)
* ``lensless/recon/admm.py``           -- ADMM (+ finite_diff, finite_diff_adj, finite_diff_gram, soft_thresh)
* ``lensless/recon/gd.py``             -- GradientDescent / NesterovGradientDescent / FISTA
* ``lensless/recon/recon.py``          -- set_data promotion, the apply() loop, background subtraction

The operation ORDER of the reference is kept literally (six FFTs per ADMM
iteration, re-padding the measurement every iteration, separate dual
updates) so that floating-point rounding follows the reference, not the
fused CUDA formulation it is used to check.

Parity pinning: the reference's own tests hold no golden values for this path
(SURVEY.md section 4), so this oracle is pinned against outputs of the
reference ITSELF, executed in the build container by
``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``
(``tests/test_oracle_golden.py`` checks oracle == reference on those).

Third-party arithmetic: the FFT is ``scipy.fft`` (pocketfft; reference pins
``scipy>=1.7``; this image has 1.18.1), ``next_fast_len`` is
``scipy.fftpack.next_fast_len`` (5-smooth sizes) -- restated here as
:func:`next_fast_len`.

Array layout is the reference's: data ``[B, D, H, W, C]``, PSF ``[D, H, W, C]``;
FFTs run over axes (-3, -2).
"""

import math

import numpy as np

try:  # pocketfft with a `workers` knob (the reference leaves it at 1 thread)
    from scipy import fft as _fft

    _HAVE_SCIPY = True
except Exception:  # pragma: no cover - scipy is in the image
    _fft = np.fft
    _HAVE_SCIPY = False


# ----------------------------------------------------------------------------
# geometry (integer work: must be bit-exact)
# ----------------------------------------------------------------------------


def next_fast_len(n):
    """Smallest 5-smooth integer (2^a 3^b 5^c) >= n.

    Restates ``scipy.fftpack.next_fast_len`` as used at rfft_convolve.py:16,112.
    """
    n = int(n)
    if n <= 6:
        return max(n, 0) if n > 0 else 0
    best = None
    p5 = 1
    while p5 < 2 * n:
        p35 = p5
        while p35 < 2 * n:
            # smallest power of two bringing p35 up to >= n
            q = -(-n // p35)  # ceil
            p2 = 1 << max(0, (q - 1).bit_length())
            cand = p35 * p2
            if cand >= n and (best is None or cand < best):
                best = cand
            p35 *= 3
        p5 *= 5
    return best


def padded_geometry(h, w):
    """(Hp, Wp, start_y, start_x) for a PSF of spatial size h x w.

    rfft_convolve.py:111-117: padded = next_fast_len(2*dim - 1);
    start = (padded - dim) // 2; end = start + dim.
    """
    Hp = next_fast_len(2 * h - 1)
    Wp = next_fast_len(2 * w - 1)
    return Hp, Wp, (Hp - h) // 2, (Wp - w) // 2


def _np_dtype(dtype):
    """recon.py:264-290 + utils/io.py:645-674: None -> float32; strings only."""
    if dtype is None:
        dtype = "float32"
    if isinstance(dtype, str):
        assert dtype in ("float32", "float64")
        return np.float32 if dtype == "float32" else np.float64
    return np.dtype(dtype).type


# ----------------------------------------------------------------------------
# RealFFTConvolve2D  (rfft_convolve.py:26-223)
# ----------------------------------------------------------------------------


class OracleConvolver:
    """Restatement of ``RealFFTConvolve2D`` (numpy branch)."""

    def __init__(self, psf, dtype=None, pad=True, norm="ortho", workers=None):
        assert psf.ndim >= 4  # rfft_convolve.py:55-57
        assert psf.shape[-1] in (1, 3)  # rfft_convolve.py:59-63
        self.dtype = _np_dtype(dtype)
        self.pad = pad
        self.norm = norm
        self.workers = workers
        self.set_psf(psf)

    # -- rfft_convolve.py:102-131
    def set_psf(self, psf):
        self.psf = np.asarray(psf).astype(self.dtype)
        D, h, w, C = self.psf.shape[-4:]
        Hp, Wp, sy, sx = padded_geometry(h, w)
        self.psf_hw = (h, w)
        self.padded_shape = [D, Hp, Wp, C]
        self.start = (sy, sx)
        self.end = (sy + h, sx + w)
        self.H = self._rfft2(self.pad_to(self.psf), norm=self.norm)
        self.Hadj = np.conj(self.H)

    def _kw(self):
        return {"workers": self.workers} if (_HAVE_SCIPY and self.workers) else {}

    def _rfft2(self, x, norm=None):
        return _fft.rfft2(x, axes=(-3, -2), norm=norm, **self._kw())

    def _irfft2(self, X):
        s = tuple(self.padded_shape[-3:-1])
        return _fft.irfft2(X, s=s, axes=(-3, -2), **self._kw())

    # -- rfft_convolve.py:84-100
    def pad_to(self, v):
        if v.ndim == 5:
            shape = [v.shape[0]] + self.padded_shape
        elif v.ndim == 4:
            shape = list(self.padded_shape)
        else:
            raise ValueError("Expected 4D or 5D tensor")
        out = np.zeros(shape, dtype=v.dtype)
        out[..., self.start[0] : self.end[0], self.start[1] : self.end[1], :] = v
        return out

    # -- rfft_convolve.py:79-82
    def crop(self, x):
        return x[..., self.start[0] : self.end[0], self.start[1] : self.end[1], :]

    def _apply(self, x, filt):
        xin = self.pad_to(x) if self.pad else x
        spec = self._rfft2(xin) * filt
        # rfft_convolve.py:160-170 / 205-216: irfft2 (backward norm) then ifftshift
        out = _fft.ifftshift(self._irfft2(spec), axes=(-3, -2))
        if self.pad:
            out = self.crop(out)
        return out

    # -- rfft_convolve.py:133-176
    def convolve(self, x):
        return self._apply(x, self.H)

    # -- rfft_convolve.py:178-223  (same shift as convolve: not an exact adjoint
    #    for odd padded sizes -- mirrored deliberately)
    def deconvolve(self, y):
        return self._apply(y, self.Hadj)


# ----------------------------------------------------------------------------
# ADMM helpers (admm.py:341-397)
# ----------------------------------------------------------------------------


def soft_thresh(x, thresh):
    """admm.py:341-346"""
    return np.sign(x) * np.maximum(0, np.abs(x) - thresh)


def finite_diff(x):
    """admm.py:349-359 -- periodic backward differences stacked on a new last axis."""
    dy = np.roll(x, 1, axis=-3) - x
    dx = np.roll(x, 1, axis=-2) - x
    return np.stack((dy, dx), axis=x.ndim)


def finite_diff_adj(u):
    """admm.py:362-370"""
    u0 = u[..., 0]
    u1 = u[..., 1]
    return (np.roll(u0, -1, axis=-3) - u0) + (np.roll(u1, -1, axis=-2) - u1)


def finite_diff_gram(shape, dtype, workers=None):
    """admm.py:373-397 -- rfft2 of the 5-point stencil; the depth > 1 branch (:388-392) is restated too although
    ADMM itself refuses depth > 1 (admm.py:91-96)."""
    g = np.zeros(shape, dtype=dtype)
    if shape[0] == 1:
        g[0, 0, 0] = 4
        g[0, 0, 1] = g[0, 0, -1] = g[0, 1, 0] = g[0, -1, 0] = -1
    else:
        g[0, 0, 0] = 6
        g[0, 0, 1] = g[0, 0, -1] = g[0, 1, 0] = g[0, -1, 0] = g[1, 0, 0] = g[-1, 0, 0] = -1
    kw = {"workers": workers} if (_HAVE_SCIPY and workers) else {}
    return _fft.rfft2(g, axes=(-3, -2), **kw)


def _promote_data(data):
    """recon.py:376-381"""
    data = np.asarray(data)
    assert data.ndim >= 3
    if data.ndim == 3:
        return data[None, None]
    if data.ndim == 4:
        return data[None]
    return data


def _apply_background(data, background):
    """recon.py:553-555"""
    if background is None:
        return data
    data = data - background
    data[data < 0] = 0
    return data


# ----------------------------------------------------------------------------
# ADMM (admm.py:24-338)
# ----------------------------------------------------------------------------


class OracleADMM:
    """Literal-order restatement of ``lensless.ADMM`` (numpy branch, TV prior, no denoiser)."""

    def __init__(
        self,
        psf,
        dtype=None,
        mu1=1e-6,
        mu2=1e-5,
        mu3=4e-5,
        tau=0.0001,
        n_iter=100,
        initial_est=None,
        workers=None,
    ):
        assert psf.ndim == 4
        if psf.shape[0] > 1:  # admm.py:93-96
            raise NotImplementedError("3D ADMM is not supported by the reference")
        self.mu1, self.mu2, self.mu3, self.tau = mu1, mu2, mu3, tau
        self.n_iter = n_iter
        self.dtype = _np_dtype(dtype)
        self.cdtype = np.complex64 if self.dtype == np.float32 else np.complex128
        self.workers = workers
        # admm.py:100-102: pad=False, norm="backward"
        self.conv = OracleConvolver(psf, dtype=self.dtype, pad=False, norm="backward", workers=workers)
        self.psf = self.conv.psf
        self.padded_shape = self.conv.padded_shape
        self.PsiTPsi = finite_diff_gram(self.padded_shape, self.dtype, workers)  # admm.py:107
        self.initial_est = None
        if initial_est is not None:
            ie = np.asarray(initial_est)
            self.initial_est = ie[None] if ie.ndim == 4 else ie
        self.data = None
        self.reset()

    def set_data(self, data):
        self.data = _promote_data(data)
        assert tuple(self.data.shape[-3:-1]) == tuple(self.psf.shape[-3:-1])

    # -- admm.py:196-230
    def reset(self):
        c = self.conv
        if self.initial_est is not None:
            self.V = self.initial_est
        else:
            self.V = np.zeros([1] + self.padded_shape, dtype=self.dtype)
        self.X = np.zeros_like(self.V)
        self.U = np.zeros_like(finite_diff(self.V))
        self.W = np.zeros_like(self.X)
        if self.V.max():
            self.HV = c.convolve(self.V)
            self.PsiV = finite_diff(self.V)
        else:
            self.HV = np.zeros_like(self.X)
            self.PsiV = np.zeros_like(self.U)
        self.xi = np.zeros_like(self.V)
        self.eta = np.zeros_like(self.U)
        self.rho = np.zeros_like(self.X)
        self.R = 1.0 / (
            self.mu1 * np.abs(c.Hadj * c.H) + self.mu2 * np.abs(self.PsiTPsi) + self.mu3
        ).astype(self.cdtype)
        self.Xdiv = 1.0 / (c.pad_to(np.ones(self.psf.shape, dtype=self.dtype)) + self.mu1)

    # -- admm.py:313-329 and the helpers it calls (232-311)
    def update(self):
        c = self.conv
        self.U = soft_thresh(self.PsiV + self.eta / self.mu2, self.tau / self.mu2)
        self.X = self.Xdiv * (self.xi + self.mu1 * self.HV + c.pad_to(self.data))
        self.W = np.maximum(self.rho / self.mu3 + self.V, 0)
        rk = (
            (self.mu3 * self.W - self.rho)
            + finite_diff_adj(self.mu2 * self.U - self.eta)
            + c.deconvolve(self.mu1 * self.X - self.xi)
        )
        self.V = c._irfft2(self.R * c._rfft2(rk))
        self.HV = c.convolve(self.V)
        self.PsiV = finite_diff(self.V)
        self.xi += self.mu1 * (self.HV - self.X)
        self.eta += self.mu2 * (self.PsiV - self.U)
        self.rho += self.mu3 * (self.V - self.W)

    # -- admm.py:331-338 (in-place clamp through the crop view, as the reference)
    def form_image(self):
        img = self.conv.crop(self.V)
        img[img < 0] = 0
        return img

    # -- recon.py:498-605 without plotting
    def apply(self, n_iter=None, reset=True, background=None):
        assert self.data is not None and self.data.shape[0] == 1
        self.data = _apply_background(self.data, background)
        if reset:
            self.reset()
        for _ in range(self.n_iter if n_iter is None else n_iter):
            self.update()
        return self.form_image()[0]


# ----------------------------------------------------------------------------
# Gradient descent family (gd.py:62-241)
# ----------------------------------------------------------------------------


class OracleGD:
    """``lensless.GradientDescent`` (numpy branch, proj = non_neg)."""

    def __init__(self, psf, dtype=None, lip_fact=1.8, n_iter=100, initial_est=None, workers=None):
        assert psf.ndim == 4
        self.dtype = _np_dtype(dtype)
        self.lip_fact = lip_fact
        self.n_iter = n_iter
        # recon.py:207,292-293 -> pad=True; RealFFTConvolve2D default norm="ortho"
        self.conv = OracleConvolver(psf, dtype=self.dtype, pad=True, norm="ortho", workers=workers)
        self.psf = self.conv.psf
        self.initial_est = None
        if initial_est is not None:
            ie = np.asarray(initial_est)
            self.initial_est = ie[None] if ie.ndim == 4 else ie
        self.data = None
        self.reset()

    def set_data(self, data):
        self.data = _promote_data(data)
        assert tuple(self.data.shape[-3:-1]) == tuple(self.psf.shape[-3:-1])

    # -- gd.py:114-126
    def reset(self):
        C = self.psf.shape[3]
        if self.initial_est is not None:
            self.x = np.array(self.initial_est, dtype=self.dtype)  # copy: the caller's array is not mutated
        else:
            flat = self.psf.reshape(-1, C)
            start = (flat.max(axis=0) + flat.min(axis=0)) / 2
            self.x = np.ones_like(self.psf[None]) * start
        HH = (self.conv.Hadj * self.conv.H).reshape(-1, C)
        # numpy branch takes max of a complex array whose imaginary part is ~0;
        # the torch branch uses max(abs(.)) (gd.py:110-112): identical in value.
        self.alpha = (self.lip_fact / np.max(np.abs(HH), axis=0)).astype(self.dtype)

    # -- gd.py:128-130
    def grad(self):
        return self.conv.deconvolve(self.conv.convolve(self.x) - self.data)

    @staticmethod
    def proj(x):  # gd.py:41-59
        return np.maximum(x, 0)

    # -- gd.py:132-134
    def update(self):
        self.x -= self.alpha * self.grad()  # in place, as gd.py:133
        self.x = self.proj(self.x)

    def form_image(self):  # gd.py:136-140
        return self.proj(self.x)

    def apply(self, n_iter=None, reset=True, background=None):
        assert self.data is not None and self.data.shape[0] == 1
        self.data = _apply_background(self.data, background)
        if reset:
            self.reset()
        for _ in range(self.n_iter if n_iter is None else n_iter):
            self.update()
        return self.form_image()[0]


class OracleNesterov(OracleGD):
    """gd.py:143-188.  ``reset()`` always restores p=0, mu=0.9 (SURVEY 8b quirk v)."""

    def reset(self):
        self.p = 0
        self.mu = 0.9
        super().reset()

    def update(self):
        p_prev = self.p
        self.p = self.mu * self.p - self.alpha * self.grad()
        self.x = self.x + (-self.mu * p_prev + (1 + self.mu) * self.p)
        self.x = self.proj(self.x)


class OracleFISTA(OracleGD):
    """gd.py:191-241.  The momentum scalar is kept a Python float so the state
    stays in the working dtype (the torch branch's behaviour; the numpy branch
    silently upcasts to float64 under NumPy >= 2)."""

    def __init__(self, psf, dtype=None, tk=1.0, **kw):
        self.initial_tk = tk
        super().__init__(psf, dtype, **kw)

    def reset(self):
        super().reset()
        self.tk = self.initial_tk
        self.xk = self.x

    def update(self):
        # gd.py:236 steps IN PLACE; right after reset() ``xk`` aliases ``x``
        # (gd.py:233), so in the first iteration the "previous" iterate is the
        # un-projected gradient step itself (SURVEY 8b quirk iii).  Invisible
        # with the default tk=1 (the momentum weight is then 0), visible for tk != 1.
        self.x -= self.alpha * self.grad()
        xk = self.proj(self.x)
        tk = (1 + math.sqrt(1 + 4 * self.tk**2)) / 2
        self.x = xk + self.dtype((self.tk - 1) / tk) * (xk - self.xk)
        self.tk = tk
        self.xk = xk


# ----------------------------------------------------------------------------
# Unrolled solvers, inference path (unrolled_admm.py:20-251, unrolled_fista.py:19-106)
# ----------------------------------------------------------------------------


class OracleUnrolledADMM(OracleADMM):
    """``lensless.UnrolledADMM.forward`` for one sample: ADMM whose mu1/mu2/mu3/tau have one entry per
    iteration (|.| of the parameter vectors, unrolled_admm.py:141-144), with R_divmat / X_divmat of
    that iteration (unrolled_admm.py:166-179) and the update order of unrolled_admm.py:222-234."""

    def __init__(self, psf, dtype=None, n_iter=5, mu1=1e-6, mu2=1e-5, mu3=4e-5, tau=0.0001, workers=None):
        def vec(v):
            a = np.abs(np.asarray(v, dtype=np.float64))
            return np.full((n_iter,), float(a)) if a.ndim == 0 else a

        self.s_mu1, self.s_mu2, self.s_mu3, self.s_tau = vec(mu1), vec(mu2), vec(mu3), vec(tau)
        super().__init__(psf, dtype=dtype, n_iter=n_iter, workers=workers)

    def reset(self):
        self.it = 0
        self._set_iter(0)
        super().reset()

    def _set_iter(self, i):
        dt = self.dtype
        self.mu1, self.mu2, self.mu3, self.tau = (dt(v[i]) for v in (self.s_mu1, self.s_mu2, self.s_mu3, self.s_tau))

    def update(self):
        c = self.conv
        self._set_iter(self.it)
        # unrolled_admm.py:166-179: per-iteration division matrices
        self.R = 1.0 / (self.mu1 * np.abs(c.Hadj * c.H) + self.mu2 * np.abs(self.PsiTPsi) + self.mu3).astype(self.cdtype)
        self.Xdiv = 1.0 / (c.pad_to(np.ones(self.psf.shape, dtype=self.dtype)) + self.mu1)
        super().update()
        self.it += 1


class OracleUnrolledFISTA(OracleGD):
    """``lensless.UnrolledFISTA.forward`` for one sample (unrolled_fista.py:84-106): out-of-place step with
    alpha[i][c], xk = proj(x), x = xk + (tk[i]-1)/tk[i+1] (xk - xk_prev); xk_prev starts as the initial estimate."""

    def __init__(self, psf, dtype=None, n_iter=5, alpha=None, tk=1.0, workers=None):
        self.s_alpha = None if alpha is None else np.abs(np.asarray(alpha, dtype=np.float64))
        tks = np.asarray(tk, dtype=np.float64)
        if tks.ndim == 0:
            t = [float(tks)]
            for i in range(n_iter):
                t.append((1 + math.sqrt(1 + 4 * t[i] ** 2)) / 2)
            tks = np.asarray(t)
        self.s_tk = np.abs(tks)
        super().__init__(psf, dtype=dtype, lip_fact=1.8, n_iter=n_iter, workers=workers)

    def reset(self):
        super().reset()
        if self.s_alpha is None:  # unrolled_fista.py:60-73: every iteration starts from 1.8 / max|H*H|
            self.s_alpha = np.tile(self.alpha.astype(np.float64)[None, :], (self.n_iter, 1))
        self.it = 0
        self.xk = self.x

    def update(self):
        i = self.it
        a = self.s_alpha[i].astype(self.dtype)
        self.x = self.x - a * self.grad()
        xk = self.proj(self.x)
        self.x = xk + self.dtype((self.s_tk[i] - 1) / self.s_tk[i + 1]) * (xk - self.xk)
        self.xk = xk
        self.it += 1


ALGOS = {
    "unrolled_admm": OracleUnrolledADMM,
    "unrolled_fista": OracleUnrolledFISTA,
    "admm": OracleADMM,
    "gd": OracleGD,
    "nesterov": OracleNesterov,
    "fista": OracleFISTA,
}


def reconstruct(algo, psf, data, n_iter, dtype="float32", workers=None, **params):
    """One-call helper used by tests / smoke / the CPU-baseline leg of bench.py."""
    solver = ALGOS[algo](psf, dtype=dtype, n_iter=n_iter, workers=workers, **params)
    solver.set_data(np.asarray(data).astype(solver.dtype))
    return solver.apply()


# ----------------------------------------------------------------------------
# seeded synthetic inputs shared by tests and bench
# ----------------------------------------------------------------------------


def synthetic_problem(h, w, C=3, D=1, seed=0, dtype=np.float32):
    """PSF = rand**8, L2-normalised (mirrors utils/io.py:375); measurement =
    pad-convolution of a rand**2 scene, max-normalised (mirrors normalize=True)."""
    rng = np.random.default_rng(seed)
    psf = rng.random((D, h, w, C)) ** 8
    psf /= np.linalg.norm(psf.ravel())
    scene = rng.random((1, D, h, w, C)) ** 2
    conv = OracleConvolver(psf, dtype=np.float64, pad=True, norm="backward")
    meas = conv.convolve(scene).sum(axis=1, keepdims=True)
    meas = np.maximum(meas, 0)
    meas /= meas.max()
    return psf.astype(dtype), meas[0].astype(dtype)
