"""
Torch restatement of the reference's torch branch for ADMM and FISTA -- TEST / MEASUREMENT
INFRASTRUCTURE, NOT PRODUCT CODE (same rule as lensless_oracle.py: only tests/, smoke() and bench.py's
baseline legs may import it; the product never does).

Why it exists: the reference cannot travel to the GPU box, and BASELINE.md (section 3, item 5) asks for the
reference's *torch* path timed on the same B200 with explicit synchronisation -- "the real bar for the fused
kernels".  This file restates that path operation for operation (the reference's op order, six FFTs per ADMM
iteration, the measurement re-padded every iteration, out-of-place torch ops), so its cost profile on the
device is the reference's: ~360 kernel launches and ~168 P of HBM traffic per ADMM iteration.

Reference lines followed:
  rfft_convolve.py:84-100 (_pad), :79-82 (_crop), :111-124 (H), :146-157 (convolve, torch), :191-202 (deconvolve)
  admm.py:150-194 (reset), :232-311 (updates), :313-329 (_update), :331-338 (_form_image), :341-370 (helpers),
  :373-397 (finite_diff_gram);  gd.py:94-126 (reset), :128-130 (_grad), :227-241 (FISTA).

Pinned by tests/test_oracle_golden.py against the reference's own torch-fp32 outputs (tests/golden/*.npz).
"""
import math

import torch


def _next_fast_len(n):
    import lensless_oracle as orc

    return orc.next_fast_len(n)


class TorchConvolver:
    def __init__(self, psf, pad=True, norm="ortho"):
        self.psf = psf
        self.pad = pad
        D, h, w, C = psf.shape
        Hp, Wp = _next_fast_len(2 * h - 1), _next_fast_len(2 * w - 1)
        self.padded_shape = [D, Hp, Wp, C]
        self.start = ((Hp - h) // 2, (Wp - w) // 2)
        self.end = (self.start[0] + h, self.start[1] + w)
        self.H = torch.fft.rfft2(self._pad(psf), norm=norm, dim=(-3, -2), s=(Hp, Wp))
        self.Hadj = torch.conj(self.H)

    def _pad(self, v):
        shape = ([v.shape[0]] if v.dim() == 5 else []) + self.padded_shape
        out = torch.zeros(shape, dtype=v.dtype, device=v.device)
        out[..., self.start[0] : self.end[0], self.start[1] : self.end[1], :] = v
        return out

    def _crop(self, x):
        return x[..., self.start[0] : self.end[0], self.start[1] : self.end[1], :]

    def _apply(self, x, filt):
        xin = self._pad(x) if self.pad else x
        s = tuple(self.padded_shape[-3:-1])
        out = torch.fft.ifftshift(
            torch.fft.irfft2(torch.fft.rfft2(xin, dim=(-3, -2)) * filt, dim=(-3, -2), s=s), dim=(-3, -2)
        )
        return self._crop(out) if self.pad else out

    def convolve(self, x):
        return self._apply(x, self.H)

    def deconvolve(self, y):
        return self._apply(y, self.Hadj)


def soft_thresh(x, thresh):
    return torch.sign(x) * torch.max(torch.abs(x) - thresh, torch.zeros_like(x))


def finite_diff(x):
    return torch.stack((torch.roll(x, 1, dims=-3) - x, torch.roll(x, 1, dims=-2) - x), dim=len(x.shape))


def finite_diff_adj(x):
    return (torch.roll(x[..., 0], -1, dims=-3) - x[..., 0]) + (torch.roll(x[..., 1], -1, dims=-2) - x[..., 1])


def finite_diff_gram(shape, dtype, device):
    gram = torch.zeros(shape, dtype=dtype, device=device)
    gram[0, 0, 0] = 4
    gram[0, 0, 1] = gram[0, 0, -1] = gram[0, 1, 0] = gram[0, -1, 0] = -1
    return torch.fft.rfft2(gram, dim=(-3, -2))


class TorchADMM:
    def __init__(self, psf, mu1=1e-6, mu2=1e-5, mu3=4e-5, tau=0.0001):
        self.mu1, self.mu2, self.mu3, self.tau = mu1, mu2, mu3, tau
        self.conv = TorchConvolver(psf, pad=False, norm="backward")
        self.psf = psf
        self.cdtype = torch.complex64 if psf.dtype == torch.float32 else torch.complex128
        self.PsiTPsi = finite_diff_gram(self.conv.padded_shape, psf.dtype, psf.device)
        self.data = None

    def set_data(self, data):
        while data.dim() < 5:
            data = data[None]
        self.data = data

    def reset(self):
        c = self.conv
        self.V = torch.zeros([1] + c.padded_shape, dtype=self.psf.dtype, device=self.psf.device)
        self.X = torch.zeros_like(self.V)
        self.U = torch.zeros_like(finite_diff(self.V))
        self.W = torch.zeros_like(self.X)
        self.HV = torch.zeros_like(self.X)
        self.PsiV = torch.zeros_like(self.U)
        self.xi = torch.zeros_like(self.V)
        self.eta = torch.zeros_like(self.U)
        self.rho = torch.zeros_like(self.X)
        self.R = (1.0 / (self.mu1 * torch.abs(c.Hadj * c.H) + self.mu2 * torch.abs(self.PsiTPsi) + self.mu3)).type(self.cdtype)
        self.Xdiv = 1.0 / (c._pad(torch.ones_like(self.psf)) + self.mu1)

    def update(self):
        c = self.conv
        self.U = soft_thresh(self.PsiV + self.eta / self.mu2, self.tau / self.mu2)
        self.X = self.Xdiv * (self.xi + self.mu1 * self.HV + c._pad(self.data))
        self.W = torch.maximum(self.rho / self.mu3 + self.V, torch.zeros_like(self.V))
        rk = (self.mu3 * self.W - self.rho) + finite_diff_adj(self.mu2 * self.U - self.eta) + c.deconvolve(self.mu1 * self.X - self.xi)
        s = tuple(c.padded_shape[-3:-1])
        self.V = torch.fft.irfft2(self.R * torch.fft.rfft2(rk, dim=(-3, -2)), dim=(-3, -2), s=s)
        self.HV = c.convolve(self.V)
        self.PsiV = finite_diff(self.V)
        self.xi += self.mu1 * (self.HV - self.X)
        self.eta += self.mu2 * (self.PsiV - self.U)
        self.rho += self.mu3 * (self.V - self.W)

    def form_image(self):
        img = self.conv._crop(self.V)
        img[img < 0] = 0
        return img

    def apply(self, n_iter):
        self.reset()
        for _ in range(n_iter):
            self.update()
        return self.form_image()[0]


class TorchFISTA:
    def __init__(self, psf, tk=1.0, lip_fact=1.8):
        self.conv = TorchConvolver(psf, pad=True, norm="ortho")
        self.psf, self.tk0, self.lip_fact = psf, tk, lip_fact
        self.data = None

    def set_data(self, data):
        while data.dim() < 5:
            data = data[None]
        self.data = data

    def reset(self):
        C = self.psf.shape[3]
        flat = self.psf.reshape(-1, C)
        start = (torch.max(flat, dim=0).values + torch.min(flat, dim=0).values) / 2
        self.x = torch.ones_like(self.psf[None, ...]) * start
        HH = (self.conv.Hadj * self.conv.H).reshape(-1, C)
        self.alpha = self.lip_fact / torch.max(torch.abs(HH), dim=0).values
        self.tk = self.tk0
        self.xk = self.x

    def update(self):
        self.x -= self.alpha * self.conv.deconvolve(self.conv.convolve(self.x) - self.data)
        xk = torch.clamp(self.x, min=0)
        tk = (1 + math.sqrt(1 + 4 * self.tk**2)) / 2
        self.x = xk + (self.tk - 1) / tk * (xk - self.xk)
        self.tk = tk
        self.xk = xk

    def apply(self, n_iter):
        self.reset()
        for _ in range(n_iter):
            self.update()
        return torch.clamp(self.x, min=0)[0]


ALGOS = {"admm": TorchADMM, "fista": TorchFISTA}
