"""numpy emulation of the engine's fused ADMM formulation (half-updated duals, 4 FFTs, spectral merge).
Test-side diagnostic only: separates formulation round-off from kernel bugs when a GPU parity test fails."""
import numpy as np
from scipy import fft

import lensless_oracle as orc


def fused_admm(psf, data, n_iter, mu1=1e-6, mu2=1e-5, mu3=4e-5, tau=1e-4, f=np.float32, workers=8):
    D, h, w, C = psf.shape
    Hp, Wp, sy, sx = orc.padded_geometry(h, w)
    c = np.complex64 if f == np.float32 else np.complex128
    P = np.zeros((C, Hp, Wp), f)
    P[:, sy:sy + h, sx:sx + w] = psf[0].transpose(2, 0, 1)
    H = fft.rfft2(P, workers=workers).astype(c)
    ky = np.arange(Hp)[:, None]
    kx = np.arange(Wp // 2 + 1)[None, :]
    S = np.exp(2j * np.pi * (((ky * (Hp // 2)) % Hp) / Hp + ((kx * (Wp // 2)) % Wp) / Wp)).astype(c)
    G = 4 - 2 * np.cos(2 * np.pi * ky / Hp) - 2 * np.cos(2 * np.pi * kx / Wp)
    absH2 = (H.real ** 2 + H.imag ** 2).astype(np.float64)
    R = (1.0 / (mu1 * absH2 + mu2 * G + mu3) / (Hp * Wp)).astype(f)
    y = np.zeros((C, Hp, Wp), f)
    y[:, sy:sy + h, sx:sx + w] = data[0].transpose(2, 0, 1)
    xdiv = np.full((C, Hp, Wp), f(1) / (f(0) + f(mu1)), f)
    xdiv[:, sy:sy + h, sx:sx + w] = f(1) / (f(1) + f(mu1))
    mu1, mu2, mu3 = f(mu1), f(mu2), f(mu3)
    thr = f(tau / float(mu2)) if False else f(tau / 1e-5 * (1e-5 / float(mu2)))
    thr = f(float(tau) / float(np.float64(mu2)))
    V = np.zeros((C, Hp, Wp), f)
    HV = V.copy(); xi_p = V.copy(); rho_p = V.copy(); e0_p = V.copy(); e1_p = V.copy()
    Hs = (S * H).astype(c); Hsa = (S * np.conj(H)).astype(c)

    def soft(x):
        return np.sign(x) * np.maximum(np.abs(x) - thr, f(0))

    for _ in range(n_iter):
        xi = xi_p + mu1 * HV
        psi0 = np.roll(V, 1, 1) - V
        psi1 = np.roll(V, 1, 2) - V
        e0 = e0_p + mu2 * psi0
        e1 = e1_p + mu2 * psi1
        rho = rho_p + mu3 * V
        u0 = soft(psi0 + e0 / mu2)
        u1 = soft(psi1 + e1 / mu2)
        x = xdiv * (xi + mu1 * HV + y)
        wv = np.maximum(rho / mu3 + V, f(0))
        xi_p = xi - mu1 * x
        e0_p = e0 - mu2 * u0
        e1_p = e1 - mu2 * u1
        rho_p = rho - mu3 * wv
        t0 = -e0_p; t1 = -e1_p
        a = -rho_p + ((np.roll(t0, -1, 1) - t0) + (np.roll(t1, -1, 2) - t1))
        A = fft.rfft2(a, workers=workers).astype(c)
        Xi = fft.rfft2(xi_p, workers=workers).astype(c)
        Vh = (R * (A - Hsa * Xi)).astype(c)
        HVh = (Hs * Vh).astype(c)
        V = (fft.irfft2(Vh, s=(Hp, Wp), workers=workers) * (Hp * Wp)).astype(f)
        HV = (fft.irfft2(HVh, s=(Hp, Wp), workers=workers) * (Hp * Wp)).astype(f)
    return np.maximum(V[:, sy:sy + h, sx:sx + w], 0).transpose(1, 2, 0)[None]
