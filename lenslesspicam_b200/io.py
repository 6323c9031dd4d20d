"""
GPU mirror of the input side of the path: ``lensless.utils.io.load_image`` / ``load_psf`` / ``load_data``
(reference: lensless/utils/io.py:20-207, 210-388, 389-600) and ``rgb2gray`` (utils/image.py:158-206).

The file is decoded on the host (``cv2.imread`` / ``np.load`` -- PNG inflate is not GPU work); the decoded uint8 / uint16
frame goes to the device once and everything the reference then does in numpy -- BGR->RGB, flips, background estimate and
subtraction, clipping, the antialiased resize (torchvision ``Resize(antialias=True)``, utils/image.py:62-67), L2 / max
normalisation, grayscale conversion -- runs as hand-written kernels in liblpc.so (csrc/lpc_io.cuh).  Outputs are float32
CUDA tensors in the reference's layout ``[depth, height, width, channels]``, ready for the solver classes.

Not mirrored (raise ``NotImplementedError``): Bayer / DNG input (``bayer=True``), 3-D ``.npy`` PSF stacks, integer outputs
(``return_float=False``), float64.
"""
import os
import warnings

import numpy as np
import torch

from . import _lib

IO_BGR, IO_FLIP_UD, IO_FLIP_LR, IO_SINGLE_PSF, IO_NORMALIZE_FIRST = 1, 2, 4, 8, 16


def _decode(fp):
    """io.py:109-112: the raw frame as the file stores it (uint8 / uint16, [H, W] or [H, W, C])"""
    assert os.path.isfile(fp)
    if "npy" in fp or "npz" in fp:
        img = np.load(fp)
    else:
        import cv2

        img = cv2.imread(fp, cv2.IMREAD_UNCHANGED)
    if img.dtype not in (np.uint8, np.uint16):
        raise NotImplementedError(f"raw frames must be uint8 or uint16, got {img.dtype}")
    return np.ascontiguousarray(img)


def _to_device(img, device):
    """one H2D copy of the decoded frame (pinned staging)"""
    t = torch.from_numpy(img.view(np.int16) if img.dtype == np.uint16 else img)
    return t.pin_memory().to(device, non_blocking=True)


def _new_shape(H, W, downsample=None, shape=None):
    """utils/image.py:48-53"""
    if shape is not None:
        return int(shape[-3]), int(shape[-2])
    new = np.array([H, W]) * (1 / downsample)
    return int(new[0]), int(new[1])


def _flags(flip, flip_ud, flip_lr, bgr_input, C):
    f = 0
    if bool(flip) ^ bool(flip_ud):
        f |= IO_FLIP_UD
    if bool(flip) ^ bool(flip_lr):
        f |= IO_FLIP_LR
    if bgr_input and C == 3:
        f |= IO_BGR
    return f


def _dev_index(device):
    device = torch.device(device)
    return device, (device.index if device.index is not None else torch.cuda.current_device())


def load_psf(fp, downsample=1, return_float=True, bg_pix=(5, 25), return_bg=False, flip=False, flip_ud=False, flip_lr=False,
             verbose=False, bayer=False, blue_gain=None, red_gain=None, dtype=np.float32, nbits_out=None, single_psf=False,
             shape=None, use_3d=False, bgr_input=True, force_rgb=False, torch_device="cuda"):
    """io.py:210-388 -> ``psf`` ``[1, h, w, C]`` float32 on ``torch_device`` (and ``bg`` ``[C]`` with ``return_bg``)."""
    if bayer or use_3d or not return_float:
        raise NotImplementedError("bayer / 3-D stack / integer-output PSF loading is not on the GPU path")
    if dtype not in (np.float32, "float32", None, torch.float32):
        raise NotImplementedError("the GPU loader is float32")
    raw = _decode(fp)
    if raw.ndim == 2 and force_rgb:
        raw = np.ascontiguousarray(np.stack([raw] * 3, axis=2))
    H, W = raw.shape[:2]
    C = 1 if raw.ndim == 2 else int(raw.shape[2])
    device, dev = _dev_index(torch_device)
    h, w = _new_shape(H, W, downsample, shape) if (downsample != 1 or shape is not None) else (H, W)
    d_raw = _to_device(raw, device)
    Cout = 1 if (single_psf and C > 1) else C
    if single_psf and C == 1:
        warnings.warn("Notice : single_psf has no effect for grayscale psf")
    psf = torch.empty((1, h, w, Cout), dtype=torch.float32, device=device)
    bg = torch.zeros((C,), dtype=torch.float32, device=device)
    flags = _flags(flip, flip_ud, flip_lr, bgr_input, C) | (IO_SINGLE_PSF if single_psf else 0)
    b0, b1 = (int(bg_pix[0]), int(bg_pix[1])) if bg_pix is not None else (0, 0)
    st = torch.cuda.current_stream(device).cuda_stream
    _lib.check(_lib.load().lpc_io_load_psf(dev, 8 * raw.dtype.itemsize, d_raw.data_ptr(), H, W, C, flags, b0, b1, h, w,
                                           psf.data_ptr(), bg.data_ptr(), st))
    if verbose:
        print(f"PSF {tuple(psf.shape)} float32, max {float(psf.max()):.3e}")
    return (psf, bg) if return_bg else psf


def load_image(fp, verbose=False, flip=False, flip_ud=False, flip_lr=False, bayer=False, black_level=None, blue_gain=None,
               red_gain=None, ccm=None, back=None, nbits_out=None, as_4d=False, downsample=None, bg=None, return_float=False,
               shape=None, dtype=None, normalize=True, bgr_input=True, torch_device="cuda", _normalize_first=False):
    """io.py:20-207, float output only -> ``[h, w, C]`` (``[1, h, w, C]`` with ``as_4d``) float32 on ``torch_device``.
    ``bg``: the ``[C]`` background level returned by :func:`load_psf` (values <= 1: scaled by the frame's max_val as in
    io.py:167-168)."""
    if bayer or back is not None:
        raise NotImplementedError("bayer input is not on the GPU path")
    if not return_float:
        raise NotImplementedError("the GPU loader returns float32 (return_float=True)")
    raw = _decode(fp)
    H, W = raw.shape[:2]
    C = 1 if raw.ndim == 2 else int(raw.shape[2])
    device, dev = _dev_index(torch_device)
    h, w = _new_shape(H, W, downsample, shape) if (downsample is not None or shape is not None) else (H, W)
    d_raw = _to_device(raw, device)
    out = torch.empty((h, w, C), dtype=torch.float32, device=device)
    bg_ptr = None
    if bg is not None:
        bg = torch.as_tensor(bg, dtype=torch.float32).to(device).reshape(-1)
        if bg.numel() == 1:
            bg = bg.expand(C)
        bg = bg.contiguous()
        bg_ptr = bg.data_ptr()
    flags = _flags(flip, flip_ud, flip_lr, bgr_input, C) | (IO_NORMALIZE_FIRST if _normalize_first else 0)
    st = torch.cuda.current_stream(device).cuda_stream
    _lib.check(_lib.load().lpc_io_load_image(dev, 8 * raw.dtype.itemsize, d_raw.data_ptr(), H, W, C, flags, bg_ptr, h, w,
                                             1 if normalize else 0, out.data_ptr(), st))
    if verbose:
        print(f"image {tuple(out.shape)} float32, max {float(out.max()):.3e}")
    return out[None] if as_4d else out


def rgb2gray(rgb, weights=None, keepchanneldim=True):
    """utils/image.py:158-206 (the numpy branch's weights 0.299 / 0.587 / 0.114, which is what load_data applies)."""
    assert rgb.is_cuda and rgb.shape[-1] == 3 and weights is None
    rgb = rgb.to(torch.float32).contiguous()
    out = torch.empty(tuple(rgb.shape[:-1]) + ((1,) if keepchanneldim else ()), dtype=torch.float32, device=rgb.device)
    device, dev = _dev_index(rgb.device)
    st = torch.cuda.current_stream(device).cuda_stream
    _lib.check(_lib.load().lpc_io_rgb2gray(dev, rgb.data_ptr(), out.data_ptr(), rgb.numel() // 3, st))
    return out


def load_data(psf_fp, data_fp, background_fp=None, return_bg=False, remove_background=True, return_float=True, downsample=None,
              bg_pix=(5, 25), plot=False, flip=False, flip_ud=False, flip_lr=False, bayer=False, blue_gain=None, red_gain=None,
              gamma=None, gray=False, dtype=None, single_psf=False, shape=None, use_torch=True, torch_device="cuda",
              normalize=False, bgr_input=True):
    """io.py:389-600 -> ``(psf, data)`` float32 CUDA tensors ``[1, h, w, C]`` (``(psf, data, bg)`` with ``return_bg``).
    The measurement is normalised at full resolution and only then resized to the PSF's shape, as the reference does
    (io.py:509-513, 545-547)."""
    assert os.path.isfile(psf_fp) and os.path.isfile(data_fp)
    if shape is None:
        assert downsample is not None
    if background_fp is not None:
        raise NotImplementedError("a measured background frame (background_fp) is not on the GPU path")
    if dtype not in (None, "float32", np.float32):
        raise NotImplementedError("the GPU loader is float32")
    if plot:
        raise NotImplementedError("plotting is out of the engine's scope")
    res = load_psf(psf_fp, downsample=downsample, return_float=return_float, bg_pix=bg_pix, return_bg=bg_pix is not None, flip=flip,
                   flip_ud=flip_ud, flip_lr=flip_lr, bayer=bayer, single_psf=single_psf, shape=shape, bgr_input=bgr_input,
                   torch_device=torch_device)
    psf, bg = res if bg_pix is not None else (res, None)
    data = load_image(data_fp, flip=flip, flip_ud=flip_ud, flip_lr=flip_lr, bayer=bayer, bg=bg, as_4d=True, return_float=return_float,
                      shape=psf.shape, normalize=normalize, bgr_input=bgr_input, torch_device=torch_device,
                      _normalize_first=(shape is None))  # with an explicit shape the reference resizes inside load_image, then normalises
    if data.shape[3] > 1 and psf.shape[3] == 1:
        warnings.warn("Warning: loaded a grayscale PSF with RGB data. Repeating PSF across channels.")
        psf = psf.repeat(1, 1, 1, data.shape[3])
    if data.shape[3] == 1 and psf.shape[3] > 1:
        warnings.warn("Warning: loaded a RGB PSF with grayscale data. Repeating data across channels.")
        data = data.repeat(1, 1, 1, psf.shape[3])
    if gray and psf.shape[3] == 3:
        psf, data = rgb2gray(psf), rgb2gray(data)
    if return_bg:
        return psf, data, (bg if bg is not None else torch.zeros((psf.shape[3],), device=psf.device))
    return psf, data
