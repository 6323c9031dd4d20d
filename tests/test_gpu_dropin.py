"""
Drop-in proof: the REFERENCE'S OWN CALLERS executed against the mirror classes.

``baseline/_ref`` (the unmodified reference, pip-installed by ``__graft_entry__.build()``; it travels to the GPU box) ships
the reference's tests, profile scripts and recon scripts.  Their code is loaded from there and run with the name
``lensless`` bound to a shim whose classes / functions are this package's:

* ``test/test_convolver.py``  -- test_pad_torch, test_conv_torch, test_conv_torch_2, unmodified, with torch's default device
  set to CUDA so the tensors they create live where the engine works;
* ``test/test_algos.py``      -- test_recon_torch for all six solver classes and test_set_initial_est_unrolled, unmodified;
  their ``load_data("data/psf/tape_rgb.png", ...)`` is served by the GPU loader from the committed crop of that very file;
* ``profile/admm.py:86-108``  -- the "PyTorch GPU" block, source lines executed verbatim;
* ``scripts/recon/admm.py``   -- the body of ``admm(config)`` (the hydra entry point: the decorator is the only thing
  removed) with a config object holding configs/recon/defaults.yaml's values;
* ``scripts/recon/gradient_descent.py`` -- same for ``gradient_descent(config)``.

Nothing here checks numbers (the parity tests do); it checks that the reference's callers run, unedited, on the mirror.
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "baseline", "_ref")
PSF_CROP = os.path.join(GOLDEN, "real_psf_crop.png")
DATA_CROP = os.path.join(GOLDEN, "real_data_crop.png")


def _need_ref():
    if not os.path.isdir(os.path.join(REF, "test")):
        pytest.skip("baseline/_ref (the unmodified reference) is not installed")


def _shim_load_data(psf_fp, data_fp, **kw):
    """the reference's bundled files are not on the GPU box: serve the committed crops of the same two frames"""
    import lenslesspicam_b200 as L

    assert psf_fp.endswith("data/psf/tape_rgb.png") and data_fp.endswith("data/raw_data/thumbs_up_rgb.png")
    dtype = kw.pop("dtype", None)
    kw.pop("use_torch", None)
    kw["torch_device"] = "cuda"
    for k in ("bayer", "blue_gain", "red_gain", "gamma", "background_fp"):
        if not kw.get(k):
            kw.pop(k, None)
    psf, data = L.load_data(PSF_CROP, DATA_CROP, **kw)
    if dtype == "float64":
        psf, data = psf.double(), data.double()
    return psf, data


@pytest.fixture
def shim(monkeypatch):
    """modules named lensless, lensless.recon.*, lensless.utils.* whose members are the mirror's"""
    import lenslesspicam_b200 as L

    mods = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        mods[name] = m
        return m

    solvers = dict(ADMM=L.ADMM, FISTA=L.FISTA, GradientDescent=L.GradientDescent, NesterovGradientDescent=L.NesterovGradientDescent,
                   UnrolledADMM=L.UnrolledADMM, UnrolledFISTA=L.UnrolledFISTA)
    top = mod("lensless", **solvers)
    top.__path__ = []
    mod("lensless.recon").__path__ = []
    mod("lensless.recon.rfft_convolve", RealFFTConvolve2D=L.RealFFTConvolve2D)
    mod("lensless.recon.admm", ADMM=L.ADMM, soft_thresh=L.soft_thresh, finite_diff=L.finite_diff, finite_diff_adj=L.finite_diff_adj,
        finite_diff_gram=L.finite_diff_gram)
    mod("lensless.recon.gd", FISTA=L.FISTA, GradientDescent=L.GradientDescent, NesterovGradientDescent=L.NesterovGradientDescent)
    mod("lensless.recon.unrolled_admm", UnrolledADMM=L.UnrolledADMM)
    mod("lensless.utils").__path__ = []
    mod("lensless.utils.io", load_data=_shim_load_data, load_image=L.load_image, load_psf=L.load_psf)
    mod("lensless.utils.plot", plot_image=lambda *a, **k: None)
    for name, m in mods.items():
        monkeypatch.setitem(sys.modules, name, m)
    return mods


def _load_module(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_test_convolver_on_the_mirror(shim):
    _need_ref()
    prev = torch.get_default_device()
    torch.set_default_device("cuda")
    try:
        m = _load_module(os.path.join(REF, "test", "test_convolver.py"), "ref_test_convolver")
        assert m.RealFFTConvolve2D.__module__.startswith("lenslesspicam_b200")
        m.test_pad_torch()
        m.test_conv_torch()
        m.test_conv_torch_2()
    finally:
        torch.set_default_device(prev)


def test_reference_test_algos_on_the_mirror(shim):
    _need_ref()
    import warnings

    m = _load_module(os.path.join(REF, "test", "test_algos.py"), "ref_test_algos")
    assert [a.__module__.split(".")[0] for a in m.standard_algos + m.trainable_algos] == ["lenslesspicam_b200"] * 6
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # 28x38 at downsample 16 is outside the fused menu: engine v1 serves it, with a warning
        for algorithm in m.standard_algos + m.trainable_algos:
            m.test_recon_torch(algorithm)
        for algorithm in m.trainable_algos:
            m.test_set_initial_est_unrolled(algorithm)


def test_reference_profile_admm_gpu_block_on_the_mirror(shim, tmp_path, capsys):
    """profile/admm.py:86-108 verbatim (its constants from :9-16, n_trials reduced)"""
    _need_ref()
    import time

    import lenslesspicam_b200 as L

    src = open(os.path.join(REF, "profile", "admm.py")).read().splitlines()
    start = next(i for i, line in enumerate(src) if line.strip() == '""" PyTorch GPU """')
    block = "\n".join(src[start:])
    assert 'torch_device="cuda"' in block and "recon = ADMM(psf, dtype=dtype)" in block
    ns = dict(load_data=_shim_load_data, ADMM=L.ADMM, time=time, psf_fp="data/psf/tape_rgb.png",
              data_fp="data/raw_data/thumbs_up_rgb.png", n_iter=5, downsample=4, gray=True, dtype="float32", n_trials=3, save=False)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(compile(block, "profile/admm.py[GPU block]", "exec"), ns)
    res = ns["res"]
    assert res.is_cuda and tuple(res.shape) == (1, 114, 152, 1) and "LenslessPiCam, PyTorch GPU (avg)" in capsys.readouterr().out


class Cfg(dict):
    """minimal stand-in for the OmegaConf node hydra hands the scripts: item and attribute access, ** unpacking"""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    @staticmethod
    def wrap(d):
        return Cfg({k: Cfg.wrap(v) if isinstance(v, dict) else v for k, v in d.items()})


def _defaults():
    """configs/recon/defaults.yaml of the reference (v1.0.8), the keys the two scripts read; save / plot off"""
    return Cfg.wrap({
        "input": {"psf": "data/psf/tape_rgb.png", "data": "data/raw_data/thumbs_up_rgb.png", "dtype": "float32", "original": None,
                  "background": None},
        "torch": True, "torch_device": "cuda",
        "preprocess": {"normalize": True, "downsample": 4, "shape": None, "flip": False, "bayer": False, "blue_gain": None,
                       "red_gain": None, "single_psf": False, "gray": False, "bg_pix": [5, 25]},
        "display": {"disp": -1, "plot": False, "gamma": None},
        "save": False,
        "gradient_descent": {"n_iter": 6, "method": "fista", "nesterov": {"p": 0, "mu": 0.9}, "fista": {"tk": 1}},
        "admm": {"n_iter": 5, "mu1": 1e-6, "mu2": 1e-5, "mu3": 4e-5, "tau": 0.0001, "unrolled": False, "checkpoint_fp": None,
                 "pre_process_model": {"network": None, "depth": 2}, "post_process_model": {"network": None, "depth": 2}},
    })


def _script_function(path, fname):
    """the FunctionDef `fname` of a reference script, compiled WITHOUT its @hydra.main decorator and without the
    script's top-level imports (hydra and matplotlib are not in the image)"""
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == fname)
    fn.decorator_list = []
    mod = ast.Module(body=[fn], type_ignores=[])
    ast.fix_missing_locations(mod)
    return compile(mod, path, "exec")


def _script_namespace():
    import pathlib as plib
    import time

    import lenslesspicam_b200 as L

    return dict(os=os, time=time, plib=plib, np=np, plt=None, to_absolute_path=lambda p: p, load_data=_shim_load_data,
                load_image=L.load_image, plot_image=lambda *a, **k: None, ADMM=L.ADMM, FISTA=L.FISTA, GradientDescent=L.GradientDescent,
                NesterovGradientDescent=L.NesterovGradientDescent, GradientDescentUpdate=L.GradientDescentUpdate)


def test_reference_scripts_recon_admm_body_on_the_mirror(shim, capsys):
    _need_ref()
    import warnings

    ns = _script_namespace()
    exec(_script_function(os.path.join(REF, "scripts", "recon", "admm.py"), "admm"), ns)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ns["admm"](_defaults())  # hydra passes unrolled / checkpoint_fp / *_process_model to the constructor: swallowed
    out = capsys.readouterr().out
    assert "Setup time" in out and "Processing time" in out


def test_reference_scripts_recon_gradient_descent_body_on_the_mirror(shim, capsys):
    _need_ref()
    import warnings

    path = os.path.join(REF, "scripts", "recon", "gradient_descent.py")
    ns = _script_namespace()
    exec(_script_function(path, "gradient_descent"), ns)
    for method in ("vanilla", "nesterov", "fista"):
        cfg = _defaults()
        cfg["gradient_descent"]["method"] = method
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns["gradient_descent"](cfg)
    assert capsys.readouterr().out.count("Processing time") == 3
