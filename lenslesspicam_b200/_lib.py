"""
ctypes binding of liblpc.so (C-ABI declared in include/lpc.h).

There is NO fallback: if the CUDA extension is missing or fails to load, every
entry point of the package raises.  PyTorch is used only for device memory,
streams and (in bench.py) torch.distributed.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# LPC_LIB_PATH: load another build of the same C-ABI (A/B experiments: e.g. -DLPC_PACKED2=0)
LIB_PATH = os.environ.get("LPC_LIB_PATH") or os.path.join(_HERE, "liblpc.so")
CSRC = os.path.join(_HERE, "csrc")

OK = 0
ALGO_CONV, ALGO_ADMM, ALGO_GD, ALGO_NESTEROV, ALGO_FISTA = 0, 1, 2, 3, 4
F32, F64 = 0, 1
NORMS = {"backward": 0, "ortho": 1, "forward": 2}
STATE_IMAGE_EST, STATE_FORWARD, STATE_XI, STATE_RHO, STATE_ETA0, STATE_ETA1 = range(6)
FLAG_NONE, FLAG_NO_GRAPH, FLAG_ENGINE_V1, FLAG_PSF_PER_SAMPLE, FLAG_PNP = 0, 1, 4, 8, 16
OPT_KEEP_X0, OPT_B2_PREFETCH, OPT_PNP_USE_DUAL = 1, 2, 3

_lock = threading.Lock()
_lib = None

_vp = ctypes.c_void_p
_i = ctypes.c_int
_d = ctypes.c_double
_ip = ctypes.POINTER(ctypes.c_int)
_dp = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes); must list every symbol include/lpc.h declares
PROTOTYPES = {
    "lpc_version": (_i, []),
    "lpc_error_name": (ctypes.c_char_p, [_i]),
    "lpc_last_error": (ctypes.c_char_p, [_vp]),
    "lpc_next_fast_len": (_i, [_i]),
    "lpc_geometry": (_i, [_i, _i, _ip, _ip, _ip, _ip]),
    "lpc_plan_create": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
    "lpc_plan_destroy": (_i, [_vp]),
    "lpc_menu_query": (_i, [_i, _i, _ip, _ip]),
    "lpc_set_psf": (_i, [_vp, _vp, _vp]),
    "lpc_set_admm_params": (_i, [_vp, _d, _d, _d, _d, _vp]),
    "lpc_set_gd_params": (_i, [_vp, _d, _d, _d, _vp]),
    "lpc_set_admm_schedule": (_i, [_vp, _i, _dp, _dp, _dp, _dp, _vp]),
    "lpc_set_fista_schedule": (_i, [_vp, _i, _dp, _dp, _vp]),
    "lpc_set_data": (_i, [_vp, _vp, _i, _i64p, _vp, _i64p, _vp, _vp]),
    "lpc_set_data_host": (_i, [_vp, _vp, _i, _vp]),
    "lpc_reset": (_i, [_vp, _vp, _vp]),
    "lpc_run": (_i, [_vp, _i, _vp]),
    "lpc_get_image": (_i, [_vp, _vp, _vp]),
    "lpc_get_image_host": (_i, [_vp, _vp, _vp]),
    "lpc_get_state": (_i, [_vp, _i, _vp, _vp]),
    "lpc_get_H": (_i, [_vp, _vp, _vp]),
    "lpc_get_alpha": (_i, [_vp, _vp, _vp]),
    "lpc_convolve": (_i, [_vp, _vp, _i64p, _vp, _i, _i, _vp]),
    "lpc_reconstruction_error": (_i, [_vp, _vp, _vp, _i, _i, _d, _vp, _vp]),
    "lpc_set_image_estimate": (_i, [_vp, _vp, _vp]),
    "lpc_set_option": (_i, [_vp, _i, _d]),
    "lpc_io_load_psf": (_i, [_i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "lpc_io_load_image": (_i, [_i, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "lpc_io_rgb2gray": (_i, [_i, _vp, _vp, ctypes.c_int64, _vp]),
    "lpc_pnp_begin": (_i, [_vp, _vp, _vp]),
    "lpc_pnp_finish": (_i, [_vp, _vp, _vp]),
    "lpc_op_soft_thresh": (_i, [_i, _i, _vp, _vp, ctypes.c_int64, _d, _vp]),
    "lpc_op_finite_diff": (_i, [_i, _i, _vp, _vp, ctypes.c_int64, _i, _i, _i, _vp]),
    "lpc_op_finite_diff_adj": (_i, [_i, _i, _vp, _vp, ctypes.c_int64, _i, _i, _i, _vp]),
    "lpc_op_finite_diff_gram": (_i, [_i, _i, _vp, _i, _i, _i, _i, _vp]),
    "lpc_bytes_per_iter": (_i, [_vp, ctypes.POINTER(_d)]),
    "lpc_launches_per_iter": (_i, [_vp, _ip, _ip]),
    "lpc_profile": (_i, [_vp, _i, _vp, ctypes.POINTER(_d), _ip]),
    "lpc_phase_name": (ctypes.c_char_p, [_vp, _i]),
    "lpc_engine_version": (_i, [_vp]),
    "lpc_plan_bytes": (_i, [_vp, ctypes.POINTER(ctypes.c_size_t)]),
}


class LpcError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def build(verbose=False):
    """Compile liblpc.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building liblpc.so failed")
    return LIB_PATH


def load():
    """Load liblpc.so and bind the prototypes.  Raises if the library is absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or `make -C lenslesspicam_b200/csrc`."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(code, plan=None):
    if code == OK:
        return
    lib = load()
    msg = lib.lpc_last_error(plan)
    msg = msg.decode() if msg else ""
    name = lib.lpc_error_name(code).decode()
    raise LpcError(code, f"{name}: {msg}")


def strides_arg(strides):
    if strides is None:
        return None
    arr = (ctypes.c_int64 * 5)(*[int(s) for s in strides])
    return arr
