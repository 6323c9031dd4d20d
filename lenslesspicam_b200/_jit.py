"""
On-demand specialisation of the fused engine for padded sizes outside the compiled menu.

The fused kernels are templates on a compile-time radix sequence (csrc/lpc_fft.cuh: every stride and twiddle offset a
constant); ``liblpc.so`` instantiates them for the padded sizes of the usual sensor dimensions.  A size outside that menu
used to fall back to the cuFFT engine (~2x slower).  Here the first solver of such a size compiles ONE extra row entry and
ONE extra column entry into a small library of its own (``nvcc``, ~15 s, cached under ``lenslesspicam_b200/_jit/``) and
creates its plans there; every other C-ABI call is a virtual call on the plan, so the rest of the package is unchanged.
Without ``nvcc`` (or with ``LPC_JIT=0``) the cuFFT engine serves the size and the mirror warns once.
"""
import ctypes
import os
import shutil
import subprocess
import sys
import tempfile

from . import _lib

JIT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_jit")
RADICES = (16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2)
_loaded = {}


def _min_pass_multiset(N, max_passes=4):
    """fewest radices (from RADICES) whose product is N; None if N is not reachable within max_passes"""
    best = {1: ()}
    frontier = [1]
    for _ in range(max_passes):
        nxt = []
        for n in frontier:
            for r in RADICES:
                m = n * r
                if N % m == 0 and m not in best:
                    best[m] = best[n] + (r,)
                    nxt.append(m)
        if N in best:
            return best[N]
        frontier = nxt
    return best.get(N)


def plan_rows(N):
    """row sequence: fewest passes, the largest odd radix first (conflict-free first-pass stores without padding,
    csrc/lpc_fft.cuh), the rest descending; every pass must keep <= 32 complex values per thread (256 threads)"""
    ms = _min_pass_multiset(N)
    if ms is None or len(ms) < 2:
        return None
    rest = sorted(ms, reverse=True)
    odd = [r for r in rest if r % 2]
    seq = rest
    if odd:
        rest.remove(max(odd))
        seq = [max(odd)] + rest
    for r in seq:
        nb = N // r
        if -(-nb // 256) * r > 32:
            return None
    return seq


def plan_cols(N):
    """column sequence: ends in a small radix (that pass is fused with the spectral update in registers)"""
    ms = _min_pass_multiset(N)
    if ms is None or len(ms) < 2:
        return None
    seq = sorted(ms, reverse=True)
    if seq[-1] > 10:  # split the smallest radix so that the in-register stage stays small
        split = {16: (4, 4), 15: (5, 3), 12: (4, 3)}[seq[-1]]
        seq = seq[:-1] + list(split)
        if len(seq) > 4:
            return None
    head, last = seq[:-1], seq[-1]
    even = [r for r in head if r % 2 == 0]
    if len(even) == 1 and last % 2 and len(head) >= 2:
        # one radix carries every factor of two: first, so that the two outer passes run warp-locally (lpc_fft.cuh: IpWl)
        head.remove(even[0])
        return [even[0]] + head + [last]
    odd = [r for r in head if r % 2]
    if odd:
        o = max(odd)
        head.remove(o)
        seq = [o] + head + [last]
    return seq


def _nvcc():
    return shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)


def library_for(Wp, Hp, verbose=True):
    """the library whose lpc_plan_create serves (Wp, Hp) on the fused engine: liblpc.so, or a cached / freshly built
    specialisation; liblpc.so (cuFFT engine for this size) when neither is possible"""
    lib = _lib.load()
    hr, hc = ctypes.c_int(), ctypes.c_int()
    lib.lpc_menu_query(int(Wp), int(Hp), ctypes.byref(hr), ctypes.byref(hc))
    if hr.value and hc.value:
        return lib
    key = (int(Wp), int(Hp))
    if key in _loaded:
        return _loaded[key]
    if os.environ.get("LPC_JIT", "1") == "0":
        return lib
    rows, cols = plan_rows(Wp), plan_cols(Hp)
    if rows is None or cols is None:
        return lib
    out = os.path.join(JIT_DIR, f"liblpc_{Wp}x{Hp}.so")
    if not os.path.exists(out):
        nvcc = _nvcc()
        if nvcc is None:
            return lib
        os.makedirs(JIT_DIR, exist_ok=True)
        tmp = tempfile.mkdtemp(prefix="lpc_jit_", dir=JIT_DIR)
        try:
            with open(os.path.join(tmp, "lpc_menu_extra.h"), "w") as f:
                f.write(f"#define LPC_ROW_MENU_EXTRA(X) X({Wp}, {', '.join(map(str, rows))})\n")
                f.write(f"#define LPC_COL_MENU_EXTRA(X) X({Hp}, {', '.join(map(str, cols))})\n")
            cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
            cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                   "-shared", "-o", os.path.join(tmp, "lib.so"), os.path.join(_lib.CSRC, "lpc_api.cu"), "-L" + cuda_lib, "-lcufft",
                   "-Xlinker", "-rpath=" + cuda_lib, "-DLPC_MENU_ONLY_EXTRA", "-DLPC_HAVE_MENU_EXTRA_H", "-I" + tmp]
            if verbose:
                print(f"lenslesspicam_b200: compiling the fused engine for padded size {Hp}x{Wp} "
                      f"(rows {rows}, columns {cols}); cached in {JIT_DIR}", file=sys.stderr)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if verbose:
                    print("lenslesspicam_b200: on-demand build failed, the cuFFT engine serves this size\n" + res.stderr[-2000:], file=sys.stderr)
                return lib
            os.replace(os.path.join(tmp, "lib.so"), out)  # atomic: concurrent ranks may race to build the same size
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    try:
        jl = ctypes.CDLL(out)
        jl.lpc_plan_create.restype, jl.lpc_plan_create.argtypes = _lib.PROTOTYPES["lpc_plan_create"]
        jl.lpc_menu_query.restype, jl.lpc_menu_query.argtypes = _lib.PROTOTYPES["lpc_menu_query"]
        jl.lpc_last_error.restype, jl.lpc_last_error.argtypes = _lib.PROTOTYPES["lpc_last_error"]
    except OSError:
        return lib
    _loaded[key] = jl
    return jl
