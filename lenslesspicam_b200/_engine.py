"""Small host-side helpers shared by the mirror classes (streams, dtype codes)."""
import torch

from . import _lib


def torch_dtype_code(dtype):
    if dtype == torch.float32:
        return _lib.F32
    if dtype == torch.float64:
        return _lib.F64
    raise ValueError(f"Unsupported dtype : {dtype}")


class EngineStream:
    """Context manager yielding the raw cudaStream_t the engine should enqueue on.

    The engine captures its iteration loop into a CUDA graph, which the legacy
    default stream cannot do; when torch's current stream IS the default stream
    the work goes to a private side stream, ordered after / before the current
    stream with events (no host synchronisation either way).
    """

    def __init__(self, device):
        self.device = torch.device(device)
        self._side = None
        self._cur = None
        self._use = None

    def __enter__(self):
        cur = torch.cuda.current_stream(self.device)
        self._cur = cur
        if cur.cuda_stream != 0:
            self._use = cur
        else:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            self._side.wait_stream(cur)
            self._use = self._side
        return self._use.cuda_stream

    def __exit__(self, *exc):
        if self._use is not self._cur:
            self._cur.wait_stream(self._use)
        return False
