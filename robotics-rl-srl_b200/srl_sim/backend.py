"""
Backend selection and buffer plumbing for the host-side env classes.

The product backend is the sm_100a library on a CUDA device; PyTorch is used only to own device
memory and streams.  There is NO CPU fallback: :func:`default_backend` raises when the CUDA
library or a GPU is missing.  :func:`use_library` lets a caller (the test-suite driving the CPU
oracle through the same ABI) substitute another library explicitly; nothing in this package calls
it.
"""
import numpy as np

from . import _abi

_override = None  # (SimLibrary, device) installed by use_library()


class Backend(object):
    """A loaded srl_sim library plus the device its handles live on."""

    def __init__(self, library, device):
        self.library = library
        self.device = int(device)
        self.on_gpu = self.device >= 0
        if self.on_gpu:
            import torch
            if not torch.cuda.is_available():
                raise _abi.SimError("a CUDA device is required (torch.cuda.is_available() is False); "
                                    "there is no CPU fallback")
            self.torch = torch
            self.torch_device = torch.device("cuda", self.device)

    # ---- buffers -----------------------------------------------------------------------------
    def empty(self, shape, dtype):
        if self.on_gpu:
            return self.torch.empty(shape, dtype=_TORCH_DTYPES(self.torch)[np.dtype(dtype).name],
                                    device=self.torch_device)
        return np.empty(shape, dtype=dtype)

    def zeros(self, shape, dtype):
        buf = self.empty(shape, dtype)
        if self.on_gpu:
            buf.zero_()
        else:
            buf[...] = 0
        return buf

    def from_host(self, array, dtype=None):
        arr = np.ascontiguousarray(array, dtype=dtype)
        if self.on_gpu:
            return self.torch.from_numpy(arr).to(self.torch_device)
        return arr

    def to_host(self, buf):
        if self.on_gpu:
            return buf.detach().cpu().numpy()
        return buf

    def stream(self):
        if self.on_gpu:
            return self.torch.cuda.current_stream(self.torch_device).cuda_stream
        return None

    def make_sim(self, env_kind, num_envs, seed=0, model_blob=None, **cfg):
        return _abi.Sim(self.library, env_kind, num_envs, self.device, seed=seed, model_blob=model_blob, **cfg)


def _TORCH_DTYPES(torch):
    return {"float32": torch.float32, "float64": torch.float64, "int32": torch.int32, "uint8": torch.uint8,
            "int64": torch.int64}


def use_library(library, device):
    """Install an explicit (library, device) pair; pass ``None`` to restore the CUDA default."""
    global _override
    _override = None if library is None else (library, device)


def default_backend(device=None):
    if _override is not None:
        lib, dev = _override
        return Backend(lib, dev if device is None else device)
    return Backend(_abi.load_cuda_library(), 0 if device is None else device)
