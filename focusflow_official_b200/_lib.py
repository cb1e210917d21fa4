"""ctypes binding of libffcorr.so (the C ABI declared in include/ffcorr.h).

There is NO CPU fallback and no alternative backend: if the shared library is missing the
import of any operator fails loudly with the build command.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libffcorr.so")
CSRC = os.path.join(_HERE, "csrc")

PREC_FP16, PREC_FP32, PREC_BF16X3, PREC_TF32 = 0, 1, 2, 3
PRECISIONS = {"fp16": PREC_FP16, "fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "tf32": PREC_TF32}
SAMPLERS = {"cpu": 0, "cuda": 1}   # FFCORR_SAMPLER_ATEN_CPU / FFCORR_SAMPLER_ATEN_CUDA
MAX_LEVELS = 8

_lock = threading.Lock()
_lib = None

_vp = ctypes.c_void_p
_i = ctypes.c_int
_PROTOS = {
    # name: (restype, argtypes)
    "ffcorr_version": (_i, []),
    "ffcorr_last_error": (ctypes.c_char_p, []),
    "ffcorr_device_info": (_i, [ctypes.POINTER(_i)] * 3),
    "ffcorr_set_l2_fetch_granularity": (_i, [_i]),
    "ffcorr_get_l2_fetch_granularity": (_i, [ctypes.POINTER(_i)]),
    "ffcorr_volume_workspace_bytes": (ctypes.c_size_t, [_i, _i, _i, _i, _i]),
    "ffcorr_volume_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, ctypes.c_size_t, _vp]),
    "ffcorr_volume_scaled_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, ctypes.c_float, _vp, ctypes.c_size_t, _vp]),
    "ffcorr_pyramid_f32": (_i, [ctypes.POINTER(_vp), _i, ctypes.c_int64, _i, _i, _vp]),
    "ffcorr_lookup_f32": (_i, [ctypes.POINTER(_vp), _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ffcorr_tiled_map_elems": (ctypes.c_int64, [_i, _i, _i]),
    "ffcorr_tiled_supported": (_i, [_i, _i, _i]),
    "ffcorr_volume_tiled_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, ctypes.c_size_t, _vp]),
    "ffcorr_pyramid_tiled_f32": (_i, [ctypes.POINTER(_vp), _i, ctypes.c_int64, _i, _i, _vp]),
    "ffcorr_build_tiled_f32": (_i, [_vp, _vp, ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _vp, ctypes.c_size_t, _vp]),
    "ffcorr_stage_operands_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, ctypes.c_size_t, _vp]),
    "ffcorr_build_tiled_chunk_f32": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _i, _vp, ctypes.c_size_t, _vp]),
    "ffcorr_lookup_tiled_chunk_f32": (_i, [ctypes.POINTER(_vp), _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ffcorr_lookup_tiled_f32": (_i, [ctypes.POINTER(_vp), _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ffcorr_untile_f32": (_i, [_vp, _vp, ctypes.c_int64, _i, _i, _vp]),
    "ffcorr_tile_f32": (_i, [_vp, _vp, ctypes.c_int64, _i, _i, _vp]),
    "ffcorr_lookup_bwd_f32": (_i, [ctypes.POINTER(_vp), _i, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ffcorr_pyramid_bwd_f32": (_i, [ctypes.POINTER(_vp), _i, ctypes.c_int64, _i, _i, _vp]),
    "ffcorr_volume_bwd_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ffcorr_pwc81_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_float, _vp]),
    "ffcorr_pwc81_bwd_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ffcorr_build_tiled_f16": (_i, [_vp, _vp, ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _vp, ctypes.c_size_t, _vp]),
    "ffcorr_convc1_packed_bytes": (ctypes.c_size_t, []),
    "ffcorr_pack_convc1_weight": (_i, [_vp, _i, _i, _vp, _vp]),
    "ffcorr_lookup_convc1_tiled_f32": (_i, [ctypes.POINTER(_vp), _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ffcorr_lookup_tiled_f16": (_i, [ctypes.POINTER(_vp), _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ffcorr_untile_f16": (_i, [_vp, _vp, ctypes.c_int64, _i, _i, _vp]),
    "ffcorr_tile_f16": (_i, [_vp, _vp, ctypes.c_int64, _i, _i, _vp]),
    "ffcorr_build_grouped_f32": (_i, [_vp, _vp, ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _vp, ctypes.c_size_t, _vp]),
    "ffcorr_lookup_grouped_f32": (_i, [ctypes.POINTER(_vp), _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ffcorr_grouped_level_elems": (ctypes.c_int64, [_i, _i, _i, _i, _i]),
    "ffcorr_ungroup_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ffcorr_group_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ffcorr_backwarp_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_float, _vp]),
}
SYMBOLS = tuple(_PROTOS)


def build(verbose: bool = False) -> str:
    """Compile libffcorr.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "-j8"], stdout=out, stderr=None if verbose else subprocess.STDOUT)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: the CUDA extension is the only backend of this package "
                        f"(no CPU or PyTorch fallback). Build it with `make -C {CSRC}` or "
                        f"`python -c 'import __graft_entry__ as g; g.build()'`."
                    )
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _PROTOS.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().ffcorr_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr_array(tensors) -> ctypes.Array:
    arr = (_vp * len(tensors))()
    for k, t in enumerate(tensors):
        arr[k] = t.data_ptr()
    return arr


def current_stream(device=None) -> int:
    """Raw cudaStream_t of torch's current stream ON `device` (default: the current device)."""
    import torch

    return torch.cuda.current_stream(device).cuda_stream


class on_device:
    """``with on_device(t0, t1, ...) as stream:`` -- every launch of this package runs inside one of these.

    Checks that all tensors live on ONE CUDA device, makes that device current for the duration of the launch (the
    kernels, the TMA descriptors and cudaFuncSetAttribute all act on the current device) and yields the raw handle of
    torch's current stream on THAT device.  ATen ops guard the device the same way; without this a model on cuda:1
    driven while cuda:0 is current would launch on device 0 with device-1 pointers."""

    def __init__(self, *tensors):
        import torch

        devs = {t.device for t in tensors if t is not None}
        if len(devs) != 1:
            raise ValueError(f"all tensors of one call must live on the same device, got {sorted(map(str, devs))}")
        (self.device,) = devs
        if self.device.type != "cuda":
            raise NotImplementedError(f"tensor on {self.device}: the B200 correlation path has no CPU implementation")
        self._guard = torch.cuda.device(self.device)

    def __enter__(self) -> int:
        self._guard.__enter__()
        return current_stream(self.device)

    def __exit__(self, *exc):
        return self._guard.__exit__(*exc)
