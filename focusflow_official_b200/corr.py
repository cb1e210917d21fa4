"""CorrBlock / AlternateCorrBlock: drop-ins for ``FF_RAFT_Core/corr.py:12-91`` backed by libffcorr (sm_100a).

Same surface as the reference: ``CorrBlock(fmap1, fmap2, num_levels=4, radius=4)`` builds the
all-pairs volume and its average-pool pyramid once (reference ``corr.py:13-27``), and
``block(coords)`` returns the ``[B, num_levels*(2r+1)^2, h, w]`` fp32 window lookup
(``corr.py:29-50``).  ``num_levels``, ``radius`` and ``corr_pyramid`` are attributes as in the
reference (``corr.py:14-16``); ``corr_pyramid[i]`` is ``[B*h*w, 1, h>>i, w>>i]``.

Differences, all deliberate:
  * CUDA only.  A CPU tensor raises ``NotImplementedError`` (there is no fallback path).
  * the contraction runs on tcgen05 tensor cores with fp16-rounded operands and fp32
    accumulation by default (the reference uses TF32 cuBLAS, ``common.py:25-27``);
    ``precision="fp32" | "bf16x3" | "tf32"`` select the other operand modes.
  * one kernel per call instead of ~60 (lookup); volume + pyramid are ONE GEMM launch (plus the operand pre-pass)
    for the tensor-core precisions -- the pyramid is then stored as 4x4-pixel tiles (``layout="tiled"``) and
    ``corr_pyramid`` is converted lazily; under autograd all lookups of a block share one row-major gradient
    buffer (``_GradSink``), and the forward pyramid is still tiled because the backward never reads it.
  * range: fp16 operands need |fmap| < 65504 (and lose precision below ~6e-5); feature maps of the trained
    networks are O(1-10).  Use ``precision="tf32"`` / ``"bf16x3"`` for unbounded activations.
  * ``sampler="cuda"`` (default) reproduces the lookups of the reference run on a GPU, ``"cpu"`` those of its CPU run
    (they differ by <= 1 ulp of the normalised coordinate, see :func:`set_sampler_semantics`); per block / per call.
  * ``channels_last=True`` returns the same ``[B, C, h, w]`` values in NHWC memory (``torch.channels_last``), the
    layout the consumer ``convc1`` (``update.py:82-83,90``) runs in, written directly by the lookup kernel.
  * ``storage="fp16"`` (opt-in) stores the pyramid as IEEE half floats in the same 4x4-pixel tiles (32 bytes each):
    fp32 accumulation and pooling, one rounding per stored value (2e-4 relative), fp32 interpolation in the lookup.
    Halves the bytes of both HBM-bound kernels; needs |corr| < 65504 and 2-4 levels.  Default ``"fp32"``.
  * every launch runs under a device guard on the device of its tensors, on torch's current stream of that device.
"""
from __future__ import annotations

import os
import weakref
from typing import List, Optional

import torch

from . import _lib

__all__ = ["CorrBlock", "AlternateCorrBlock", "set_sampler_semantics", "get_sampler_semantics", "coords_grid", "correlation_volume", "correlation_pyramid", "lookup",
           "tiled_pyramid", "lookup_tiled", "tile_levels", "untile_levels"]

DEFAULT_PRECISION = "fp16"
# storage of the pyramid inside CorrBlock when no gradient is needed: "tiled" (4x4-pixel tiles, the fast
# path) or "rowmajor" (the reference layout; always used under autograd and for precision="fp32")
DEFAULT_LAYOUT = os.environ.get("FFCORR_LAYOUT", "tiled")


_default_sampler = "cuda"


def set_sampler_semantics(which: str) -> None:
    """Default ``sampler`` of blocks / calls that do not name one: ``"cuda"`` (initial value) reproduces the lookups of
    the reference run on a GPU (ATen's CUDA kernel multiplies by the fp32 reciprocal of ``W-1`` in ``utils.py:61-62``),
    ``"cpu"`` those of its CPU run (ATen divides) -- the form the golden vectors in ``tests/golden`` were made with.
    The two differ by <= 1 ulp of the normalised coordinate.  This is a Python-side default only: the library has no
    global state, the choice travels with every call (``include/ffcorr.h``: ``sampler``)."""
    global _default_sampler
    if which not in _lib.SAMPLERS:
        raise ValueError(f"sampler semantics must be 'cpu' or 'cuda', got {which!r}")
    _default_sampler = which


def get_sampler_semantics() -> str:
    return _default_sampler


def _sampler_code(sampler) -> int:
    try:
        return _lib.SAMPLERS[sampler or _default_sampler]
    except KeyError:
        raise ValueError(f"sampler must be 'cpu' or 'cuda', got {sampler!r}") from None


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise NotImplementedError(
            f"{name} is on {t.device}: the B200 correlation path has no CPU implementation"
        )


def coords_grid(batch: int, ht: int, wd: int, device) -> torch.Tensor:
    """``utils/utils.py:74-77``: [B, 2, ht, wd], channel 0 = x, channel 1 = y."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def _level_shapes(b: int, h: int, w: int, num_levels: int):
    return [(b * h * w, 1, h >> i, w >> i) for i in range(num_levels)]


# ------------------------------------------------------------------------------------
# raw (non-differentiable) launches
# ------------------------------------------------------------------------------------
def _volume_pyramid_raw(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int, precision: int) -> List[torch.Tensor]:
    b, d, h, w = fmap1.shape
    if (h >> (num_levels - 1)) < 1 or (w >> (num_levels - 1)) < 1:
        raise ValueError(f"{h}x{w} feature map is too small for {num_levels} pyramid levels")
    L = _lib.lib()
    levels = [torch.empty(s, device=fmap1.device, dtype=torch.float32) for s in _level_shapes(b, h, w, num_levels)]
    ws_bytes = L.ffcorr_volume_workspace_bytes(b, d, h, w, precision)
    ws = torch.empty(max(ws_bytes, 1), device=fmap1.device, dtype=torch.uint8)
    with _lib.on_device(fmap1, fmap2) as stream:
        _lib.check(L.ffcorr_volume_f32(fmap1.data_ptr(), fmap2.data_ptr(), levels[0].data_ptr(), b, d, h, w, precision,
                                       ws.data_ptr() if ws_bytes else None, ws_bytes, stream), "ffcorr_volume_f32")
        _lib.check(L.ffcorr_pyramid_f32(_lib.ptr_array(levels), num_levels, b * h * w, h, w, stream), "ffcorr_pyramid_f32")
    # `ws` may be freed here: the caching allocator is stream-ordered on the current stream.
    return levels


def _tiled_elems(h: int, w: int, level: int) -> int:
    return int(_lib.lib().ffcorr_tiled_map_elems(h, w, level))


def tiled_supported(h: int, w: int, num_levels: int) -> bool:
    return bool(_lib.lib().ffcorr_tiled_supported(num_levels, h, w))


STORAGES = {"fp32": torch.float32, "fp16": torch.float16}


def _storage_dtype(storage) -> torch.dtype:
    try:
        return STORAGES[storage or "fp32"]
    except KeyError:
        raise ValueError(f"storage must be 'fp32' or 'fp16', got {storage!r}") from None


def _volume_pyramid_tiled_raw(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int, precision: int,
                              fused: bool = True, storage: str = "fp32", grouped: bool = False) -> List[torch.Tensor]:
    """Volume + pyramid in the tiled layout: level i is [B*h*w, tiled_map_elems(h, w, i)], fp32 or (storage="fp16") half.

    fused=True (default): one GEMM launch writes every level (ffcorr_build_tiled_f32 / _f16); fused=False: the GEMM
    writes level 0 and the standalone pooling kernel re-reads it (bit-identical results, kept for tests; fp32 only)."""
    b, d, h, w = fmap1.shape
    L = _lib.lib()
    dtype = _storage_dtype(storage)
    if grouped:
        # "G32": [B * groups of 32 queries, tiles, 32 queries x 16 floats] -- recognised by its three dimensions
        if dtype != torch.float32 or not fused or not (2 <= num_levels <= 4):
            raise ValueError("layout='grouped' is produced by the fp32 fused build: 2 to 4 levels")
        ng = (h * w + 31) // 32
        levels = [torch.empty((b * ng, _tiled_elems(h, w, i) // 16, 512), device=fmap1.device, dtype=torch.float32)
                  for i in range(num_levels)]
        ws_bytes = L.ffcorr_volume_workspace_bytes(b, d, h, w, precision)
        ws = torch.empty(max(ws_bytes, 1), device=fmap1.device, dtype=torch.uint8)
        with _lib.on_device(fmap1, fmap2) as stream:
            _lib.check(L.ffcorr_build_grouped_f32(fmap1.data_ptr(), fmap2.data_ptr(), _lib.ptr_array(levels), num_levels, b, d, h, w,
                                                  precision, ws.data_ptr(), ws_bytes, stream), "ffcorr_build_grouped_f32")
        return levels
    levels = [torch.empty((b * h * w, _tiled_elems(h, w, i)), device=fmap1.device, dtype=dtype)
              for i in range(num_levels)]
    ws_bytes = L.ffcorr_volume_workspace_bytes(b, d, h, w, precision)
    ws = torch.empty(max(ws_bytes, 1), device=fmap1.device, dtype=torch.uint8)
    with _lib.on_device(fmap1, fmap2) as stream:
        if dtype == torch.float16:
            if not fused or not (2 <= num_levels <= 4):
                raise ValueError("storage='fp16' is produced by the fused build: 2 to 4 levels")
            _lib.check(L.ffcorr_build_tiled_f16(fmap1.data_ptr(), fmap2.data_ptr(), _lib.ptr_array(levels), num_levels, b, d, h, w,
                                                precision, ws.data_ptr(), ws_bytes, stream), "ffcorr_build_tiled_f16")
            return levels
        if fused:
            _lib.check(L.ffcorr_build_tiled_f32(fmap1.data_ptr(), fmap2.data_ptr(), _lib.ptr_array(levels), num_levels, b, d, h, w,
                                                precision, ws.data_ptr(), ws_bytes, stream), "ffcorr_build_tiled_f32")
            return levels
        _lib.check(L.ffcorr_volume_tiled_f32(fmap1.data_ptr(), fmap2.data_ptr(), levels[0].data_ptr(), b, d, h, w, precision,
                                             ws.data_ptr(), ws_bytes, stream), "ffcorr_volume_tiled_f32")
        _lib.check(L.ffcorr_pyramid_tiled_f32(_lib.ptr_array(levels), num_levels, b * h * w, h, w, stream),
                   "ffcorr_pyramid_tiled_f32")
    return levels


def _alloc_lookup_out(coords: torch.Tensor, num_levels: int, radius: int, channels_last: bool):
    """-> (storage handed to the kernel, the [B, C, h, w] tensor the caller sees).  channels_last: NHWC storage, viewed
    as NCHW with torch.channels_last strides -- what cuDNN's tensor-core kernels for the consumer conv want."""
    b, _, h, w = coords.shape
    k = 2 * radius + 1
    if channels_last:
        store = torch.empty((b, h, w, num_levels * k * k), device=coords.device, dtype=torch.float32)
        return store, store.permute(0, 3, 1, 2)
    store = torch.empty((b, num_levels * k * k, h, w), device=coords.device, dtype=torch.float32)
    return store, store


def _lookup_tiled_raw(levels, level_ptrs, coords: torch.Tensor, radius: int, sampler: int = 1,
                      channels_last: bool = False) -> torch.Tensor:
    b, _, h, w = coords.shape
    if levels[0].dtype == torch.float16:
        # the fp16-stored pyramid is read by the channels-last kernel; NCHW callers get a converted copy
        store, out = _alloc_lookup_out(coords, len(levels), radius, True)
        with _lib.on_device(coords, levels[0]) as stream:
            _lib.check(_lib.lib().ffcorr_lookup_tiled_f16(level_ptrs, len(levels), coords.data_ptr(), store.data_ptr(), b, h, w,
                                                          radius, sampler, 1, stream), "ffcorr_lookup_tiled_f16")
        return out if channels_last else out.contiguous()
    store, out = _alloc_lookup_out(coords, len(levels), radius, channels_last)
    fn = "ffcorr_lookup_grouped_f32" if levels[0].dim() == 3 else "ffcorr_lookup_tiled_f32"
    with _lib.on_device(coords, levels[0]) as stream:
        _lib.check(getattr(_lib.lib(), fn)(level_ptrs, len(levels), coords.data_ptr(), store.data_ptr(), b, h, w, radius,
                                           sampler, int(channels_last), stream), fn)
    return out


def untile_levels(tiled_levels, b: int, h: int, w: int) -> List[torch.Tensor]:
    """Tiled (fp32 / fp16) or grouped levels -> the reference's [B*h*w, 1, h>>i, w>>i] tensors."""
    out = []
    for i, t in enumerate(tiled_levels):
        hi, wi = h >> i, w >> i
        dst = torch.empty((b * h * w, 1, hi, wi), device=t.device, dtype=torch.float32)
        if t.dim() == 3:
            with _lib.on_device(t) as stream:
                _lib.check(_lib.lib().ffcorr_ungroup_f32(t.data_ptr(), dst.data_ptr(), b, h * w, hi, wi, stream), "ffcorr_ungroup_f32")
            out.append(dst)
            continue
        fn = "ffcorr_untile_f16" if t.dtype == torch.float16 else "ffcorr_untile_f32"
        with _lib.on_device(t) as stream:
            _lib.check(getattr(_lib.lib(), fn)(t.data_ptr(), dst.data_ptr(), b * h * w, hi, wi, stream), fn)
        out.append(dst)
    return out


def tile_levels(levels, storage: str = "fp32") -> List[torch.Tensor]:
    """Reference-layout levels ([Q, 1, h_i, w_i]) -> tiled storage, fp32 or fp16 (used by tests and tools)."""
    out = []
    dtype = _storage_dtype(storage)
    fn = "ffcorr_tile_f16" if dtype == torch.float16 else "ffcorr_tile_f32"
    for lv in levels:
        _require_cuda(lv, "level")
        lv = lv.float().contiguous()
        q, _, hi, wi = lv.shape
        dst = torch.empty((q, _tiled_elems(hi, wi, 0)), device=lv.device, dtype=dtype)
        with _lib.on_device(lv) as stream:
            _lib.check(getattr(_lib.lib(), fn)(lv.data_ptr(), dst.data_ptr(), q, hi, wi, stream), fn)
        out.append(dst)
    return out


def tiled_pyramid(fmap1, fmap2, num_levels: int = 4, precision=None, fused: bool = True, storage: str = "fp32") -> List[torch.Tensor]:
    """Volume + pyramid in the tiled layout (inference only)."""
    fmap1, fmap2 = _prep(fmap1, "fmap1"), _prep(fmap2, "fmap2")
    return _volume_pyramid_tiled_raw(fmap1, fmap2, num_levels, _precision_code(precision), fused, storage or "fp32")


def lookup_tiled(tiled_levels, coords: torch.Tensor, radius: int = 4, level_ptrs=None, sampler=None,
                 channels_last: bool = False) -> torch.Tensor:
    _require_cuda(coords, "coords")
    coords = coords.float().contiguous()
    return _lookup_tiled_raw(tiled_levels, level_ptrs if level_ptrs is not None else _lib.ptr_array(tiled_levels), coords, radius,
                             _sampler_code(sampler), channels_last)


def _lookup_raw(levels, level_ptrs, coords: torch.Tensor, radius: int, sampler: int = 1, channels_last: bool = False) -> torch.Tensor:
    b, _, h, w = coords.shape
    store, out = _alloc_lookup_out(coords, len(levels), radius, channels_last)
    with _lib.on_device(coords, levels[0]) as stream:
        _lib.check(_lib.lib().ffcorr_lookup_f32(level_ptrs, len(levels), coords.data_ptr(), store.data_ptr(), b, h, w, radius,
                                                sampler, int(channels_last), stream), "ffcorr_lookup_f32")
    return out


# ------------------------------------------------------------------------------------
# autograd wrappers (reference: ATen autograd of corr.py:26,45,58; coords never need a
# gradient because the caller detaches them, raft.py:216-217)
# ------------------------------------------------------------------------------------
class _GradSink:
    """Gradient of the pyramid of ONE CorrBlock, shared by all its lookups.

    The reference's autograd materialises a pyramid-sized gradient per lookup and sums them (12-32 adds of
    ~345 MB each at config 5).  Here every lookup's backward scatters straight into one zero-initialised
    buffer (``ffcorr_lookup_bwd_f32`` accumulates), and ``_VolumePyramid.backward`` -- which autograd runs after
    all lookups because they all consume its ``anchor`` output -- turns it into the fmap gradients."""

    def __init__(self, shapes, device):
        self.shapes = shapes
        self.device = device
        self.levels = None
        self.zero = None

    def buffers(self):
        if self.levels is None:
            self.levels = [torch.zeros(s, device=self.device, dtype=torch.float32) for s in self.shapes]
        return self.levels

    def take(self):
        lv, self.levels = self.levels, None     # a second backward (retain_graph) starts from zero again
        return lv

    def anchor_grad(self):
        if self.zero is None:
            self.zero = torch.zeros(1, device=self.device, dtype=torch.float32)
        return self.zero


class _VolumePyramid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap1, fmap2, num_levels, precision, sink, tiled=False, storage="fp32"):
        # tiled=True: the forward pyramid is stored as 4x4 tiles (fused build); the backward never reads it -- it
        # needs the feature maps and the ROW-MAJOR gradient pyramid the lookups scatter into the sink -- so the
        # storage order of the forward is free.  The tiled levels are not differentiable outputs themselves.
        if tiled:
            levels = _volume_pyramid_tiled_raw(fmap1, fmap2, num_levels, precision, True, storage)
        else:
            levels = _volume_pyramid_raw(fmap1, fmap2, num_levels, precision)
        ctx.save_for_backward(fmap1, fmap2)
        ctx.num_levels = num_levels
        ctx.precision = precision
        ctx.sink = sink
        ctx.set_materialize_grads(False)
        if sink is None:
            return tuple(levels)
        if tiled:
            ctx.mark_non_differentiable(*levels)
        return (*levels, torch.zeros(1, device=fmap1.device, dtype=torch.float32))   # + the anchor

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *grads):
        fmap1, fmap2 = ctx.saved_tensors
        b, d, h, w = fmap1.shape
        L = _lib.lib()
        shapes = _level_shapes(b, h, w, ctx.num_levels)
        g = ctx.sink.take() if ctx.sink is not None else None
        direct = grads[:ctx.num_levels]              # gradients of levels used directly (corr_pyramid[i] in a loss)
        if g is None:
            g = []
            for i, (gi, s) in enumerate(zip(direct, shapes)):
                if gi is None:
                    g.append(torch.zeros(s, device=fmap1.device, dtype=torch.float32))
                else:
                    gi = gi.contiguous().float()
                    g.append(gi.clone())     # updated in place below (pooling adjoint, in-place transpose)
        else:
            for gl, gi in zip(g, direct):
                if gi is not None:
                    gl.add_(gi.reshape(gl.shape))
        g1 = torch.empty_like(fmap1) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(fmap2) if ctx.needs_input_grad[1] else None
        with _lib.on_device(fmap1, fmap2, g[0]) as stream:
            _lib.check(L.ffcorr_pyramid_bwd_f32(_lib.ptr_array(g), ctx.num_levels, b * h * w, h, w, stream),
                       "ffcorr_pyramid_bwd_f32")
            _lib.check(L.ffcorr_volume_bwd_f32(g[0].data_ptr(), fmap1.data_ptr(), fmap2.data_ptr(),
                                               g1.data_ptr() if g1 is not None else None,
                                               g2.data_ptr() if g2 is not None else None, b, d, h, w, ctx.precision, stream),
                       "ffcorr_volume_bwd_f32")
        return g1, g2, None, None, None, None, None


class _UntileWithGrad(torch.autograd.Function):
    """``corr_pyramid`` of a block that stores its forward pyramid tiled while gradients are tracked: row-major copies
    whose gradients (a level used directly in a loss) go to the block's sink like those of the lookups."""

    @staticmethod
    def forward(ctx, anchor, sink, b, h, w, *tiled):
        ctx.sink = sink
        ctx.set_materialize_grads(False)
        return tuple(untile_levels(tiled, b, h, w))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *grads):
        bufs = ctx.sink.buffers()
        for buf, g in zip(bufs, grads):
            if g is not None:
                buf.add_(g.reshape(buf.shape))
        return (ctx.sink.anchor_grad(), None, None, None, None, *([None] * len(bufs)))


class _Lookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords, radius, sink, anchor, sampler, channels_last, *levels):
        tiled = levels[0].dim() == 2                     # [B*N, map_elems] tiles vs [B*N, 1, h, w] rows
        raw = _lookup_tiled_raw if tiled else _lookup_raw
        out = raw(levels, _lib.ptr_array(levels), coords, radius, sampler, channels_last)
        ctx.save_for_backward(coords)
        ctx.radius = radius
        ctx.sampler = sampler
        ctx.sink = sink
        ctx.shapes = [tuple(l.shape) for l in levels]
        if tiled and sink is None:
            raise RuntimeError("a tiled pyramid is differentiable only through its CorrBlock")
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        (coords,) = ctx.saved_tensors
        b, _, h, w = coords.shape
        gout = gout.contiguous().float()                 # NCHW, whatever layout the forward's output had
        shared = ctx.sink is not None
        glv = ctx.sink.buffers() if shared else [torch.zeros(s, device=coords.device, dtype=torch.float32) for s in ctx.shapes]
        with _lib.on_device(coords, gout, glv[0]) as stream:
            _lib.check(_lib.lib().ffcorr_lookup_bwd_f32(_lib.ptr_array(glv), len(glv), coords.data_ptr(), gout.data_ptr(),
                                                        b, h, w, ctx.radius, ctx.sampler, stream), "ffcorr_lookup_bwd_f32")
        if shared:   # the gradient reaches _VolumePyramid through the sink; the anchor only orders the two
            return (None, None, None, ctx.sink.anchor_grad(), None, None, *([None] * len(glv)))
        return (None, None, None, None, None, None, *glv)


def _prep(fmap: torch.Tensor, name: str) -> torch.Tensor:
    _require_cuda(fmap, name)
    if fmap.dim() != 4:
        raise ValueError(f"{name} must be [B, D, h, w], got {tuple(fmap.shape)}")
    return fmap.float().contiguous()


def _precision_code(precision) -> int:
    if isinstance(precision, int):
        return precision
    try:
        return _lib.PRECISIONS[precision or DEFAULT_PRECISION]
    except KeyError:
        raise ValueError(f"unknown precision {precision!r}; one of {sorted(_lib.PRECISIONS)}") from None


def correlation_pyramid(fmap1, fmap2, num_levels: int = 4, precision=None) -> List[torch.Tensor]:
    """Volume + pyramid (``corr.py:18-27``), differentiable w.r.t. both feature maps."""
    fmap1, fmap2 = _prep(fmap1, "fmap1"), _prep(fmap2, "fmap2")
    if fmap1.shape != fmap2.shape:
        raise ValueError(f"fmap shapes differ: {tuple(fmap1.shape)} vs {tuple(fmap2.shape)}")
    code = _precision_code(precision)
    if torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad):
        return list(_VolumePyramid.apply(fmap1, fmap2, num_levels, code, None))
    return _volume_pyramid_raw(fmap1, fmap2, num_levels, code)


def correlation_volume(fmap1, fmap2, precision=None) -> torch.Tensor:
    """``CorrBlock.corr`` (``corr.py:52-60``): [B, h, w, 1, h, w]."""
    b, _, h, w = fmap1.shape
    return correlation_pyramid(fmap1, fmap2, 1, precision)[0].view(b, h, w, 1, h, w)


def lookup(levels, coords: torch.Tensor, radius: int = 4, level_ptrs=None, sampler=None, channels_last: bool = False) -> torch.Tensor:
    _require_cuda(coords, "coords")
    coords = coords.float().contiguous()
    code = _sampler_code(sampler)
    if torch.is_grad_enabled() and any(l.requires_grad for l in levels):
        return _Lookup.apply(coords, radius, None, None, code, channels_last, *levels)
    return _lookup_raw(levels, level_ptrs if level_ptrs is not None else _lib.ptr_array(levels), coords, radius, code, channels_last)


class CorrBlock:
    def __init__(self, fmap1, fmap2, num_levels: int = 4, radius: int = 4, precision: Optional[str] = None,
                 layout: Optional[str] = None, sampler: Optional[str] = None, channels_last: bool = False,
                 storage: Optional[str] = None):
        self.num_levels = num_levels
        self.radius = radius
        self._sched, self._sched_used = None, 0
        self.storage = storage or "fp32"
        _storage_dtype(self.storage)
        self.sampler = sampler or _default_sampler
        self._sampler = _sampler_code(sampler)
        self.channels_last = bool(channels_last)
        if fmap1.device != fmap2.device:
            raise ValueError(f"fmap1 is on {fmap1.device}, fmap2 on {fmap2.device}")
        b, _, h, w = fmap1.shape
        self._shape = (b, h, w)
        self._sink = None
        self._grad_tiled = False
        layout = layout or DEFAULT_LAYOUT
        if layout not in ("tiled", "rowmajor", "grouped"):
            raise ValueError(f"layout must be 'tiled', 'grouped' or 'rowmajor', got {layout!r}")
        self._grouped = layout == "grouped"
        if self._grouped:
            layout = "tiled"        # same kernels, same values; only the storage order of the tiles differs
        needs_grad = torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad)
        code = _precision_code(precision)
        if (h >> (num_levels - 1)) < 1 or (w >> (num_levels - 1)) < 1:
            raise ValueError(f"{h}x{w} feature map is too small for {num_levels} pyramid levels")
        self._tiled = (layout == "tiled" and not needs_grad and code != _lib.PREC_FP32 and b > 0
                       and fmap1.is_cuda and tiled_supported(h, w, num_levels))
        if self._tiled:
            f1, f2 = _prep(fmap1, "fmap1"), _prep(fmap2, "fmap2")
            if f1.shape != f2.shape:
                raise ValueError(f"fmap shapes differ: {tuple(f1.shape)} vs {tuple(f2.shape)}")
            self._levels = _volume_pyramid_tiled_raw(f1, f2, num_levels, code, True, self.storage,
                                                     self._grouped and 2 <= num_levels <= 4 and self.storage == "fp32")
            self._rowmajor = None                      # materialised on first access of .corr_pyramid
        elif needs_grad:
            f1, f2 = _prep(fmap1, "fmap1"), _prep(fmap2, "fmap2")
            if f1.shape != f2.shape:
                raise ValueError(f"fmap shapes differ: {tuple(f1.shape)} vs {tuple(f2.shape)}")
            self._sink = _GradSink(_level_shapes(b, h, w, num_levels), f1.device)
            # forward pyramid tiled (fused build, faster lookups) whenever the tensor-core path is allowed: the
            # backward only needs the feature maps and the row-major gradient pyramid in the sink
            self._grad_tiled = (layout == "tiled" and code != _lib.PREC_FP32 and b > 0 and tiled_supported(h, w, num_levels))
            outs = _VolumePyramid.apply(f1, f2, num_levels, code, self._sink, self._grad_tiled,
                                        self.storage if self._grad_tiled else "fp32")
            self._levels, self._anchor = list(outs[:-1]), outs[-1]
            self._rowmajor = None if self._grad_tiled else self._levels
        else:
            self._levels = correlation_pyramid(fmap1, fmap2, num_levels, precision)
            self._rowmajor = self._levels
        if self.storage == "fp16" and self._levels[0].dtype != torch.float16:
            raise ValueError("storage='fp16' needs the tiled layout, a tensor-core precision and 2-4 levels that fit it "
                             f"(layout={layout!r}, precision={precision!r}, {num_levels} levels on a {h}x{w} map)")
        self._ptrs = _lib.ptr_array(self._levels)

    @property
    def corr_pyramid(self) -> List[torch.Tensor]:
        """The reference's ``[B*h*w, 1, h>>i, w>>i]`` list (``corr.py:16,23-27``); converted lazily when the
        block stores its pyramid in the tiled layout."""
        if self._rowmajor is None:
            b, h, w = self._shape
            if self._grad_tiled and torch.is_grad_enabled():
                self._rowmajor = list(_UntileWithGrad.apply(self._anchor, self._sink, b, h, w, *self._levels))
            else:
                self._rowmajor = untile_levels(self._levels, b, h, w)
        return self._rowmajor

    def __call__(self, coords: torch.Tensor) -> torch.Tensor:
        b, two, h, w = coords.shape
        if (b, h, w) != self._shape or two != 2:
            raise ValueError(f"coords {tuple(coords.shape)} does not match the volume built for B,h,w={self._shape}")
        if self._tiled:
            return lookup_tiled(self._levels, coords, self.radius, self._ptrs, self.sampler, self.channels_last)
        if self._sink is not None and torch.is_grad_enabled():
            _require_cuda(coords, "coords")
            return _Lookup.apply(coords.float().contiguous(), self.radius, self._sink, self._anchor, self._sampler,
                                 self.channels_last, *self._levels)
        if self._grad_tiled:
            return lookup_tiled(self._levels, coords, self.radius, self._ptrs, self.sampler, self.channels_last)
        return lookup(self._levels, coords, self.radius, self._ptrs, self.sampler, self.channels_last)

    def supports_lookup_conv(self, conv) -> bool:
        """Whether :meth:`lookup_conv` can replace ``relu(conv(self(coords)))`` for this block and this 1x1 convolution."""
        w = getattr(conv, "weight", None)
        return bool(self._tiled and not self._grouped and self._levels[0].dtype == torch.float32 and self.num_levels == 4
                    and self.radius == 4 and w is not None and tuple(w.shape) == (256, 324, 1, 1) and conv.bias is not None
                    and w.device == self._levels[0].device and w.dtype == torch.float32 and not torch.is_grad_enabled())

    def lookup_conv(self, coords: torch.Tensor, conv) -> torch.Tensor:
        """``relu(conv(self(coords)))`` for ``conv = BasicMotionEncoder.convc1`` (``update.py:82-83,90``) in ONE launch:
        the 324 samples per query go straight into the tensor cores and never reach global memory.  Returns
        ``[B, 256, h, w]`` fp32 in channels_last storage.  Inference only; fp16 operands, fp32 accumulate."""
        b, two, h, w = coords.shape
        if (b, h, w) != self._shape or two != 2:
            raise ValueError(f"coords {tuple(coords.shape)} does not match the volume built for B,h,w={self._shape}")
        if not self.supports_lookup_conv(conv):
            raise ValueError("lookup_conv needs the fp32 tiled pyramid (4 levels, radius 4), a Conv2d(324, 256, 1) with bias "
                             "on the same device, and no autograd")
        _require_cuda(coords, "coords")
        coords = coords.float().contiguous()
        packed = _packed_convc1(conv)
        bias = conv.bias.detach().contiguous()
        store = torch.empty((b, h, w, 256), device=coords.device, dtype=torch.float32)
        # the kernel's tile scheduler wants a zeroed int32 per launch: one memset buys 64 of them
        if self._sched is None or self._sched_used == self._sched.numel():
            self._sched = torch.zeros(64, device=coords.device, dtype=torch.int32)
            self._sched_used = 0
        counter = self._sched[self._sched_used:self._sched_used + 1]
        self._sched_used += 1
        with _lib.on_device(coords, self._levels[0], packed, bias) as stream:
            _lib.check(_lib.lib().ffcorr_lookup_convc1_tiled_f32(self._ptrs, 4, coords.data_ptr(), packed.data_ptr(), bias.data_ptr(),
                                                                 store.data_ptr(), counter.data_ptr(), b, h, w, 4, self._sampler,
                                                                 stream), "ffcorr_lookup_convc1_tiled_f32")
        return store.permute(0, 3, 1, 2)

    @staticmethod
    def corr(fmap1, fmap2, precision: Optional[str] = None):
        return correlation_volume(fmap1, fmap2, precision)


_packed_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _packed_convc1(conv) -> torch.Tensor:
    """convc1.weight in the fused kernel's operand order (fp16), re-packed only when the weight tensor changes."""
    w = conv.weight
    hit = _packed_cache.get(conv)
    if hit is not None and hit[0] == (w.data_ptr(), w._version, w.device):
        return hit[1]
    L = _lib.lib()
    packed = torch.empty(L.ffcorr_convc1_packed_bytes(), device=w.device, dtype=torch.uint8)
    w2 = w.detach().reshape(w.shape[0], w.shape[1]).contiguous()
    with _lib.on_device(w2) as stream:
        _lib.check(L.ffcorr_pack_convc1_weight(w2.data_ptr(), w2.shape[0], w2.shape[1], packed.data_ptr(), stream),
                   "ffcorr_pack_convc1_weight")
    _packed_cache[conv] = ((w.data_ptr(), w._version, w.device), packed)
    return packed


class AlternateCorrBlock:
    """Memory-bounded drop-in for the reference's ``AlternateCorrBlock`` (``corr.py:63-91``): same constructor, same
    ``__call__(coords)``, same values as :class:`CorrBlock` (bit-identical), but the pyramid is never stored -- every
    lookup recomputes it for ``chunk`` queries at a time with the fused tensor-core build and throws it away, so
    memory is O(chunk * h*w) instead of O((h*w)^2).  Inference only (the reference's version has no backward
    either: ``alt_cuda_corr`` is forward-only in every shipped config)."""

    def __init__(self, fmap1, fmap2, num_levels: int = 4, radius: int = 4, precision: Optional[str] = None,
                 chunk: Optional[int] = None, max_pyramid_bytes: int = 512 << 20, sampler: Optional[str] = None,
                 channels_last: bool = False):
        self.num_levels = num_levels
        self.radius = radius
        self.sampler = sampler or _default_sampler
        self._sampler = _sampler_code(sampler)
        self.channels_last = bool(channels_last)
        f1, f2 = _prep(fmap1, "fmap1").detach(), _prep(fmap2, "fmap2").detach()
        if f1.shape != f2.shape:
            raise ValueError(f"fmap shapes differ: {tuple(f1.shape)} vs {tuple(f2.shape)}")
        b, d, h, w = f1.shape
        self._shape = (b, d, h, w)
        self._code = _precision_code(precision)
        if self._code == _lib.PREC_FP32 or not tiled_supported(h, w, num_levels):
            raise ValueError("AlternateCorrBlock needs a tensor-core precision (fp16 / tf32 / bf16x3) and a shape the "
                             "tiled kernels support")
        n = h * w
        per_query = 4 * sum(_tiled_elems(h, w, i) for i in range(num_levels))
        if chunk is None:   # as many queries as fit the budget, in whole GEMM row tiles
            chunk = max(128, (max_pyramid_bytes // max(1, b * per_query)) // 128 * 128)
        # whole 32-query lookup units: the lookup picks its separable / exact-per-tap path per unit, so only then is
        # every query evaluated exactly like in CorrBlock (the two paths differ by ~1e-7 relative)
        self.chunk = int(min(max(32, chunk // 32 * 32), (n + 31) // 32 * 32))
        L = _lib.lib()
        self._ws_bytes = L.ffcorr_volume_workspace_bytes(b, d, h, w, self._code)
        self._ws = torch.empty(max(self._ws_bytes, 1), device=f1.device, dtype=torch.uint8)
        with _lib.on_device(f1, f2) as stream:
            _lib.check(L.ffcorr_stage_operands_f32(f1.data_ptr(), f2.data_ptr(), num_levels, b, d, h, w, self._code,
                                                   self._ws.data_ptr(), self._ws_bytes, stream), "ffcorr_stage_operands_f32")
        self._levels = [torch.empty((b * self.chunk, _tiled_elems(h, w, i)), device=f1.device, dtype=torch.float32)
                        for i in range(num_levels)]
        self._ptrs = _lib.ptr_array(self._levels)

    def __call__(self, coords: torch.Tensor) -> torch.Tensor:
        b, d, h, w = self._shape
        _require_cuda(coords, "coords")
        if tuple(coords.shape) != (b, 2, h, w):
            raise ValueError(f"coords {tuple(coords.shape)} does not match B,h,w={(b, h, w)}")
        coords = coords.detach().float().contiguous()
        store, out = _alloc_lookup_out(coords, self.num_levels, self.radius, self.channels_last)
        L = _lib.lib()
        n = h * w
        with _lib.on_device(coords, self._ws) as stream:
            for q0 in range(0, n, self.chunk):
                nq = min(self.chunk, n - q0)            # q0 stays a multiple of 32
                _lib.check(L.ffcorr_build_tiled_chunk_f32(self._ptrs, self.num_levels, b, d, h, w, q0, nq, self._code,
                                                          self._ws.data_ptr(), self._ws_bytes, stream),
                           "ffcorr_build_tiled_chunk_f32")
                _lib.check(L.ffcorr_lookup_tiled_chunk_f32(self._ptrs, self.num_levels, coords.data_ptr(), store.data_ptr(),
                                                           b, h, w, q0, nq, self.radius, self._sampler, int(self.channels_last),
                                                           stream), "ffcorr_lookup_tiled_chunk_f32")
        return out
