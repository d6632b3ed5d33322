"""Host-side mirror of the hot-path surface named in BASELINE.json's north_star:

    segment(skel[B,T,V,C]) -> logits[B,T,K]
    align(a, b)            -> (cost, path)

The reference has no operator / plugin interface to mirror ("not
in reference"; stages named at /root/reference/README.md:17-22, 27-34, 44-52),
so these two calls ARE the interface.  Everything here is plumbing: ctypes over
the C ABI of include/golfer_b200.h, with PyTorch used only for device memory and
streams.  There is no CPU fallback: if the CUDA library is missing or no sm_100
device is present these calls raise `GolferError`.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Dict, Optional, Tuple

import numpy as np

from .config import V0, GolfSegConfig
from .params import make_params, pack_blob

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libgolfer_b200.so"
GS_MAX_BLOCKS = 8
GS_MAX_BRANCHES = 8
PRECISIONS = {"fp32": 0, "bf16": 1}

# every symbol include/golfer_b200.h declares (tests check the .so exports all of them)
ABI_SYMBOLS = (
    "gs_abi_version", "gs_last_error", "gs_create", "gs_destroy", "gs_segment",
    "gs_segment_host", "gs_segment_host_submit", "gs_segment_host_wait", "gs_segment_features", "gs_align", "gs_align_host", "gs_align_host_submit", "gs_align_host_wait", "gs_pair_cost",
    "gs_compare", "gs_normalize_pose", "gs_align_phase", "gs_set_align_encoder", "gs_align_embed", "gs_launch_count", "gs_workspace_bytes", "gs_last_kernel_ms",
    "gs_profile_enable", "gs_profile_reset", "gs_profile_kernels", "gs_profile_read", "gs_profile_read_block", "gs_debug_read",
)


class GolferError(RuntimeError):
    """Raised when the CUDA library is missing or a C-ABI call fails."""


class _GsConfig(ctypes.Structure):
    _fields_ = [
        ("num_joints", ctypes.c_int32), ("in_channels", ctypes.c_int32),
        ("num_partitions", ctypes.c_int32), ("num_blocks", ctypes.c_int32),
        ("widths", ctypes.c_int32 * GS_MAX_BLOCKS),
        ("num_branches", ctypes.c_int32), ("kernel_size", ctypes.c_int32),
        ("dilations", ctypes.c_int32 * GS_MAX_BRANCHES),
        ("se_reduction", ctypes.c_int32), ("stj_reduction", ctypes.c_int32),
        ("num_classes", ctypes.c_int32), ("precision", ctypes.c_int32),
    ]


def library_path() -> str:
    return os.path.join(_HERE, "lib", _LIB_NAME)


_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen libgolfer_b200.so and declare the prototypes.  Fails loudly if the
    extension has not been built (`python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            raise GolferError(
                f"{path} is missing: the CUDA extension is not built and there is no CPU "
                "fallback. Run __graft_entry__.build().")
        try:
            L = ctypes.CDLL(path)
        except OSError as e:  # pragma: no cover
            raise GolferError(f"cannot load {path}: {e}") from e
        vp, i32, u8p = ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p
        L.gs_abi_version.restype = ctypes.c_int
        L.gs_last_error.restype = ctypes.c_char_p
        L.gs_create.argtypes = [ctypes.POINTER(vp), i32, ctypes.POINTER(_GsConfig), vp,
                                ctypes.c_size_t, i32, i32]
        L.gs_destroy.argtypes = [vp]
        L.gs_segment.argtypes = [vp, vp, vp, u8p, i32, i32, vp]
        L.gs_segment_host.argtypes = [vp, vp, vp, u8p, i32, i32]
        L.gs_segment_host_submit.argtypes = [vp, vp, vp, u8p, i32, i32, ctypes.POINTER(ctypes.c_int)]
        L.gs_segment_host_wait.argtypes = [vp, i32]
        L.gs_segment_features.argtypes = [vp, vp, i32, vp, i32, i32, vp]
        L.gs_align.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]
        L.gs_align_host.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
        L.gs_align_host_submit.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, ctypes.POINTER(ctypes.c_int)]
        L.gs_align_host_wait.argtypes = [vp, i32]
        L.gs_pair_cost.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
        L.gs_compare.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
        L.gs_normalize_pose.argtypes = [vp, vp, vp, i32, i32, i32, ctypes.c_float, vp]
        L.gs_align_phase.argtypes = [vp, vp, vp, vp, vp, ctypes.c_float, i32, i32, i32, i32, i32, vp, vp, vp, vp]
        L.gs_set_align_encoder.argtypes = [vp, vp, ctypes.c_size_t]
        L.gs_align_embed.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]
        L.gs_launch_count.argtypes = [vp]
        L.gs_launch_count.restype = ctypes.c_int64
        L.gs_workspace_bytes.argtypes = [vp]
        L.gs_workspace_bytes.restype = ctypes.c_size_t
        L.gs_last_kernel_ms.argtypes = [vp]
        L.gs_last_kernel_ms.restype = ctypes.c_float
        L.gs_debug_read.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_size_t]
        L.gs_debug_read.restype = ctypes.c_int
        L.gs_profile_enable.argtypes = [vp, i32]
        L.gs_profile_reset.argtypes = [vp]
        L.gs_profile_kernels.restype = ctypes.c_int
        L.gs_profile_read.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_char_p),
                                      ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64),
                                      ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.gs_profile_read_block.argtypes = [vp, i32, i32, ctypes.POINTER(ctypes.c_double),
                                            ctypes.POINTER(ctypes.c_int64)]
        for name in ("gs_profile_enable", "gs_profile_reset", "gs_profile_read", "gs_profile_read_block"):
            getattr(L, name).restype = ctypes.c_int
        for name in ("gs_create", "gs_destroy", "gs_segment", "gs_segment_host", "gs_segment_host_submit",
                     "gs_segment_host_wait", "gs_align_host_submit", "gs_align_host_wait",
                     "gs_segment_features", "gs_align", "gs_align_host", "gs_pair_cost",
                     "gs_compare", "gs_normalize_pose", "gs_align_phase", "gs_set_align_encoder", "gs_align_embed"):
            getattr(L, name).restype = ctypes.c_int
        if L.gs_abi_version() != 1:
            raise GolferError("libgolfer_b200.so ABI version mismatch")
        _lib = L
        return L


def _check(rc: int, what: str):
    if rc != 0:
        msg = load_library().gs_last_error()
        raise GolferError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")


def _torch():
    import torch
    return torch


def _cfg_struct(cfg: GolfSegConfig, precision: str) -> _GsConfig:
    if precision not in PRECISIONS:
        raise GolferError(f"precision must be one of {sorted(PRECISIONS)}")
    s = _GsConfig()
    s.num_joints, s.in_channels = cfg.num_joints, cfg.in_channels
    s.num_partitions, s.num_blocks = cfg.num_partitions, cfg.num_blocks
    for i, w in enumerate(cfg.widths):
        s.widths[i] = w
    s.num_branches, s.kernel_size = cfg.num_branches, cfg.kernel_size
    for i, d in enumerate(cfg.dilations):
        s.dilations[i] = d
    s.se_reduction, s.stj_reduction = cfg.se_reduction, cfg.stj_reduction
    s.num_classes, s.precision = cfg.num_classes, PRECISIONS[precision]
    return s


def _stream_ptr(torch) -> int:
    return int(torch.cuda.current_stream().cuda_stream)


class Context:
    """Owns one gs_ctx (weights + workspace on one device)."""

    def __init__(self, device: int = 0, cfg: Optional[GolfSegConfig] = None,
                 blob: Optional[np.ndarray] = None, precision: str = "bf16",
                 max_B: int = 0, max_T: int = 0):
        self._L = load_library()
        self._h = ctypes.c_void_p()
        self.device = int(device)
        if cfg is None:
            rc = self._L.gs_create(ctypes.byref(self._h), self.device, None, None, 0, 0, 0)
        else:
            blob = np.ascontiguousarray(blob, dtype=np.float32)
            cs = _cfg_struct(cfg, precision)
            rc = self._L.gs_create(ctypes.byref(self._h), self.device, ctypes.byref(cs),
                                   blob.ctypes.data, blob.nbytes, int(max_B), int(max_T))
        _check(rc, "gs_create")

    @property
    def handle(self):
        return self._h

    def launch_count(self) -> int:
        return int(self._L.gs_launch_count(self._h))

    def workspace_bytes(self) -> int:
        return int(self._L.gs_workspace_bytes(self._h))

    def last_kernel_ms(self) -> float:
        return float(self._L.gs_last_kernel_ms(self._h))

    def debug_read(self, name: str, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        _check(self._L.gs_debug_read(self._h, name.encode(), out.ctypes.data, nbytes), "gs_debug_read")
        return out

    def profile(self, on: bool):
        _check(self._L.gs_profile_enable(self._h, 1 if on else 0), "gs_profile_enable")

    def profile_reset(self):
        _check(self._L.gs_profile_reset(self._h), "gs_profile_reset")

    def profile_read(self):
        """{kernel name: dict(ms, launches, flops, bytes)} for kernels launched while profiling."""
        out = {}
        for k in range(self._L.gs_profile_kernels()):
            name = ctypes.c_char_p()
            ms, fl, by = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
            n = ctypes.c_int64()
            _check(self._L.gs_profile_read(self._h, k, ctypes.byref(name), ctypes.byref(ms), ctypes.byref(n),
                                           ctypes.byref(fl), ctypes.byref(by)), "gs_profile_read")
            if n.value:
                per_block = {}
                for blk in range(GS_MAX_BLOCKS + 1):
                    bms, bn = ctypes.c_double(), ctypes.c_int64()
                    _check(self._L.gs_profile_read_block(self._h, k, blk, ctypes.byref(bms), ctypes.byref(bn)),
                           "gs_profile_read_block")
                    if bn.value:
                        per_block[blk] = dict(ms=bms.value, launches=bn.value)
                out[name.value.decode()] = dict(ms=ms.value, launches=n.value, flops=fl.value, bytes=by.value,
                                                blocks=per_block)
        return out

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.gs_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class Ticket:
    """One batch in flight behind Segmenter.submit: keeps the host buffers alive until it has been waited for."""
    __slots__ = ("id", "skel", "logits", "labels")

    def __init__(self, tid, skel, logits, labels):
        self.id, self.skel, self.logits, self.labels = tid, skel, logits, labels


class Segmenter:
    """segment(skel[B,T,V,C]) -> logits[B,T,K] on one B200.

    CUDA tensor in -> CUDA tensor out (asynchronous on the current stream).
    CPU tensor / ndarray in -> same kind out, through gs_segment_host (copies inside).
    """

    def __init__(self, cfg: GolfSegConfig = V0, params: Optional[Dict[str, np.ndarray]] = None,
                 seed: int = 1234, precision: str = "bf16", device: int = 0,
                 max_B: int = 256, max_T: int = 300):
        self.cfg, self.precision = cfg, precision
        self.params = params if params is not None else make_params(cfg, seed)
        self.blob = pack_blob(cfg, self.params)
        self.max_B, self.max_T = int(max_B), int(max_T)
        self.ctx = Context(device, cfg, self.blob, precision, max_B, max_T)

    # -- helpers -----------------------------------------------------------
    def _shape(self, shape) -> Tuple[int, int]:
        if len(shape) != 4 or shape[2] != self.cfg.num_joints or shape[3] != self.cfg.in_channels:
            raise GolferError(
                f"skel must be [B,T,{self.cfg.num_joints},{self.cfg.in_channels}], got {tuple(shape)}")
        B, T = int(shape[0]), int(shape[1])
        if B > self.max_B or T > self.max_T:
            raise GolferError(f"B={B},T={T} exceeds the context's max_B={self.max_B},max_T={self.max_T}")
        return B, T

    def segment(self, skel, return_labels: bool = False, out=None):
        """skel [B,T,V,C] -> logits [B,T,K] fp32 (and labels [B,T] u8).  Device tensors run on the current
        stream and return device tensors; host tensors / arrays go through the chunked, copy-overlapped host
        entry point and return when the result is in host memory.  `out` (host path only): a pinned fp32
        [B,T,K] tensor to receive the logits instead of a fresh pinned allocation per call."""
        torch = _torch()
        L = self.ctx._L
        K = self.cfg.num_classes
        if isinstance(skel, torch.Tensor) and skel.is_cuda:
            B, T = self._shape(skel.shape)
            x = skel.contiguous().float()
            logits = torch.empty((B, T, K), dtype=torch.float32, device=x.device)
            labels = torch.empty((B, T), dtype=torch.uint8, device=x.device) if return_labels else None
            if B and T:
                _check(L.gs_segment(self.ctx.handle, x.data_ptr(), logits.data_ptr(),
                                    labels.data_ptr() if return_labels else None, B, T,
                                    _stream_ptr(torch)), "gs_segment")
            return (logits, labels) if return_labels else logits
        was_numpy = not isinstance(skel, torch.Tensor)
        x = torch.as_tensor(np.asarray(skel) if was_numpy else skel, dtype=torch.float32).contiguous()
        B, T = self._shape(x.shape)
        if out is not None:
            if (not isinstance(out, torch.Tensor) or out.is_cuda or out.dtype != torch.float32
                    or tuple(out.shape) != (B, T, K) or not out.is_contiguous()):
                raise GolferError(f"out must be a contiguous host fp32 tensor of shape {(B, T, K)}")
            logits = out
        else:
            logits = torch.empty((B, T, K), dtype=torch.float32, pin_memory=True)
        labels = torch.empty((B, T), dtype=torch.uint8, pin_memory=True) if return_labels else None
        if B and T:
            _check(L.gs_segment_host(self.ctx.handle, x.data_ptr(), logits.data_ptr(),
                                     labels.data_ptr() if return_labels else None, B, T),
                   "gs_segment_host")
        if was_numpy:
            logits = logits.numpy()
            labels = labels.numpy() if return_labels else None
        return (logits, labels) if return_labels else logits

    __call__ = segment

    # -- pipelined host entry point ----------------------------------------
    def submit(self, skel, logits_out=None, labels_out=None) -> "Ticket":
        """Enqueue one HOST batch and return at once (gs_segment_host_submit): the input copy of this batch overlaps
        the kernels of the previous one, its result copy overlaps the kernels of the next one; at most two batches
        are in flight (a third submit first waits for the oldest).  `skel` must be a pinned fp32 host tensor
        [B,T,V,C] that stays untouched until `wait`; `logits_out` / `labels_out` are pinned host tensors to fill
        (allocated when omitted).  Returns a ticket for `wait`."""
        torch = _torch()
        if not isinstance(skel, torch.Tensor) or skel.is_cuda or skel.dtype != torch.float32 or not skel.is_contiguous():
            raise GolferError("submit: skel must be a contiguous fp32 host tensor (pinned for overlap)")
        B, T = self._shape(skel.shape)
        K = self.cfg.num_classes
        if logits_out is None:
            logits_out = torch.empty((B, T, K), dtype=torch.float32, pin_memory=True)
        if (logits_out.is_cuda or logits_out.dtype != torch.float32 or tuple(logits_out.shape) != (B, T, K)
                or not logits_out.is_contiguous()):
            raise GolferError(f"submit: logits_out must be a contiguous host fp32 tensor of shape {(B, T, K)}")
        if labels_out is not None and (labels_out.is_cuda or labels_out.dtype != torch.uint8
                                       or tuple(labels_out.shape) != (B, T) or not labels_out.is_contiguous()):
            raise GolferError(f"submit: labels_out must be a contiguous host u8 tensor of shape {(B, T)}")
        tk = ctypes.c_int(-1)
        _check(self.ctx._L.gs_segment_host_submit(self.ctx.handle, skel.data_ptr(), logits_out.data_ptr(),
                                                  labels_out.data_ptr() if labels_out is not None else None, B, T,
                                                  ctypes.byref(tk)), "gs_segment_host_submit")
        return Ticket(tk.value, skel, logits_out, labels_out)

    def wait(self, ticket: "Ticket"):
        """Block until the batch behind `ticket` is complete; returns (logits, labels) host tensors."""
        _check(self.ctx._L.gs_segment_host_wait(self.ctx.handle, ticket.id), "gs_segment_host_wait")
        return ticket.logits, ticket.labels

    def segment_stream(self, batches, return_labels: bool = False):
        """Generator over an iterable of pinned host batches: yields each batch's logits (and labels) in order while
        the next batch is already copying and computing."""
        torch = _torch()
        prev = None
        for x in batches:
            B, T = self._shape(x.shape)
            lab = torch.empty((B, T), dtype=torch.uint8, pin_memory=True) if return_labels else None
            tk = self.submit(x, None, lab)
            if prev is not None:
                lo, la = self.wait(prev)
                yield (lo, la) if return_labels else lo
            prev = tk
        if prev is not None:
            lo, la = self.wait(prev)
            yield (lo, la) if return_labels else lo

    def features(self, skel, block: int):
        """Block output after both attention gates, fp32 [B,T,V,C] (parity hook)."""
        torch = _torch()
        x = skel.contiguous().float()
        B, T = self._shape(x.shape)
        C = self.cfg.widths[block]
        out = torch.empty((B, T, self.cfg.num_joints, C), dtype=torch.float32, device=x.device)
        _check(self.ctx._L.gs_segment_features(self.ctx.handle, x.data_ptr(), int(block),
                                               out.data_ptr(), B, T, _stream_ptr(torch)),
               "gs_segment_features")
        return out


_default_segmenters: Dict[tuple, Segmenter] = {}
_default_align_ctx: Dict[int, Context] = {}


def _current_device(torch, t=None) -> int:
    if t is not None and isinstance(t, torch.Tensor) and t.is_cuda:
        return t.device.index
    if not torch.cuda.is_available():
        raise GolferError("no CUDA device: golfer_b200 has no CPU fallback")
    return torch.cuda.current_device()


def _dev_tensor(torch, t, name: str, dev: int, dtype):
    """`t` as a contiguous tensor of `dtype` on cuda:`dev`; anything else raises instead of handing a kernel a
    host or foreign-device pointer."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise GolferError(f"{name} must be a CUDA tensor")
    if t.device.index != dev:
        raise GolferError(f"{name} is on cuda:{t.device.index}, expected cuda:{dev} (all tensors of one call share a device)")
    return t.to(dtype).contiguous()


def segment(skel, precision: str = "bf16", cfg: GolfSegConfig = V0, seed: int = 1234):
    """Module-level convenience: seeded random-init weights (the reference ships none)."""
    torch = _torch()
    dev = _current_device(torch, skel)
    B, T = int(skel.shape[0]), int(skel.shape[1])
    key = (dev, precision, cfg.config_hash(), seed)
    seg = _default_segmenters.get(key)
    if seg is None or seg.max_B < B or seg.max_T < T:
        # grow-only: alternating shapes must not rebuild the workspace on every call
        nB, nT = max(B, 1), max(T, 1)
        if seg is not None:
            nB, nT = max(nB, seg.max_B), max(nT, seg.max_T)
            seg.ctx.close()
        seg = Segmenter(cfg, None, seed, precision, dev, nB, nT)
        _default_segmenters[key] = seg
    return seg.segment(skel)


def _align_ctx(dev: int) -> Context:
    ctx = _default_align_ctx.get(dev)
    if ctx is None:
        ctx = Context(dev)
        _default_align_ctx[dev] = ctx
    return ctx


def align_batch(a, b, ctx: Optional[Context] = None, want_path: bool = True):
    """a [N,Ta,V,Cc], b [N,Tb,V,Cc] -> (cost [N] f32, path [N,Ta+Tb-1,2] i32 (-1 padded), path_len [N] i32)."""
    torch = _torch()
    on_dev = isinstance(a, torch.Tensor) and a.is_cuda
    if on_dev:
        a = a.contiguous().float()
        b = _dev_tensor(torch, b, "b", a.device.index, torch.float32)
    else:
        was_numpy = not isinstance(a, torch.Tensor)
        a = torch.as_tensor(np.asarray(a) if was_numpy else a, dtype=torch.float32).contiguous()
        b = torch.as_tensor(np.asarray(b) if was_numpy else b, dtype=torch.float32).contiguous()
    if a.dim() != 4 or b.dim() != 4 or a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
        raise GolferError(f"align expects a [N,Ta,V,Cc] and b [N,Tb,V,Cc]; got {tuple(a.shape)} {tuple(b.shape)}")
    N, Ta, V, Cc = (int(s) for s in a.shape)
    Tb = int(b.shape[1])
    if Cc < 2:
        raise GolferError("align needs at least (x, y) channels")
    if Ta < 1 or Tb < 1:
        raise GolferError("align needs at least one frame per sequence")
    dev = _current_device(torch, a if on_dev else None)
    ctx = ctx or _align_ctx(dev)
    maxL = Ta + Tb - 1
    kw = dict(device=a.device) if on_dev else dict(pin_memory=True)
    cost = torch.empty((N,), dtype=torch.float32, **kw)
    path = torch.empty((N, maxL, 2), dtype=torch.int32, **kw) if want_path else None
    plen = torch.empty((N,), dtype=torch.int32, **kw) if want_path else None
    if N:
        if on_dev:
            _check(ctx._L.gs_align(ctx.handle, a.data_ptr(), b.data_ptr(), N, Ta, Tb, V, Cc,
                                   cost.data_ptr(), path.data_ptr() if want_path else None,
                                   plen.data_ptr() if want_path else None, _stream_ptr(torch)),
                   "gs_align")
        else:
            _check(ctx._L.gs_align_host(ctx.handle, a.data_ptr(), b.data_ptr(), N, Ta, Tb, V, Cc,
                                        cost.data_ptr(), path.data_ptr() if want_path else None,
                                        plen.data_ptr() if want_path else None), "gs_align_host")
    return cost, path, plen


class AlignTicket:
    """One alignment batch in flight behind align_submit: keeps its host buffers alive until align_wait."""
    __slots__ = ("id", "ctx", "a", "b", "cost", "path", "plen")

    def __init__(self, tid, ctx, a, b, cost, path, plen):
        self.id, self.ctx, self.a, self.b, self.cost, self.path, self.plen = tid, ctx, a, b, cost, path, plen


def align_submit(a, b, ctx: Optional[Context] = None, want_path: bool = True, device: int = 0) -> AlignTicket:
    """Pipelined host alignment (gs_align_host_submit): a [N,Ta,V,Cc], b [N,Tb,V,Cc] contiguous fp32 HOST tensors
    (pinned for overlap) -> a ticket; at most two batches in flight, the copies of one under the sweep of the other.
    `align_wait(ticket)` returns (cost, path, path_len) host tensors."""
    torch = _torch()
    for name, t in (("a", a), ("b", b)):
        if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise GolferError(f"align_submit: {name} must be a contiguous fp32 host tensor (pinned for overlap)")
    if a.dim() != 4 or b.dim() != 4 or a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
        raise GolferError(f"align expects a [N,Ta,V,Cc] and b [N,Tb,V,Cc]; got {tuple(a.shape)} {tuple(b.shape)}")
    N, Ta, V, Cc = (int(s) for s in a.shape)
    Tb = int(b.shape[1])
    if N < 1 or Cc < 2 or Ta < 1 or Tb < 1:
        raise GolferError("align_submit needs at least one pair, one frame per sequence and (x, y) channels")
    ctx = ctx or _align_ctx(device)
    maxL = Ta + Tb - 1
    cost = torch.empty((N,), dtype=torch.float32, pin_memory=True)
    path = torch.empty((N, maxL, 2), dtype=torch.int32, pin_memory=True) if want_path else None
    plen = torch.empty((N,), dtype=torch.int32, pin_memory=True) if want_path else None
    tk = ctypes.c_int(-1)
    _check(ctx._L.gs_align_host_submit(ctx.handle, a.data_ptr(), b.data_ptr(), N, Ta, Tb, V, Cc, cost.data_ptr(),
                                       path.data_ptr() if want_path else None,
                                       plen.data_ptr() if want_path else None, ctypes.byref(tk)),
           "gs_align_host_submit")
    return AlignTicket(tk.value, ctx, a, b, cost, path, plen)


def align_wait(ticket: AlignTicket):
    _check(ticket.ctx._L.gs_align_host_wait(ticket.ctx.handle, ticket.id), "gs_align_host_wait")
    return ticket.cost, ticket.path, ticket.plen


def align(a, b):
    """align(a, b) -> (cost, path).

    One pair: a [Ta,V,Cc], b [Tb,V,Cc] -> (cost scalar tensor, path [L,2] int32).
    Batch: a [N,Ta,V,Cc], b [N,Tb,V,Cc] -> (cost [N], path [N,Ta+Tb-1,2], rows past each
    pair's length filled with -1)."""
    torch = _torch()
    single = (a.ndim == 3)
    if single:
        a, b = a[None], b[None]
    cost, path, plen = align_batch(a, b)
    if single:
        n = int(plen[0])
        return cost[0], path[0, :n]
    return cost, path


def pair_cost(a, b, ctx: Optional[Context] = None):
    """Cost matrices only: a [N,Ta,V,Cc], b [N,Tb,V,Cc] (CUDA tensors) -> [N,Ta,Tb] fp32."""
    torch = _torch()
    dev = _current_device(torch, a)
    a = _dev_tensor(torch, a, "a", dev, torch.float32)
    b = _dev_tensor(torch, b, "b", dev, torch.float32)
    if a.dim() != 4 or b.dim() != 4 or a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
        raise GolferError(f"pair_cost expects a [N,Ta,V,Cc] and b [N,Tb,V,Cc]; got {tuple(a.shape)} {tuple(b.shape)}")
    N, Ta, V, Cc = (int(s) for s in a.shape)
    Tb = int(b.shape[1])
    ctx = ctx or _align_ctx(dev)
    out = torch.empty((N, Ta, Tb), dtype=torch.float32, device=a.device)
    if N:
        _check(ctx._L.gs_pair_cost(ctx.handle, a.data_ptr(), b.data_ptr(), N, Ta, Tb, V, Cc,
                                   out.data_ptr(), _stream_ptr(torch)), "gs_pair_cost")
    return out


def compare(a, b, path, path_len, ctx: Optional[Context] = None):
    """"Compare 2 skeleton" (README.md:50-52): per aligned step, per joint distance.
    a [N,Ta,V,Cc], b [N,Tb,V,Cc], path [N,Ta+Tb-1,2], path_len [N] (CUDA) -> [N,Ta+Tb-1,V] fp32."""
    torch = _torch()
    dev = _current_device(torch, a)
    a = _dev_tensor(torch, a, "a", dev, torch.float32)
    b = _dev_tensor(torch, b, "b", dev, torch.float32)
    if a.dim() != 4 or b.dim() != 4 or a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
        raise GolferError(f"compare expects a [N,Ta,V,Cc] and b [N,Tb,V,Cc]; got {tuple(a.shape)} {tuple(b.shape)}")
    N, Ta, V, Cc = (int(s) for s in a.shape)
    Tb = int(b.shape[1])
    path = _dev_tensor(torch, path, "path", dev, torch.int32)              # int64 paths are converted, not reinterpreted
    path_len = _dev_tensor(torch, path_len, "path_len", dev, torch.int32)
    if tuple(path.shape) != (N, Ta + Tb - 1, 2) or tuple(path_len.shape) != (N,):
        raise GolferError(f"path must be [N,Ta+Tb-1,2] and path_len [N]; got {tuple(path.shape)} {tuple(path_len.shape)}")
    if N:
        # the kernel indexes a and b with the path entries: reject anything outside the two sequences
        L = path_len.clamp(min=0).to(torch.int64)
        live = torch.arange(Ta + Tb - 1, device=path.device)[None, :] < L[:, None]
        bad = (path_len < 1) | (path_len > Ta + Tb - 1)
        oob = live & ((path[..., 0] < 0) | (path[..., 0] >= Ta) | (path[..., 1] < 0) | (path[..., 1] >= Tb))
        if bool(bad.any()) or bool(oob.any()):
            raise GolferError("compare: path / path_len hold entries outside the two sequences")
    ctx = ctx or _align_ctx(dev)
    out = torch.empty((N, Ta + Tb - 1, V), dtype=torch.float32, device=a.device)
    if N:
        _check(ctx._L.gs_compare(ctx.handle, a.data_ptr(), b.data_ptr(), path.data_ptr(),
                                 path_len.data_ptr(), N, Ta, Tb, V, Cc, out.data_ptr(),
                                 _stream_ptr(torch)), "gs_compare")
    return out


def normalize_pose(kp, min_score: float = 0.3, ctx: Optional[Context] = None):
    """Pose-estimation keypoints -> network input (C ABI gs_normalize_pose).

    kp [B,T,V,3] or [T,V,3] = (x, y, score) in image coordinates, COCO-17 joint order.  Returns the
    hip-centred, torso-scaled skeletons with low-score joints zeroed, same shape, fp32.  Device tensors stay on
    the device (current stream); host tensors / arrays are copied in and out."""
    torch = _torch()
    on_dev = isinstance(kp, torch.Tensor) and kp.is_cuda
    was_numpy = not isinstance(kp, torch.Tensor)
    x = kp if on_dev else torch.as_tensor(np.asarray(kp) if was_numpy else kp)
    x = x.to(torch.float32).contiguous()
    single = x.dim() == 3
    if single:
        x = x.unsqueeze(0)
    if x.dim() != 4 or x.shape[-1] != 3 or x.shape[2] < 13:
        raise GolferError(f"normalize_pose expects [B,T,V>=13,3] (x, y, score); got {tuple(kp.shape)}")
    B, T, V = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
    if T < 1:
        raise GolferError("normalize_pose needs at least one frame")
    dev = _current_device(torch, x if on_dev else None)
    ctx = ctx or _align_ctx(dev)
    xd = x if on_dev else x.to(f"cuda:{dev}")
    out = torch.empty_like(xd)
    if B:
        with torch.cuda.device(dev):
            _check(ctx._L.gs_normalize_pose(ctx.handle, xd.data_ptr(), out.data_ptr(), B, T, V,
                                            ctypes.c_float(min_score), _stream_ptr(torch)), "gs_normalize_pose")
    if single:
        out = out[0]
    if on_dev:
        return out
    out = out.cpu()
    return out.numpy() if was_numpy else out


def align_phase(a, b, labels_a, labels_b, penalty: float, ctx: Optional[Context] = None, want_path: bool = True):
    """Phase-conditioned alignment (C ABI gs_align_phase): `align_batch` with `penalty` added
    to every cell whose frames carry different phase labels.  Device tensors only (the labels come straight
    from `Segmenter.segment(..., return_labels=True)`): a [N,Ta,V,Cc], b [N,Tb,V,Cc] fp32, labels_a [N,Ta],
    labels_b [N,Tb] u8 -> (cost [N], path [N,Ta+Tb-1,2] (-1 padded), path_len [N])."""
    torch = _torch()
    for t in (a, b, labels_a, labels_b):
        if not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise GolferError("align_phase takes device tensors")
    dev = a.device.index
    a = _dev_tensor(torch, a, "a", dev, torch.float32)
    b = _dev_tensor(torch, b, "b", dev, torch.float32)
    if a.dim() != 4 or b.dim() != 4 or a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
        raise GolferError(f"align_phase expects a [N,Ta,V,Cc] and b [N,Tb,V,Cc]; got {tuple(a.shape)} {tuple(b.shape)}")
    N, Ta, V, Cc = (int(x) for x in a.shape)
    Tb = int(b.shape[1])
    if tuple(labels_a.shape) != (N, Ta) or tuple(labels_b.shape) != (N, Tb):
        raise GolferError(f"labels must be [N,Ta] and [N,Tb]; got {tuple(labels_a.shape)} {tuple(labels_b.shape)}")
    if Cc < 2 or Ta < 1 or Tb < 1:
        raise GolferError("align_phase needs (x, y) channels and at least one frame per sequence")
    la = _dev_tensor(torch, labels_a, "labels_a", dev, torch.uint8)
    lb = _dev_tensor(torch, labels_b, "labels_b", dev, torch.uint8)
    ctx = ctx or _align_ctx(dev)
    maxL = Ta + Tb - 1
    cost = torch.empty((N,), dtype=torch.float32, device=a.device)
    path = torch.empty((N, maxL, 2), dtype=torch.int32, device=a.device) if want_path else None
    plen = torch.empty((N,), dtype=torch.int32, device=a.device) if want_path else None
    if N:
        _check(ctx._L.gs_align_phase(ctx.handle, a.data_ptr(), b.data_ptr(), la.data_ptr(), lb.data_ptr(),
                                     ctypes.c_float(penalty), N, Ta, Tb, V, Cc, cost.data_ptr(),
                                     path.data_ptr() if want_path else None,
                                     plen.data_ptr() if want_path else None, _stream_ptr(torch)), "gs_align_phase")
    return cost, path, plen


class EmbedAligner:
    """Learned alignment embedding (C ABI gs_set_align_encoder / gs_align_embed): a per-frame
    encoder, the frame-to-frame cost as a tensor-core GEMM of the embeddings, then the DTW of `align`.
    `blob`: fp32 {W1 [34,128], b1, W2 [128,128], b2} (AlignEmbedConfig v0; the reference ships neither encoder nor
    loss, so the weights are the caller's - the tests and the bench use seeded random ones)."""

    def __init__(self, blob: np.ndarray, device: int = 0):
        self.ctx = Context(device)
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        _check(self.ctx._L.gs_set_align_encoder(self.ctx.handle, blob.ctypes.data, blob.nbytes), "gs_set_align_encoder")

    def align(self, a, b, want_path: bool = True, want_cost_matrix: bool = False):
        """a [N,Ta,17,Cc], b [N,Tb,17,Cc] CUDA fp32 -> (cost [N], path [N,Ta+Tb-1,2] (-1 padded), path_len [N]
        [, cost matrix [N,max(Ta,Tb),min(Ta,Tb)]: rows = frames of the LONGER sequence])."""
        torch = _torch()
        dev = self.ctx.device
        a = _dev_tensor(torch, a, "a", dev, torch.float32)
        b = _dev_tensor(torch, b, "b", dev, torch.float32)
        if a.dim() != 4 or b.dim() != 4 or a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
            raise GolferError(f"align expects a [N,Ta,V,Cc] and b [N,Tb,V,Cc]; got {tuple(a.shape)} {tuple(b.shape)}")
        N, Ta, V, Cc = (int(s) for s in a.shape)
        Tb = int(b.shape[1])
        if Cc < 2 or Ta < 1 or Tb < 1:
            raise GolferError("align needs (x, y) channels and at least one frame per sequence")
        cost = torch.empty((N,), dtype=torch.float32, device=a.device)
        path = torch.empty((N, Ta + Tb - 1, 2), dtype=torch.int32, device=a.device) if want_path else None
        plen = torch.empty((N,), dtype=torch.int32, device=a.device) if want_path else None
        cm = torch.empty((N, max(Ta, Tb), min(Ta, Tb)), dtype=torch.float32, device=a.device) if want_cost_matrix else None
        if N:
            with torch.cuda.device(dev):
                _check(self.ctx._L.gs_align_embed(self.ctx.handle, a.data_ptr(), b.data_ptr(), N, Ta, Tb, V, Cc,
                                                  cost.data_ptr(), path.data_ptr() if want_path else None,
                                                  plen.data_ptr() if want_path else None,
                                                  cm.data_ptr() if want_cost_matrix else None, _stream_ptr(torch)),
                       "gs_align_embed")
        return (cost, path, plen, cm) if want_cost_matrix else (cost, path, plen)
