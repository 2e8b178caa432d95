"""Drop-in mirror of dl_cs/mri/transforms.py (FFT, SenseModel) on the B200 kernels.

Same class names, constructor arguments, call signature, attributes and assertions as the
reference (`SenseModel(maps, weights)(data, adjoint=False)`, transforms.py:49-111;
`FFT(ndims, norm)(data, adjoint, centered)`, transforms.py:12-46), so the unrolled solvers
that receive the operator by dependency injection (`forward(self, y, A, x0)`,
unrolledswin.py:91, unrolled.py:86) run unchanged.  Arithmetic happens in the hand-written
sm_100a kernels behind the C ABI of include/dlsg_b200.h; torch is used for device memory,
streams and autograd plumbing only.  There is no CPU path: CPU tensors raise.

Sampled lines: when the weights are a k-t sampling mask (whole ky lines of W are zero, which is what
`get_mask` of undersampled data gives, dl_cs/mri/utils.py:69-79) the operator runs the Y transform first and
keeps only the lines that carry weight (csrc/sense_lines.cuh); same results, a fraction of the HBM traffic.
The plan is built on first use (one pass over W and one scalar read-back) and cached with the weights.
`DLSG_SAMPLED_LINES=0` forces the dense kernels (`=force` also takes the path for E > 2, where it is the slower one).  On that path `A(y, adjoint=True)` also accepts a k-space that was
left in *pinned host memory*: the kernels read its sampled lines straight over PCIe, so the zero-filled k-space
(the bulk of a reconstruction's input, scripts/reconstruct.py:132-171) never needs a dense device copy.

Beyond the reference interface the operator offers two fused entries the solvers may use
when present (they fall back to two calls otherwise):
    A.normal(x)              == A(A(x), adjoint=True)
    A.dc_step(x, ATy, step)  == x + step * (A(A(x), adjoint=True) - ATy)      (unrolledswin.py:109)
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch
from torch import nn

from .. import _lib

_WORKSPACE_MB_DEFAULT = 4096   # measured: an unchunked intermediate beats L2-sized chunks (profiles/r1a_workspace_sweep.log)


def _workspace_budget() -> int:
    return int(float(os.environ.get("DLSG_WORKSPACE_MB", _WORKSPACE_MB_DEFAULT)) * (1 << 20))


def _require_cuda_c64(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} is on {t.device}: dl_swin_gan_b200 runs on CUDA (sm_100a) only and has no CPU fallback")
    if t.dtype != torch.complex64:
        raise TypeError(f"{what} must be complex64 (got {t.dtype}); the B200 kernels compute in fp32")


# --------------------------------------------------------------------------------------
# low-level launcher shared by SenseModel and FFT
# --------------------------------------------------------------------------------------
_LINES_MAX_FRACTION = 0.5      # use the sampled-line kernels when at most this share of the ky lines carries weight

# Host-resident k-space handed to the asynchronous line gather: (tensor, event recorded after the launch).  The kernel
# reads the pinned buffer over PCIe after the call has returned, so the buffer must neither be freed nor recycled by
# torch's pinned-memory allocator before that event has completed; entries are dropped once it has.
_HOST_INFLIGHT = []


def _retain_host_buffer(t: torch.Tensor, device) -> None:
    _HOST_INFLIGHT[:] = [(h, e) for (h, e) in _HOST_INFLIGHT if not e.query()]
    ev = torch.cuda.Event()
    with torch.cuda.device(device):
        ev.record(torch.cuda.current_stream())
    _HOST_INFLIGHT.append((t, ev))


def _sampled_lines_enabled() -> bool:
    return os.environ.get("DLSG_SAMPLED_LINES", "1") != "0"


def _launch_lines(kind: str, data, maps5, wexp, dims, lines, aty=None, step=None) -> torch.Tensor:
    """Same contract as _launch, through the sampled-line entry points; lines = (plan int32 tensor, lmax)."""
    lib = _lib.load()
    B, E, C, T, Y, X = dims
    dev = maps5.device                      # `data` may be pinned host memory for kind == "adj_host"
    plan, lmax = lines
    out = torch.empty((B, C, T, Y, X) if kind == "fwd" else (B, E, T, Y, X), dtype=torch.complex64, device=dev)
    wstr = (ctypes.c_longlong * 5)(*wexp.stride())
    nbytes = lib.dlsg_sense_lines_workspace_bytes(B, C, T, X, lmax, _workspace_budget())
    work = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        tail = (work.data_ptr(), nbytes, B, E, C, T, Y, X, stream)
        head = (maps5.data_ptr(), wexp.data_ptr(), wstr, plan.data_ptr(), lmax)
        if kind == "fwd":
            st = lib.dlsg_sense_forward_lines_c64(data.data_ptr(), *head, out.data_ptr(), *tail)
        elif kind == "adj":
            st = lib.dlsg_sense_adjoint_lines_c64(data.data_ptr(), *head, out.data_ptr(), *tail)
        elif kind == "adj_host":
            st = lib.dlsg_sense_adjoint_lines_gather_c64(data.data_ptr(), *head, out.data_ptr(), *tail)
        else:
            epi = {"normal": 0, "dc": 1, "reg": 2}[kind]
            if isinstance(step, torch.Tensor):
                s_host, s_dev = 0.0, step.data_ptr()
            else:
                s_host, s_dev = (0.0 if step is None else float(step)), None
            st = lib.dlsg_sense_normal_lines_c64(data.data_ptr(), aty.data_ptr() if aty is not None else None, *head,
                                                 s_host, s_dev, epi, out.data_ptr(), *tail)
    _lib.check(st, f"dlsg_sense_{kind}_lines")
    return out


def _launch(kind: str, data: torch.Tensor, maps5: Optional[torch.Tensor], wexp: Optional[torch.Tensor],
            dims, ndims: int, aty: Optional[torch.Tensor] = None, step=None) -> torch.Tensor:
    """kind in {'fwd','adj','normal','reg','dc'}; dims = (B,E,C,T,Y,X); tensors already validated/contiguous."""
    lib = _lib.load()
    B, E, C, T, Y, X = dims
    dev = data.device
    out_shape = (B, C, T, Y, X) if kind == "fwd" else (B, E, T, Y, X)
    out = torch.empty(out_shape, dtype=torch.complex64, device=dev)
    if out.numel() == 0 or data.numel() == 0:
        return out.zero_()
    if wexp is not None:
        wptr = wexp.data_ptr()
        wstr = (ctypes.c_longlong * 5)(*wexp.stride())
    else:
        wptr, wstr = None, None
    nbytes = lib.dlsg_sense_workspace_bytes(B, E, C, T, Y, X, ndims, _workspace_budget())
    work = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=dev)
    mptr = maps5.data_ptr() if maps5 is not None else None
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        tail = (work.data_ptr(), nbytes, B, E, C, T, Y, X, ndims, stream)
        if kind == "fwd":
            st = lib.dlsg_sense_forward_c64(data.data_ptr(), mptr, wptr, wstr, out.data_ptr(), *tail)
        elif kind == "adj":
            st = lib.dlsg_sense_adjoint_c64(data.data_ptr(), mptr, wptr, wstr, out.data_ptr(), *tail)
        elif kind == "normal":
            st = lib.dlsg_sense_normal_c64(data.data_ptr(), mptr, wptr, wstr, out.data_ptr(), *tail)
        elif kind == "reg":
            if isinstance(step, torch.Tensor):
                lam_host, lam_dev = 0.0, step.data_ptr()
            else:
                lam_host, lam_dev = float(step), None
            st = lib.dlsg_sense_normal_reg_c64(data.data_ptr(), mptr, wptr, wstr, lam_host, lam_dev, out.data_ptr(), *tail)
        elif kind == "dc":
            if isinstance(step, torch.Tensor):
                step_host, step_dev = 0.0, step.data_ptr()
            else:
                step_host, step_dev = float(step), None
            st = lib.dlsg_sense_dc_step_c64(data.data_ptr(), aty.data_ptr(), mptr, wptr, wstr,
                                            step_host, step_dev, out.data_ptr(), *tail)
        else:  # pragma: no cover
            raise AssertionError(kind)
    _lib.check(st, f"dlsg_sense_{kind}")
    return out


# --------------------------------------------------------------------------------------
# FFT  (dl_cs/mri/transforms.py:12-46)
# --------------------------------------------------------------------------------------
class FFT(nn.Module):
    """Centered-optional N-dimensional (i)FFT over the trailing `ndims` axes, norm='ortho' by default."""

    def __init__(self, ndims: int, norm: Optional[str] = "ortho") -> None:
        super().__init__()
        self.ndims = ndims
        self.norm = norm
        self.fft_dims = [i for i in range(-1, -1 - ndims, -1)]

    def forward(self, data: torch.Tensor, adjoint: Optional[bool] = False, centered: Optional[bool] = False):
        assert torch.is_complex(data)  # force complex (transforms.py:33)
        _require_cuda_c64(data, "FFT input")
        if self.ndims == 0:
            return data.clone()
        if self.ndims not in (1, 2):
            raise NotImplementedError(f"FFT over {self.ndims} trailing axes: the B200 kernels cover ndims 1 and 2")
        if data.dim() < self.ndims:
            raise ValueError("FFT input has fewer dimensions than ndims")
        if centered:
            data = torch.fft.ifftshift(data, dim=self.fft_dims)   # index permutation only
        out = _FftFn.apply(data, self.ndims, bool(adjoint), self.norm)
        if centered:
            out = torch.fft.fftshift(out, dim=self.fft_dims)
        return out


def _fft_raw(data: torch.Tensor, ndims: int, inverse: bool) -> torch.Tensor:
    """ortho (i)FFT of the trailing ndims axes through the SENSE entry points with unit maps."""
    shape = data.shape
    X = shape[-1]
    Y = shape[-2] if ndims == 2 else 1
    flat = data.contiguous().reshape(-1, 1, 1, Y, X)      # [B,E=1,T=1,Y,X]; C = 1
    B = flat.shape[0]
    out = _launch("adj" if inverse else "fwd", flat, None, None, (B, 1, 1, 1, Y, X), 2)
    return out.reshape(shape)


class _FftFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, ndims, inverse, norm):
        ctx.cfg = (ndims, inverse, norm)
        out = _fft_raw(data, ndims, inverse)
        n = 1
        for d in data.shape[-ndims:]:
            n *= d
        # kernels are ortho; other torch.fft norm modes are a scalar away
        if norm in (None, "backward"):
            out = out * (n ** -0.5 if inverse else n ** 0.5)
        elif norm == "forward":
            out = out * (n ** 0.5 if inverse else n ** -0.5)
        elif norm != "ortho":
            raise ValueError(f"unknown norm {norm!r}")
        return out

    @staticmethod
    def backward(ctx, grad):
        ndims, inverse, norm = ctx.cfg
        # VJP of a DFT is the conjugate-transpose DFT; the non-ortho scalings swap roles
        back_norm = {"ortho": "ortho", None: "forward", "backward": "forward", "forward": "backward"}[norm]
        return _FftFn.apply(grad, ndims, not inverse, back_norm), None, None, None


# --------------------------------------------------------------------------------------
# SenseModel  (dl_cs/mri/transforms.py:49-111)
# --------------------------------------------------------------------------------------
class SenseModel(nn.Module):
    """Forward  y = (W * F * S) x   and adjoint  x = (S^H * F^H * W) y   SENSE operations.

    Args:
        maps (torch.Tensor[complex64]): [batch, emaps, coils, 1, y, x]
        weights (torch.Tensor, optional): real, broadcastable to [batch, coils, t, y, x]
            (the usual sampling mask is [batch, 1, t, y, x]); None means 1.0.
    The constructor does no device work (the training loop builds one per step,
    scripts/train_swin.py:116); layout preparation happens on first use and is cached.
    """

    def __init__(self, maps: torch.Tensor, weights: Optional[torch.Tensor] = None, max_lines: Optional[int] = None) -> None:
        """`max_lines` (not in the reference; optional): the caller's promise that no (slice, frame) of `weights` has more than this many
        non-zero ky lines - what a k-t mask generator knows about its own pattern.  With it the sampled-line plan is built without
        the one scalar read-back, so constructing and using the operator never synchronises the host with the stream (a pipeline
        of chunks on several streams then overlaps as enqueued).  Lines beyond the promise would be treated as unsampled;
        `lines_overflowed()` tells (it synchronises)."""
        super().__init__()
        assert torch.is_complex(maps)  # force complex (transforms.py:70)
        if max_lines is not None and int(max_lines) < 1:
            raise ValueError("max_lines must be a positive number of ky lines")
        self.max_lines = None if max_lines is None else int(max_lines)
        self.maps = maps
        ndims = len(list(maps[0, 0, 0, ...].squeeze().shape))
        self.fft = FFT(ndims)
        if weights is None:
            self.weights = 1.0
        else:
            self.weights = weights
        self._maps5 = None
        self._maps_key = None
        self._w32 = None
        self._w_key = None
        self._line_plans = {}

    # ---- reference interface ---------------------------------------------------------
    def forward(self, data: torch.Tensor, adjoint: Optional[bool] = False) -> torch.Tensor:
        assert torch.is_complex(data)  # transforms.py:104
        return _SenseFn.apply(data, self, "adj" if adjoint else "fwd")

    def _adjoint_op(self, data: torch.Tensor) -> torch.Tensor:
        return self.forward(data, adjoint=True)

    def _forward_op(self, data: torch.Tensor) -> torch.Tensor:
        return self.forward(data, adjoint=False)

    # ---- fused entries ---------------------------------------------------------------
    def normal(self, data: torch.Tensor) -> torch.Tensor:
        """A(A(x), adjoint=True) in one fused chain (k-space stays in L2)."""
        assert torch.is_complex(data)
        return _SenseFn.apply(data, self, "normal")

    def normal_reg(self, data: torch.Tensor, lamda) -> torch.Tensor:
        """A(A(x), adjoint=True) + lamda * x  — `model_normal` of the MoDL / HQS solvers (unrolledswin.py:154) in one chain."""
        assert torch.is_complex(data)
        if not isinstance(lamda, torch.Tensor):
            lamda = torch.as_tensor(float(lamda), dtype=torch.float32)
        return _RegFn.apply(data, lamda, self)

    def dc_step(self, x: torch.Tensor, aty: torch.Tensor, step) -> torch.Tensor:
        """x + step * (A(A(x), adjoint=True) - aty)   — unrolledswin.py:109 in one fused chain."""
        assert torch.is_complex(x) and torch.is_complex(aty)
        if not isinstance(step, torch.Tensor):
            step = torch.as_tensor(float(step), dtype=torch.float32)   # host scalar, no device work
        return _DcFn.apply(x, aty, step, self)

    # ---- preparation -------------------------------------------------------------------
    @staticmethod
    def _tensor_key(t):
        """Identity of a tensor's current contents: object, storage address and in-place version (None inside inference mode)."""
        try:
            ver = t._version
        except RuntimeError:               # inference tensors do not track versions
            ver = None
        return (id(t), t.data_ptr(), ver)

    def _prepared_maps(self) -> torch.Tensor:
        key = self._tensor_key(self.maps)
        if self._maps5 is None or key != self._maps_key or key[2] is None:
            m = self.maps
            _require_cuda_c64(m, "SenseModel maps")
            if m.dim() != 6 or m.shape[3] != 1:
                raise ValueError(f"maps must be [batch, emaps, coils, 1, y, x]; got {tuple(m.shape)}")
            self._maps5 = m.detach()[:, :, :, 0].contiguous()
            self._maps_key = key
        return self._maps5

    def _prepared_weights(self, dims, device):
        """-> fp32 tensor expanded (stride-0 broadcast, no copy) to [B,C,T,Y,X], or None for 1.0."""
        w = self.weights
        B, E, C, T, Y, X = dims
        if not isinstance(w, torch.Tensor):
            if float(w) == 1.0:
                return None
            if self._w32 is None or self._w_key != ("scalar", float(w)):
                self._w32 = torch.full((1,), float(w), dtype=torch.float32, device=device)
                self._w_key = ("scalar", float(w))
                self._line_plans = {}
            return self._w32.expand(B, C, T, Y, X)
        key = self._tensor_key(w)
        if self._w32 is None or key != self._w_key or key[2] is None:     # `A.weights` reassigned or edited in place: re-derive
            if torch.is_complex(w):
                raise TypeError("SenseModel weights must be real (the reference documents a float mask)")
            if not w.is_cuda:
                raise RuntimeError(f"SenseModel weights are on {w.device}; CUDA tensors only (no CPU fallback)")
            w32 = w.detach()
            if w32.dtype != torch.float32:
                w32 = w32.to(torch.float32)
            self._w32 = w32
            self._w_key = key
            self._line_plans = {}                                          # plans describe the previous weights
        w32 = self._w32
        if w32.dim() > 5:
            raise ValueError(f"weights must broadcast to [batch, coils, t, y, x]; got {tuple(w32.shape)}")
        return w32.expand(B, C, T, Y, X)   # raises like torch broadcasting does when incompatible

    def _line_plan(self, dims, wexp, ndims, any_emaps=False):
        """-> (plan, lmax) when the weights leave at most half of the ky lines sampled, else None (dense kernels).

        One and two ESPIRiT map sets (every reference configuration) run the second-generation column kernels; with more
        the first-generation kernels read the extra sets through the cache and measure slower than the dense kernels
        (tools/compare_paths.py), so E > 2 stays dense unless `any_emaps` (the host-resident k-space adjoint, which only
        exists on this path)."""
        if wexp is None or ndims != 2 or not isinstance(self.weights, torch.Tensor) or not _sampled_lines_enabled():
            return None
        if dims[1] > 2 and not any_emaps and os.environ.get("DLSG_SAMPLED_LINES") != "force":
            return None
        key = tuple(dims)
        if key not in self._line_plans:
            if self.max_lines is None and torch.cuda.is_current_stream_capturing():
                raise RuntimeError(
                    "SenseModel builds its sampled-line plan on first use (one scalar read-back), which is not allowed while a CUDA "
                    "graph is being captured: call the operator once before the capture")
            lib = _lib.load()
            B, E, C, T, Y, X = dims
            plan = None
            if B * T > 0 and lib.dlsg_sense_lines_supported(E, Y, X, 1):
                buf = torch.empty(int(lib.dlsg_sense_line_plan_bytes(B, T, Y)) // 4, dtype=torch.int32, device=wexp.device)
                wstr = (ctypes.c_longlong * 5)(*wexp.stride())
                with torch.cuda.device(wexp.device):
                    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                    _lib.check(lib.dlsg_sense_line_plan_build_capped(wexp.data_ptr(), wstr, B, C, T, Y, X, int(self.max_lines or 0),
                                                                     buf.data_ptr(), stream), "dlsg_sense_line_plan_build_capped")
                if self.max_lines is not None:
                    lmax = min(self.max_lines, Y)                      # the caller's promise: no read-back, no host synchronisation
                else:
                    lmax = max(1, int(buf[: B * T].max().item()))      # the one read-back: sizes the intermediate
                if lmax <= _LINES_MAX_FRACTION * Y and lib.dlsg_sense_lines_supported(E, Y, X, lmax):
                    plan = (buf, lmax)
            self._line_plans[key] = plan
        return self._line_plans[key]

    def lines_overflowed(self) -> bool:
        """True when a `max_lines` promise was broken for any plan built so far (some frame had more sampled lines; they were dropped).
        Reads the plans back, i.e. synchronises."""
        for key, plan in self._line_plans.items():
            if plan is None:
                continue
            B, _, _, T, Y, _ = key
            cls0 = B * T * (1 + 3 * Y)
            if bool(plan[0][cls0: cls0 + B * T * 36].view(B * T, 36)[:, 35].any().item()):
                return True
        return False

    def _dims(self, data: torch.Tensor, kind: str):
        maps5 = self._prepared_maps()
        Bm, E, C, Y, X = maps5.shape
        if data.dim() != 5:
            raise ValueError(f"SENSE data must be 5-D ([batch, emaps|coils, t, y, x]); got {tuple(data.shape)}")
        B, ch, T, Yd, Xd = data.shape
        want = C if kind == "adj" else E
        if (Yd, Xd) != (Y, X) or ch != want or B != Bm:
            raise ValueError(
                f"SENSE {kind}: data {tuple(data.shape)} does not match maps {tuple(self.maps.shape)} "
                f"(expected [{Bm}, {want}, T, {Y}, {X}])")
        return (B, E, C, T, Y, X)

    def _kernel_ndims(self, dims) -> int:
        nd = self.fft.ndims
        Y, X = dims[4], dims[5]
        if nd == 2:
            return 2
        if nd == 1:
            # reference: FFT over the last axis only.  With Y == 1 the 2-D path is identical (and faster).
            return 2 if Y == 1 else 1
        if nd == 0:
            return 2   # Y == X == 1: the transform is the identity either way
        raise NotImplementedError(
            f"maps with a non-singleton 4th axis imply a {nd}-D FFT (transforms.py:76); the B200 kernels cover 1-D/2-D")

    def _run_adjoint_from_host(self, data: torch.Tensor) -> torch.Tensor:
        """A^H y for a zero-filled k-space left in pinned host memory: the kernels read its sampled lines over PCIe."""
        if data.dtype != torch.complex64:
            raise TypeError(f"SENSE input must be complex64 (got {data.dtype}); the B200 kernels compute in fp32")
        dims = self._dims(data, "adj")
        maps5 = self._prepared_maps()
        wexp = self._prepared_weights(dims, maps5.device)
        ndims = self._kernel_ndims(dims)
        lines = self._line_plan(dims, wexp, ndims, any_emaps=True) if data.numel() else None
        if lines is None:
            raise RuntimeError(
                "host-resident k-space is only read in place by the sampled-line kernels (weights with empty ky lines, "
                "2-D, 32*R axes); move the k-space to the device for this operator")
        if not data.is_contiguous():
            # .contiguous() would make a pageable copy the GPU cannot read: ask for the right buffer instead
            raise ValueError("host-resident k-space must be a contiguous pinned tensor (a strided view of a pinned tensor is not)")
        host = data.detach()
        out = _launch_lines("adj_host", host, maps5, wexp, dims, lines)
        _retain_host_buffer(host, maps5.device)        # keeps the pinned buffer alive until the gather has run
        return out

    def _mark_stream_use(self, device) -> None:
        """The operator's cached device tensors (contiguous maps, fp32 weights, line plans) may have been created on another
        stream than the one now launching kernels that read them (e.g. operator built on a transfer stream, solver run on a
        compute stream): tell the caching allocator, so that their memory is not recycled under the running kernels."""
        if torch.cuda.is_current_stream_capturing():
            return
        st = torch.cuda.current_stream(device)
        for t in (self._maps5, self._w32):
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(st)
        for plan in self._line_plans.values():
            if plan is not None:
                plan[0].record_stream(st)

    def _run(self, kind: str, data: torch.Tensor, aty=None, step=None) -> torch.Tensor:
        out = self._run_impl(kind, data, aty, step)
        self._mark_stream_use(out.device)
        return out

    def _run_impl(self, kind: str, data: torch.Tensor, aty=None, step=None) -> torch.Tensor:
        if kind == "adj" and not data.is_cuda and data.is_pinned():
            return self._run_adjoint_from_host(data)
        _require_cuda_c64(data, "SENSE input")
        dims = self._dims(data, kind)
        maps5 = self._prepared_maps()
        if maps5.device != data.device:
            raise RuntimeError(f"SENSE input on {data.device} but maps on {maps5.device}")
        wexp = self._prepared_weights(dims, data.device)
        if wexp is not None and wexp.device != data.device:
            raise RuntimeError(f"SENSE input on {data.device} but weights on {wexp.device}")
        data = data.detach().contiguous()
        if kind == "reg" and isinstance(step, torch.Tensor):
            if not (step.is_cuda and step.dtype == torch.float32 and step.numel() == 1 and step.device == data.device):
                step = float(step)
            else:
                step = step.detach()
        if kind == "dc":
            _require_cuda_c64(aty, "ATy")
            if aty.shape != data.shape:
                raise ValueError(f"ATy {tuple(aty.shape)} must have the shape of x {tuple(data.shape)}")
            aty = aty.detach().contiguous()
            if isinstance(step, torch.Tensor):
                if step.is_cuda and step.dtype == torch.float32 and step.numel() == 1 and step.device == data.device:
                    step = step.detach()
                else:
                    step = float(step)
        ndims = self._kernel_ndims(dims)
        lines = self._line_plan(dims, wexp, ndims) if data.numel() else None
        if lines is not None:
            return _launch_lines(kind, data, maps5, wexp, dims, lines, aty, step)
        return _launch(kind, data, maps5, wexp, dims, ndims, aty, step)


class _SenseFn(torch.autograd.Function):
    """A is linear: VJP(A) = A^H, VJP(A^H) = A, VJP(A^H A) = A^H A.  maps / weights carry no gradient."""

    _BACK = {"fwd": "adj", "adj": "fwd", "normal": "normal"}

    @staticmethod
    def forward(ctx, data, model, kind):
        ctx.model, ctx.kind = model, kind
        return model._run(kind, data)

    @staticmethod
    def backward(ctx, grad):
        return _SenseFn.apply(grad, ctx.model, _SenseFn._BACK[ctx.kind]), None, None


class _RegFn(torch.autograd.Function):
    """x -> A^H A x + lamda x ; self-adjoint in x, d/dlamda = x."""

    @staticmethod
    def forward(ctx, x, lamda, model):
        ctx.model = model
        ctx.save_for_backward(x, lamda)
        return model._run("reg", x, step=lamda)

    @staticmethod
    def backward(ctx, grad):
        x, lamda = ctx.saved_tensors
        gx = glam = None
        if ctx.needs_input_grad[0]:
            gx = _RegFn.apply(grad, lamda, ctx.model)
        if ctx.needs_input_grad[1]:
            glam = (x.conj() * grad).real.sum().reshape(lamda.shape).to(lamda.device)
        return gx, glam, None


class _DcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, aty, step, model):
        ctx.model = model
        ctx.save_for_backward(x, aty, step)
        return model._run("dc", x, aty=aty, step=step)

    @staticmethod
    def backward(ctx, grad):
        x, aty, step = ctx.saved_tensors
        model = ctx.model
        s = step.to(device=grad.device, dtype=torch.float32)
        gx = gaty = gstep = None
        if ctx.needs_input_grad[0]:
            gx = grad + s * _SenseFn.apply(grad, model, "normal")
        if ctx.needs_input_grad[1]:
            gaty = -s * grad
        if ctx.needs_input_grad[2]:
            resid = _SenseFn.apply(x, model, "normal") - aty
            gstep = (resid.conj() * grad).real.sum().reshape(step.shape).to(step.device)
        return gx, gaty, gstep, None
