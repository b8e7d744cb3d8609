"""Drop-in ``LPSR`` module: same constructor, ``forward`` signature and ``state_dict`` layout as the reference class
``my_models.lpsr.LPSR`` (reference my_models/lpsr.py:252-274), with the arithmetic done by the sm_100a kernels behind
the C ABI (include/lpsr_b200.h).

What is mirrored from the reference
  * ctor ``LPSR(num_channels, num_features, growth_rate, num_blocks, num_layers, scale_factor, out_channels=1)``;
    ``scale_factor`` is accepted and ignored exactly like the reference (lpsr.py:261,272).
  * the 158-key ``state_dict``: parameters are registered under the same module paths, in the same
    order, with the same aliasing (``auto_encoder.GA.0`` is ``auto_encoder.encoder``; ``rdn.rdbs.{1,3,5,7}`` are the one
    shared ``rdn.csar``) and the two never-executed RDBs ``rdn.rdbs.{4,6}`` (Q3), so ``load_state_dict(strict=True)`` of
    ``weights/best_model.pth`` works and ``state_dict()`` round-trips.  Parameter holders are created in the reference's
    construction order with stock torch initialisers, so ``torch.manual_seed(s); LPSR(...)`` gives the same random
    init as the reference.
  * ``forward(x[B,3,H,W] fp32) -> [B,1,ceil4(H),ceil4(W)]`` fp32 on x's device; pad-to-4 is never cropped back (Q8).

What is different: inference only (no autograd through the kernels), CUDA sm_100 only -- a CPU tensor raises instead of
silently falling back -- and a ``precision`` knob ("fp32" parity mode, "bf16"/"fp16" tensor-core modes).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import capi

_CANONICAL = (3, 32, 16, 4, 4, 1)


class _Holder(nn.Module):
    """Parameter container; has no forward of its own (the kernels consume the parameters)."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            setattr(self, k, v)


def _slots(*mods) -> nn.Sequential:
    # nn.Sequential only for its integer child names ("0", "1", ...): parameter-free slots hold nn.Identity so the
    # indices of the parameterised slots match the reference's Sequentials.
    return nn.Sequential(*[m if m is not None else nn.Identity() for m in mods])


def _conv(cin, cout, k, bias=True, groups=1):
    return nn.Conv2d(cin, cout, k, padding=k // 2, bias=bias, groups=groups)


def _dconv(cin, cout, k=5):
    # depthwise kxk (+bias) then pointwise 1x1 (+bias), registered as "dConv.0" / "dConv.1"  (lpsr.py:8-28)
    return _Holder(dConv=_slots(_conv(cin, cin, k, groups=cin), _conv(cin, cout, 1)))


def _rdb(feat, growth, layers):
    # "layers.i.conv", "lff", "alpha"  (lpsr.py:43-61); registration order: layers, lff, alpha
    h = _Holder(layers=_slots(*[_Holder(conv=_conv(feat + growth * i, growth, 3)) for i in range(layers)]),
                lff=_conv(feat + growth * layers, feat, 1))
    h.alpha = nn.Parameter(torch.tensor(1.0))
    return h


def _csar(feat):
    # "conv_in.{0,2}", "ca.block.{2,4}", "sa.block.{0,2}", "conv_out"  (lpsr.py:156-178)
    return _Holder(
        conv_in=_slots(_conv(feat, feat, 3), None, _conv(feat, feat, 3)),
        ca=_Holder(block=_slots(None, None, nn.Linear(feat, feat // 4), None, nn.Linear(feat // 4, feat), None, None)),
        sa=_Holder(block=_slots(_conv(feat, 2 * feat, 1), None, _conv(2 * feat, feat, 1), None)),
        conv_out=_conv(2 * feat, feat, 1))


class LPSR(nn.Module):
    def __init__(self, num_channels, num_features, growth_rate, num_blocks, num_layers, scale_factor, out_channels=1,
                 precision: Optional[str] = None):
        super().__init__()
        dims = (num_channels, num_features, growth_rate, num_blocks, num_layers, out_channels)
        if dims != _CANONICAL:
            raise ValueError(f"lpsr_b200 kernels are specialised on LPSR{_CANONICAL[:5]} with out_channels=1 (the literals "
                             f"at every reference call site, e.g. inference/run.py:124); got {dims}")
        self.scale_factor = scale_factor  # ignored, as in the reference
        self._dims = dims
        self.precision = precision or os.environ.get("LPSR_B200_PRECISION", "fp32")
        if self.precision not in capi.PREC:
            raise ValueError(f"precision must be one of {sorted(capi.PREC)}, got {self.precision!r}")
        C3, F, G, NB, NL, OC = dims
        E = 4 * C3
        # ---- AutoEncoder parameter holders, reference registration order (lpsr.py:64-104)
        ae = _Holder(conv_in=_conv(C3, E, 3, bias=False))
        ae.encoder = _slots(_dconv(E, E), None, None, _dconv(4 * E, E), None, None)
        ae.decoder = _slots(_dconv(4 * E, 4 * E), None, None, _dconv(E, 4 * E), None, None)
        ae.GA = _slots(ae.encoder, ae.decoder)        # alias keys auto_encoder.GA.{0,1}.* 
        ae.conv_out = _conv(E, C3, 3, bias=False)
        self.auto_encoder = ae
        # ---- RDN parameter holders (lpsr.py:189-212)
        rdn = _Holder(shallowF1=_conv(C3, F, 7), shallowF2=_conv(F, F, 3), csar=_csar(F))
        rdn.rdbs = nn.ModuleList()
        for _ in range(NB):
            rdn.rdbs.append(_rdb(F, G, NL))
            rdn.rdbs.append(rdn.csar)                 # one shared CSAR registered 4x
        rdn.gff = _slots(_conv(F * NB, F, 1), _conv(F, F, 3))
        self.rdn = rdn
        self.final_conv = _conv(F, OC, 3)
        self._reset_native_state()

    # ---- native state (not part of state_dict, never copied or pickled) --------------------------------
    _NATIVE_ATTRS = ("_handles", "_weights_key", "_workspaces", "_live_refs", "_ws_streams")

    def _reset_native_state(self):
        # raw lpsr_handle pointers, the (data_ptr, version) key of the weights they were packed from, scratch tensors, and the
        # (owner module, attribute) path of every live tensor (resolved once per handle)
        self.__dict__["_handles"] = {}
        self.__dict__["_weights_key"] = {}
        self.__dict__["_workspaces"] = {}
        self.__dict__["_live_refs"] = {}
        self.__dict__["_ws_streams"] = {}      # streams that used each workspace, most recent last

    def __getstate__(self):
        # copy.deepcopy / pickle / torch.save(model): a copy must never share (and later double-free) the native handles
        state = self.__dict__.copy()
        for k in self._NATIVE_ATTRS:
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._reset_native_state()

    # ------------------------------------------------------------------------------------------------
    def live_tensors(self) -> Dict[str, torch.Tensor]:
        """The 64 tensors the executed graph reads, keyed by their first state_dict name."""
        lib = capi.load_library()
        sd = dict(self.named_parameters(remove_duplicate=True))
        names = self._live_names(lib)
        return {n: sd[n] for n in names}

    def _live_names(self, lib):
        h = next(iter(self._handles.values()), None)
        if h is None:
            raise capi.LpsrError("no native handle yet")
        return [lib.lpsr_live_tensor_name(h, i).decode() for i in range(lib.lpsr_num_live_tensors(h))]

    def set_precision(self, precision: str) -> "LPSR":
        if precision not in capi.PREC:
            raise ValueError(f"precision must be one of {sorted(capi.PREC)}")
        self.precision = precision
        return self

    def _handle(self, device: torch.device) -> int:
        lib = capi.load_library()
        key = (device.index if device.index is not None else torch.cuda.current_device(), self.precision)
        h = self._handles.get(key)
        if h is None:
            cfg = capi.LpsrConfig(capi.ABI_VERSION, key[0], *self._dims[:5], self._dims[5], capi.PREC[self.precision])
            out = C.c_void_p()
            capi.check(lib.lpsr_create(C.byref(out), C.byref(cfg)), None, "lpsr_create")
            h = out.value
            self._handles[key] = h
        # (re)pack when parameters changed: load_state_dict, .to(), optimiser steps bump _version / data_ptr.  The per-call cost is
        # 64 getattr + data_ptr reads (the call site runs batch 1, inference/run.py:200-202): the name -> (module, attribute) paths
        # are resolved once per handle instead of walking named_parameters() on every forward.
        refs = self._live_refs.get(key)
        if refs is None:
            n = lib.lpsr_num_live_tensors(h)
            refs = []
            for i in range(n):
                nm = lib.lpsr_live_tensor_name(h, i).decode()
                mod_path, _, attr = nm.rpartition(".")
                refs.append((nm, self.get_submodule(mod_path) if mod_path else self, attr))
            self._live_refs[key] = refs
        n = len(refs)
        live = [getattr(owner, attr) for _, owner, attr in refs]
        wkey = tuple((t.data_ptr(), t._version) for t in live)
        if self._weights_key.get(key) != wkey:
            descs = (capi.LpsrTensorDesc * n)()
            keep = []
            for i, (nm, _, _) in enumerate(refs):
                t = live[i].detach()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.float().contiguous()
                if t.is_cuda and t.device.index != key[0]:
                    t = t.cpu()
                keep.append(t)
                descs[i] = capi.LpsrTensorDesc(nm.encode(), t.data_ptr(), t.numel(), 1 if t.is_cuda else 0, 0)
            if any(t.is_cuda for t in keep):
                torch.cuda.synchronize(key[0])
            capi.check(lib.lpsr_load_weights(h, descs, n), h, "lpsr_load_weights")
            self._weights_key[key] = wkey
        return h

    def _workspace(self, h: int, dev: int, B: int, H: int, W: int) -> torch.Tensor:
        key = (dev, B, H, W, self.precision)
        ws = self._workspaces.get(key)
        if ws is None:
            lib = capi.load_library()
            nbytes = lib.lpsr_workspace_bytes(h, B, H, W)
            # keep at most one workspace per device/precision: shapes change rarely at the call sites.  Kernels of an earlier forward
            # (possibly on another stream) may still be using the old scratch: tell the caching allocator before dropping it.
            for k in [k for k in self._workspaces if k[0] == dev and k[4] == self.precision]:
                old = self._workspaces.pop(k)
                for st in self._ws_streams.pop(k, ()):
                    old.record_stream(st)
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=torch.device("cuda", dev))
            self._workspaces[key] = ws
        return ws

    @staticmethod
    def _aligned_ptr(ws: torch.Tensor) -> int:
        return (ws.data_ptr() + 255) // 256 * 256

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or x.shape[1] != self._dims[0]:
            raise RuntimeError(f"LPSR.forward expects a [B,{self._dims[0]},H,W] tensor, got {tuple(getattr(x, 'shape', ()))}")
        if not x.is_cuda:
            raise RuntimeError("lpsr_b200.LPSR runs on CUDA sm_100 only; there is no CPU fallback (move the input with .to('cuda'))")
        p0 = self.final_conv.weight
        if p0.device != x.device:
            raise RuntimeError(f"Input is on {x.device} but the module's parameters are on {p0.device} (call model.to(device))")
        lib = capi.load_library()
        dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(dev):
            xin = x.detach()
            if xin.dtype != torch.float32:
                xin = xin.float()
            xin = xin.contiguous()
            B, _, H, W = xin.shape
            h = self._handle(x.device)
            Hp, Wp = (H + 3) // 4 * 4, (W + 3) // 4 * 4
            y = torch.empty((B, self._dims[5], Hp, Wp), dtype=torch.float32, device=x.device)
            if B == 0:
                return y
            ws = self._workspace(h, dev, B, H, W)
            cur = torch.cuda.current_stream(dev)
            # the scratch is shared by every forward of this shape: a forward on a different stream than the previous one waits for it
            # (not while `cur` is being captured into a CUDA graph: a capturing stream cannot wait on work outside its graph -- the caller orders
            # the capture after earlier forwards, as torch.cuda.graph does by synchronising the device first)
            used = self._ws_streams.setdefault((dev, B, H, W, self.precision), [])
            if not torch.cuda.is_current_stream_capturing():
                if used and used[-1] != cur:
                    cur.wait_stream(used[-1])
                if cur not in used:
                    used.append(cur)
                elif used[-1] != cur:
                    used.remove(cur); used.append(cur)
            capi.check(lib.lpsr_forward(h, xin.data_ptr(), y.data_ptr(), B, H, W, self._aligned_ptr(ws),
                                        ws.numel() - (self._aligned_ptr(ws) - ws.data_ptr()), cur.cuda_stream), h, "lpsr_forward")
        return y

    @torch.no_grad()
    def forward_host(self, x_cpu: torch.Tensor, out: Optional[torch.Tensor] = None, device: Optional[int] = None) -> torch.Tensor:
        """`sr_model(x.to(device)).cpu()` of the reference call site (inference/run.py:201-202) as ONE C-ABI call on
        host buffers: H2D copy, forward and D2H copy happen inside ``lpsr_forward_host``."""
        if not isinstance(x_cpu, torch.Tensor) or x_cpu.dim() != 4 or x_cpu.shape[1] != self._dims[0]:
            raise RuntimeError(f"LPSR.forward_host expects a [B,{self._dims[0]},H,W] tensor, got {tuple(getattr(x_cpu, 'shape', ()))}")
        if x_cpu.is_cuda:
            raise RuntimeError("forward_host takes a CPU tensor")
        lib = capi.load_library()
        dev = device if device is not None else (self.final_conv.weight.device.index or 0)
        x_cpu = x_cpu.float().contiguous()
        B, _, H, W = x_cpu.shape
        Hp, Wp = (H + 3) // 4 * 4, (W + 3) // 4 * 4
        shape = (B, self._dims[5], Hp, Wp)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, pin_memory=torch.cuda.is_available())
        elif (not isinstance(out, torch.Tensor) or out.is_cuda or out.dtype != torch.float32 or tuple(out.shape) != shape
              or not out.is_contiguous()):
            raise RuntimeError(f"forward_host: `out` must be a contiguous CPU float32 tensor of shape {shape}")
        if B == 0 or H == 0 or W == 0:
            return out
        h = self._handle(torch.device("cuda", dev))
        capi.check(lib.lpsr_forward_host(h, x_cpu.data_ptr(), out.data_ptr(), B, H, W), h, "lpsr_forward_host")
        return out

    @staticmethod
    def canonical_key(key: str) -> str:
        """Map any of the 158 state_dict keys to the name of the tensor it aliases."""
        if key.startswith("auto_encoder.GA.0."):
            return "auto_encoder.encoder." + key[len("auto_encoder.GA.0."):]
        if key.startswith("auto_encoder.GA.1."):
            return "auto_encoder.decoder." + key[len("auto_encoder.GA.1."):]
        if key.startswith("rdn.rdbs."):
            parts = key.split(".")
            if int(parts[2]) % 2 == 1:
                return "rdn.csar." + ".".join(parts[3:])
        return key

    def load_live_weights(self, live: Dict[str, "torch.Tensor"]) -> "LPSR":
        """Load the 64 executed tensors (numpy arrays or tensors keyed by their canonical names) by expanding them to
        the full 158-key layout; the never-executed RDBs keep their current values."""
        sd = self.state_dict()
        for k in sd:
            kk = self.canonical_key(k)
            if kk in live:
                sd[k] = torch.as_tensor(live[kk], dtype=torch.float32).reshape(sd[k].shape)
        self.load_state_dict(sd, strict=True)
        return self

    # ---- test / introspection helpers -----------------------------------------------------------------
    def launch_count(self, B: int, H: int, W: int, device: Optional[torch.device] = None) -> int:
        lib = capi.load_library()
        h = self._handle(device or self.final_conv.weight.device)
        return lib.lpsr_forward_launch_count(h, B, H, W)

    @torch.no_grad()
    def read_tap(self, name: str, x_shape, channels: int, div: int = 1) -> torch.Tensor:
        """Intermediate of the LAST forward for input shape ``x_shape`` as fp32 NCHW (tests only)."""
        lib = capi.load_library()
        B, _, H, W = x_shape
        dev = self.final_conv.weight.device
        h = self._handle(dev)
        Hp, Wp = (H + 3) // 4 * 4, (W + 3) // 4 * 4
        out = torch.empty((B, channels, Hp // div, Wp // div), dtype=torch.float32, device=dev)
        ws = self._workspace(h, dev.index or 0, B, H, W)
        capi.check(lib.lpsr_debug_read_tap(h, name.encode(), out.data_ptr(), out.numel(), B, H, W, self._aligned_ptr(ws),
                                           torch.cuda.current_stream(dev).cuda_stream), h, "lpsr_debug_read_tap")
        return out

    def __del__(self):
        try:
            handles = self.__dict__.get("_handles") or {}
            if handles:
                lib = capi.load_library()
                for h in list(handles.values()):
                    lib.lpsr_destroy(h)
                handles.clear()
        except Exception:
            pass
