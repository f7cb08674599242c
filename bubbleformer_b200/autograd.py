"""torch.autograd glue: one Function per sub-network (embed, temporal block, spatial block, debed) so
parameter gradients become ready block by block during backward (bucketed all-reduce can overlap), plus
the flat fp32 master-weight / bf16 operand bank.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import ops


ACTIVE_SINK = None      # set by bubbleformer_b200.parallel.GradSink


class KernelFn(torch.autograd.Function):
    """Generic bridge.  `spec` provides

        spec.names                      parameter names, in the order of *params
        spec.forward(x, aux, pd, save)  -> (out, saved)
        spec.backward(dout, pd, saved, grads, need_dx) -> (dx or None, daux or None)

    `grads` is a dict name -> zero-initialised fp32 tensor (views of one flat buffer) that the kernels
    accumulate into.
    """

    @staticmethod
    def forward(ctx, spec, x, aux, *params):
        pd = dict(zip(spec.names, params))
        save = any(ctx.needs_input_grad)
        out, saved = spec.forward(x, aux, pd, save)
        ctx.spec, ctx.pd, ctx.saved = spec, pd, saved
        ctx.need_dx = ctx.needs_input_grad[1]
        ctx.has_aux = aux is not None
        sink = ACTIVE_SINK
        ctx.sink = sink if (sink is not None and save and params and sink.owns(params[0])) else None
        return out

    @staticmethod
    def backward(ctx, dout):
        spec, pd, saved = ctx.spec, ctx.pd, ctx.saved
        if saved is None:
            raise RuntimeError("bubbleformer_b200: backward called twice or without saved activations")
        names = spec.names
        dout = dout.contiguous()
        if ctx.sink is not None:
            # gradients accumulate straight into the persistent flat buffer; its segment is all-reduced from here
            grads = ctx.sink.views(pd)
            dx, daux = spec.backward(dout, pd, saved, grads, ctx.need_dx)
            ctx.sink.segment_done(pd)
            ctx.saved = None
            return (None, dx, daux if ctx.has_aux else None) + (None,) * len(names)
        sizes = [pd[n].numel() for n in names]
        offs, tot = [], 0
        for s in sizes:
            offs.append(tot)
            tot += (s + 7) // 8 * 8          # keep every view 32-byte aligned
        flat = torch.zeros(max(tot, 1), dtype=torch.float32, device=dout.device)
        grads = {n: flat[o:o + s].view(pd[n].shape) for n, o, s in zip(names, offs, sizes)}
        dx, daux = spec.backward(dout, pd, saved, grads, ctx.need_dx)
        ctx.saved = None
        return (None, dx, daux if ctx.has_aux else None) + tuple(grads[n] for n in names)


def run(spec, x, aux, pd: Dict[str, torch.Tensor]):
    return KernelFn.apply(spec, x, aux, *[pd[n] for n in spec.names])


# ---------------------------------------------------------------------------------------------
# weight bank
# ---------------------------------------------------------------------------------------------
class WeightBank:
    """All parameters of a model as views of ONE flat fp32 buffer, mirrored by a flat bf16 buffer.

    One `bf_cast16` launch per forward refreshes every bf16 GEMM operand (the 1x1-conv / linear weights are
    used exactly as PyTorch stores them, (N, K) row-major, for forward, dgrad and wgrad alike).
    """

    def __init__(self, module: torch.nn.Module):
        self.module = module
        self.flat: Optional[torch.Tensor] = None
        self.flat16: Optional[torch.Tensor] = None
        self.offsets: Dict[int, int] = {}
        self.params: List[torch.nn.Parameter] = []

    def _ok(self) -> bool:
        if self.flat is None or not self.params:
            return False
        base = self.flat.data_ptr()
        if self.params[0].device != self.flat.device:
            return False
        return all(p.data_ptr() == base + 4 * self.offsets[id(p)] for p in self.params)

    def ensure(self) -> None:
        if self._ok():
            return
        params = [p for p in self.module.parameters()]
        if not params:
            return
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("bubbleformer_b200 modules run on CUDA only (no CPU fallback): call .cuda() first")
        offs, tot = {}, 0
        for p in params:
            if p.dtype != torch.float32:
                raise RuntimeError("bubbleformer_b200 keeps fp32 master parameters; bf16/fp16 operand copies are "
                                   "made internally (do not call .half()/.bfloat16() on the model)")
            offs[id(p)] = tot
            tot += (p.numel() + 7) // 8 * 8
        flat = torch.zeros(tot, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p in params:
                o = offs[id(p)]
                flat[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = flat[o:o + p.numel()].view(p.shape)
        self.flat, self.offsets, self.params = flat, offs, params
        self.flat16 = torch.empty(tot, dtype=torch.bfloat16, device=dev)
        self._mirror_version = None          # a rebuilt bank has an uninitialised mirror

    def invalidate(self) -> None:
        """Force the next forward to re-cast the bf16 mirror (call after changing parameters through `.data`, which
        does not bump their version counters)."""
        self._mirror_version = None

    def refresh(self) -> None:
        self.ensure()
        from . import engine
        if engine.EXACT:                     # fp32 validation configuration: the GEMMs read the fp32 masters
            return
        # the fused optimiser step rewrites the mirror itself; any other in-place change of a parameter (copy_,
        # torch.optim, load_state_dict) bumps that parameter's version counter
        if getattr(self, "_mirror_version", None) == self._versions():
            return
        ops.cast16(self.flat, self.flat16)

    def _versions(self) -> int:
        return sum(p._version for p in self.params)

    def mark_mirror_fresh(self) -> None:
        self._mirror_version = self._versions()

    def w16(self, p: torch.Tensor) -> torch.Tensor:
        from . import engine
        if engine.EXACT:
            return p.detach().view(p.shape[0], -1)
        o = (p.data_ptr() - self.flat.data_ptr()) // 4
        if not (0 <= o and o + p.numel() <= self.flat.numel()):
            raise RuntimeError("bubbleformer_b200: parameter is not part of this model's flat weight buffer")
        return self.flat16[o:o + p.numel()].view(p.shape[0], -1)


def adhoc_w16(p: torch.Tensor) -> torch.Tensor:
    """bf16 operand copy of one weight (stand-alone layer use; the full model uses WeightBank)."""
    from . import engine
    if engine.EXACT:
        return p.detach().contiguous().view(p.shape[0], -1)
    out = torch.empty(p.shape[0], p.numel() // p.shape[0], dtype=torch.bfloat16, device=p.device)
    ops.cast16(p.detach().contiguous().reshape(-1), out.reshape(-1))
    return out
