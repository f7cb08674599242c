"""Data-parallel plumbing: flat gradient sink + bucketed NCCL all-reduce overlapped with backward.

Mirrors what upstream gets from Lightning's `strategy="ddp"` (scripts/train.py:158-163): one process per GPU,
gradients averaged over ranks each step.  Here every autograd Function of the model (embed, each temporal /
spatial block, debed) accumulates its parameter gradients straight into one persistent flat fp32 buffer laid
out like the weight bank, and as soon as a Function's backward has finished its contiguous segment is
all-reduced on a side stream while the remaining backward kernels keep running.  `finish()` joins the
streams.  With world size 1 the sink still removes the per-backward gradient allocations and zero fills.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class GradSink:
    def __init__(self, model: torch.nn.Module, process_group=None, bucket_bytes: int = 8 << 20):
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        params = list(model.parameters())
        self.offsets: Dict[int, int] = {}
        tot = 0
        for p in params:
            self.offsets[id(p)] = tot
            tot += (p.numel() + 7) // 8 * 8
        dev = params[0].device
        self.flat = torch.zeros(tot, dtype=torch.float32, device=dev)
        for p in params:
            o = self.offsets[id(p)]
            p.grad = self.flat[o:o + p.numel()].view(p.shape)
        self.params = params
        self._views = {id(p): p.grad for p in params}
        if self.world > 1:
            self._sync_parameters()
        self.bucket_elems = bucket_bytes // 4
        self.comm_stream = torch.cuda.Stream(device=dev) if (self.world > 1 and dev.type == "cuda") else None
        if self.comm_stream is not None:
            # The all-reduce kernels run beside the backward kernels: leave them their SMs (include/bubbleformer_b200.h,
            # bf_set_reserved_sms).  BF_RESERVED_SMS overrides; default = NCCL_MAX_CTAS when that is set, else 0.
            import os
            from . import _lib
            n = os.environ.get("BF_RESERVED_SMS", os.environ.get("NCCL_MAX_CTAS", "0"))
            try:
                n = min(max(int(n or 0), 0), 32)     # a large NCCL_MAX_CTAS must not take a fifth of the machine away
            except ValueError:
                n = 0
            _lib.check(_lib.lib.bf_set_reserved_sms(n), "bf_set_reserved_sms")
        self._pending: List = []
        self._lo: Optional[int] = None
        self._hi: Optional[int] = None
        model._grad_sink = self
        from . import autograd
        autograd.ACTIVE_SINK = self

    def _sync_parameters(self) -> None:
        """Rank 0's parameters become everyone's (what DistributedDataParallel does at construction, upstream
        scripts/train.py:163): averaging the gradients of replicas that started from different weights is not
        data-parallel training.  One broadcast of the flat weight bank when the model has one, else of a packed copy."""
        bank = getattr(self.model, "_bank", None)
        if bank is not None and bank.flat is not None:
            dist.broadcast(bank.flat, src=0, group=self.group)
            bank.invalidate()
            return
        with torch.no_grad():
            packed = torch.cat([p.detach().reshape(-1) for p in self.params])
            dist.broadcast(packed, src=0, group=self.group)
            o = 0
            for p in self.params:
                p.copy_(packed[o:o + p.numel()].view(p.shape))
                o += p.numel()

    def owns(self, p: torch.Tensor) -> bool:
        """Ownership is by parameter identity, never by what `p.grad` currently points at: `zero_grad(set_to_none=True)`
        (torch's default) must not silently route gradients around the flat buffer."""
        return id(p) in self.offsets

    def _attach(self, p: torch.Tensor) -> bool:
        """Make `p.grad` the parameter's view of the flat buffer again.  Returns True when it had been lost (set to
        None or replaced), in which case the caller zeroes the segment: a fresh gradient starts from zero."""
        v = self._views[id(p)]
        g = p.grad
        if g is not None and g.data_ptr() == v.data_ptr():
            return False
        p.grad = v
        return True

    def close(self) -> None:
        from . import autograd
        if autograd.ACTIVE_SINK is self:
            autograd.ACTIVE_SINK = None

    # -- per step -------------------------------------------------------------------------------
    def begin_step(self) -> None:
        """Zero the flat gradient buffer (one memset) before backward."""
        self.flat.zero_()
        self._pending.clear()
        self._lo = self._hi = None
        self._done: List = []

    def views(self, named: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Gradient views for the parameters of one autograd Function."""
        out = {}
        for n, p in named.items():
            if id(p) not in self.offsets:
                raise RuntimeError(f"GradSink: parameter {n} is not part of the model this sink was built for")
            if self._attach(p):
                self._views[id(p)].zero_()           # zero_grad(set_to_none=True) semantics: start from zero
            out[n] = self._views[id(p)]
        return out

    def _offset_of(self, p: torch.Tensor) -> int:
        return self.offsets[id(p)]

    def segment_done(self, named: Dict[str, torch.Tensor]) -> None:
        """Called when one Function's backward kernels have been enqueued: maybe launch a bucket all-reduce."""
        if self.world == 1:
            return
        lo = min(self._offset_of(p) for p in named.values())
        hi = max(self._offset_of(p) + p.numel() for p in named.values())
        self._lo = lo if self._lo is None else min(self._lo, lo)
        self._hi = hi if self._hi is None else max(self._hi, hi)
        # Buckets fill in reverse parameter order (debed first, embed last).  Near the front of the buffer, i.e. at the
        # end of backward, every segment is flushed at once so that only a small reduction is still in flight when
        # backward ends (the exposed tail of the step).
        if self._hi - self._lo >= self.bucket_elems or lo < self.bucket_elems:
            self._flush()

    def _flush(self) -> None:
        if self._lo is None:
            return
        self._done.append((self._lo, self._hi))
        self._reduce(self._lo, self._hi)
        self._lo = self._hi = None

    def _reduce(self, lo: int, hi: int) -> None:
        seg = self.flat[lo:hi]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(seg, op=dist.ReduceOp.AVG, group=self.group)
        else:                                   # gloo (CPU tests): no AVG
            dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.group)
            seg.div_(self.world)

    def finish(self) -> None:
        """Flush the last bucket, reduce whatever no Function owns (parameters differentiated by plain torch
        autograd, e.g. the FiLM MLP), and make the compute stream wait for all reductions."""
        # parameters differentiated by plain torch autograd (none in the shipped models): if their .grad was replaced by
        # a fresh tensor (zero_grad(set_to_none=True) before backward), fold it back into the flat buffer
        for p in self.params:
            g, v = p.grad, self._views[id(p)]
            if g is None:
                p.grad = v
            elif g.data_ptr() != v.data_ptr():
                v.copy_(g)
                p.grad = v
        if self.world == 1:
            return
        self._flush()
        pos = 0
        for lo, hi in sorted(self._done):
            if lo > pos:
                self._reduce(pos, lo)
            pos = max(pos, hi)
        if pos < self.flat.numel():
            self._reduce(pos, self.flat.numel())
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)


class GraphedTrainStep:
    """Forward + loss + backward (+ the bucketed gradient all-reduce) of fixed shapes, captured once in a CUDA graph.

    At config 2 an eager step issues ~600 library launches plus the torch glue around them from Python, and the GPU
    waits on the host for ~4 ms of a 33 ms step; replaying the captured step removes that.  Inputs live in static
    buffers (`x`, `tgt`, `cond`: the tensors passed at construction are adopted as those buffers); gradients land in
    the GradSink's flat buffer, so an optimizer step can follow each replay.  Stochastic-depth masks are redrawn on
    every replay (torch's graph-safe Philox offsets).  Construct it before any eager backward of the same model (or
    after every reference to earlier losses is gone): AccumulateGrad nodes kept alive from an eager step on the
    default stream would pull the capture onto that stream and invalidate it.
    """

    def __init__(self, model: torch.nn.Module, loss_fn, sink: GradSink, x: torch.Tensor, tgt: torch.Tensor,
                 cond: Optional[torch.Tensor] = None, warmup: int = 3):
        if not x.is_cuda:
            raise RuntimeError("bubbleformer_b200 runs on CUDA tensors only (no CPU fallback)")
        self.model, self.loss_fn, self.sink = model, loss_fn, sink
        self.x, self.tgt, self.cond = x.detach(), tgt.detach(), (cond.detach() if cond is not None else None)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):      # kernel attributes, allocator pools, NCCL channels: outside the capture
                self._step()
        bank = getattr(model, "_bank", None)
        if bank is not None:
            # the weight-mirror cast must be part of the graph even if an optimiser step has just refreshed the mirror:
            # replays must pick up any later change of the parameters
            bank._mirror_version = None
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(x.device)
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()
        self.launches_per_step = _lib.launch_count() - n0      # library kernels recorded in the graph

    def _step(self) -> torch.Tensor:
        self.sink.begin_step()
        y = self.model(self.x) if self.cond is None else self.model(self.x, self.cond)
        loss = self.loss_fn(y, self.tgt)
        loss.backward()
        self.sink.finish()
        return loss.detach()

    def __call__(self, x: torch.Tensor, tgt: torch.Tensor, cond: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Replays the step on (x, tgt, cond); returns the static loss tensor (clone it to keep it)."""
        for dst, src in ((self.x, x), (self.tgt, tgt), (self.cond, cond)):
            if dst is not None and src is not None and src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss
