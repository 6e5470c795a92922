"""FusedSGD: torch.optim.SGD(momentum, nesterov=True, weight_decay per group) as ONE launch over all parameters
(csrc/sgd.cu), a drop-in for the optimizer the reference builds in trainer/trainer.py:215-217.  It is a real
torch.optim.Optimizer (param_groups / state_dict / LambdaLR work unchanged); momentum buffers are views of one flat
buffer exposed per parameter as state[p]['momentum_buffer'] like torch's SGD.  The gradients are zeroed in the same pass
(optimizer.zero_grad() becomes a no-op for the arena)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import EtbSgdChunk, ETB_EMA_CHUNK


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=0.01, momentum=0.0, weight_decay=0.0, nesterov=True):
        if not nesterov or momentum <= 0:
            raise NotImplementedError("FusedSGD implements the reference's configuration: Nesterov momentum > 0")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=nesterov, dampening=0))
        self._table = None
        self._hyper = None

    def _build(self):
        ps = [(gi, p) for gi, g in enumerate(self.param_groups) for p in g["params"] if p.requires_grad]
        for _, p in ps:
            _lib.require_cuda(p)
            if p.grad is None:
                raise RuntimeError("FusedSGD needs materialised gradients (use parallel.GradArena or run a backward first)")
            if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                raise RuntimeError("FusedSGD expects contiguous fp32 parameters and gradients")
        dev = ps[0][1].device
        total = sum(p.numel() for _, p in ps)
        old = [self.state[p].get("momentum_buffer") for _, p in ps]
        self._flat = torch.zeros(total, dtype=torch.float32, device=dev)
        chunks, o = [], 0
        for (gi, p), ob in zip(ps, old):
            buf = self._flat[o:o + p.numel()].view_as(p)
            if ob is not None:
                buf.copy_(ob)
            self.state[p]["momentum_buffer"] = buf
            for s in range(0, p.numel(), ETB_EMA_CHUNK):
                c = EtbSgdChunk()
                n = min(ETB_EMA_CHUNK, p.numel() - s)
                c.p, c.g, c.buf, c.n, c.group = p.data_ptr() + 4 * s, p.grad.data_ptr() + 4 * s, buf.data_ptr() + 4 * s, n, gi
                chunks.append(c)
            o += p.numel()
        arr = (EtbSgdChunk * len(chunks))(*chunks)
        self._table = torch.from_numpy(np.frombuffer(arr, dtype=np.uint8).copy()).to(dev)
        self._n = len(chunks)
        self._key = tuple((p.data_ptr(), p.grad.data_ptr()) for _, p in ps)
        self._ps = ps
        self._hyper = torch.zeros(4 * len(self.param_groups), dtype=torch.float32, device=dev)
        self._hyper_host = None

    def _sync_hyper(self):
        vals = []
        for g in self.param_groups:
            vals += [float(g["lr"]), float(g["momentum"]), float(g["weight_decay"]), 0.0]
        if vals != self._hyper_host:       # H2D only when the schedule / warm-up changed something
            self._hyper.copy_(torch.tensor(vals, dtype=torch.float32))
            self._hyper_host = vals

    @torch.no_grad()
    def step(self, closure=None, zero_grad=True):
        if closure is not None:
            raise NotImplementedError
        if self._table is None or self._key != tuple((p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0) for _, p in self._ps):
            self._build()
        if not torch.cuda.is_current_stream_capturing():
            self._sync_hyper()
        _lib.check(_lib.lib().etb_sgd_step(_lib.ptr(self._table), self._n, _lib.ptr(self._hyper), int(zero_grad), _lib.stream_ptr()),
                   "etb_sgd_step")

    def load_state_dict(self, state_dict):
        """torch's load_state_dict replaces state[p]['momentum_buffer'] by fresh tensors; copy them into the flat buffer the
        kernel reads (and keep the per-parameter views), so a resumed run really continues with the restored momentum
        (trainer/trainer.py:251: `self.optimizer.load_state_dict(ckpt['optimizer'])`)."""
        super().load_state_dict(state_dict)
        if self._table is not None:
            with torch.no_grad():
                o = 0
                for _, p in self._ps:
                    view = self._flat[o:o + p.numel()].view_as(p)
                    loaded = self.state[p].get("momentum_buffer")
                    if loaded is not None and loaded.data_ptr() != view.data_ptr():
                        view.copy_(loaded)
                    self.state[p]["momentum_buffer"] = view
                    o += p.numel()
        self._hyper_host = None      # force the next step to push lr / momentum / weight_decay again

    def refresh_hyper(self):
        """Push the current lr / momentum / weight_decay of the param groups to device memory (call before replaying a
        captured graph that contains step())."""
        if self._hyper is not None:
            self._sync_hyper()
