"""Data parallelism of the SSOD step: one process per GPU, NCCL over NVLink, and exactly ONE collective per
step -- a SUM all-reduce of the flat fp32 gradient arena (191.8 MB for YOLOv5l-SSOD).

The reference wraps the student in DDP (trainer/trainer.py:313), multiplies both losses by WORLD_SIZE
(trainer/ssod_trainer.py:638-639,647-648) and lets DDP average the bucketed gradients: mean(W*g_r) == sum(g_r).  Here the
losses are left unscaled and the arena is summed once, after backward.  BatchNorm statistics stay per rank (SyncBN is off
in every shipped config) and the teacher EMA is updated locally from the identical post-all-reduce weights, so no other
communication exists."""
import torch


class GradArena:
    """All gradients of `params` as views of one contiguous fp32 buffer.  Every view starts on a 16-byte boundary (the
    wgrad / BN-backward kernels accumulate into the views with float4 accesses); the pad floats stay zero."""

    ALIGN = 4   # floats

    @classmethod
    def _offsets(cls, params):
        offs, o = [], 0
        for p in params:
            offs.append(o)
            o += (p.numel() + cls.ALIGN - 1) // cls.ALIGN * cls.ALIGN
        return offs, o

    def __init__(self, params, device=None):
        self.params = [p for p in params if p.requires_grad]
        self.offsets, n = self._offsets(self.params)
        dev = device if device is not None else self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o:o + p.numel()].view_as(p)

    def zero(self):
        self.flat.zero_()

    def check_views(self):
        """True while every p.grad still aliases the arena (optimizer.zero_grad(set_to_none=True) would break it)."""
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * o:
                return False
        return True

    def all_reduce_sum(self, world_size, group=None):
        if world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat
