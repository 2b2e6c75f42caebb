"""Data-parallel harness for the path.

The path shards by batch sample -- the six cameras of one sample stay together because the pool
reduces over cameras (BV2:508-514) and the render samples one sample's volume from its own six
cameras (BV2:419).  There is no exchange step in the forward or in the data-gradient backward,
so ranks never talk inside the path.  The only collective is what the reference's DDP does
(base_cli.py:84,105): one all-reduce (mean) per step over a flat fp32 bucket of parameter
gradients.  The path itself owns one parameter (``density.beta``); the bucket is sized like the
parameters adjacent to the path (the two lift convs and the three 3-D heads, 479,543 floats =
1.9 MB) so the message is representative: latency-bound on NVLink 5 / NVSwitch, launched on
NCCL's own stream right after the render backward so it overlaps the lift backward.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

# mapping_along_depth 396,288 + channel_lower 73,728 + density_conv 433 + seg_conv 7,794 + rgb_conv 1,299 + beta 1
# (BV2:171-176, 189-198)
ADJACENT_PARAM_FLOATS = 396_288 + 73_728 + 433 + 7_794 + 1_299 + 1


def shard_samples(global_batch: int, world_size: int, rank: int) -> List[int]:
    """Sample b lives on rank b mod G."""
    return [b for b in range(global_batch) if b % world_size == rank]


class GradBucket:
    """Flat fp32 gradient bucket, all-reduced (mean) once per step.

    Nothing but the wait sits on the critical path after the collective: the mean is taken by NCCL itself
    (``ReduceOp.AVG``; backends without it -- gloo in the CPU tests -- get the gradients pre-scaled by 1/G while
    they are packed, before the reduce is launched), and only the packed gradients (not the whole bucket) are
    copied back."""

    def __init__(self, device, world_size: int, numel: int = ADJACENT_PARAM_FLOATS,
                 group: Optional[dist.ProcessGroup] = None):
        self.world_size = world_size
        self.group = group
        self.flat = torch.zeros(numel, dtype=torch.float32, device=device)
        self._handle = None
        self._views: List[torch.Tensor] = []
        self._targets: List[torch.Tensor] = []
        self._used = 0
        self._avg = None      # does the backend reduce with AVG?

    def _native_avg(self) -> bool:
        if self._avg is None:
            self._avg = self.world_size > 1 and dist.get_backend(self.group) == "nccl"
        return self._avg

    def pack(self, grads: Sequence[torch.Tensor], scale: float = 1.0) -> None:
        off = 0
        self._views, self._targets = [], []
        for g in grads:
            n = g.numel()
            if off + n > self.flat.numel():
                raise ValueError("GradBucket: gradients exceed the bucket")
            view = self.flat[off:off + n]
            if scale == 1.0:
                view.copy_(g.reshape(-1))
            else:
                torch.mul(g.reshape(-1), scale, out=view)
            self._views.append(view)
            self._targets.append(g)
            off += n
        self._used = off

    def allreduce_async(self, grads: Sequence[torch.Tensor]) -> None:
        """Pack + launch the all-reduce without blocking the caller's stream of work."""
        if self.world_size <= 1:
            self.pack(grads)
            return
        if self._native_avg():
            self.pack(grads)
            self._handle = dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:
            self.pack(grads, 1.0 / self.world_size)
            if self._used < self.flat.numel():
                # the rest of the bucket stands for the adjacent parameters' gradients; keep it bounded across steps
                self.flat[self._used:].mul_(1.0 / self.world_size)
            self._handle = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def wait(self) -> None:
        """Finish the all-reduce and write the reduced values back into the gradients."""
        if self._handle is not None:
            self._handle.wait()
            self._handle = None
        for view, tgt in zip(self._views, self._targets):
            tgt.copy_(view.view_as(tgt))

    def allreduce(self, grads: Sequence[torch.Tensor]) -> None:
        self.allreduce_async(grads)
        self.wait()


def train_step(mod, depth, ctx, vols, mats_prep, cotangents, bucket: GradBucket, has_bda: bool = True, plan=None):
    """One forward+backward of lift+pool+render with fixed cotangents (BASELINE configs[2]).

    Order: render forward/backward first, then launch the bucket all-reduce (beta's gradient is
    final), then the lift forward/backward underneath it, then wait."""
    from . import ops
    den, sem, feat, rgb = vols
    beta = mod.density.beta
    for t in (depth, ctx, den, sem, feat, rgb, beta):
        t.grad = None
    rend = ops.render_fwd(den, sem, rgb, feat, beta, mats_prep, None, mod.cfg_id, has_bda, 3)
    torch.autograd.backward(list(rend), list(cotangents[1:]))
    bucket.allreduce_async([beta.grad])
    vox, _ = ops.lift_pool_fwd(depth, ctx, mats_prep, mod.cfg_id, has_bda, False, True, plan)
    torch.autograd.backward([vox], [cotangents[0]])
    bucket.wait()
    return vox, rend


class GraphedTrainStep:
    """``train_step`` replayed from two CUDA graphs.

    At B = 1 per GPU the step is ~45 short launches plus the Python of two custom ops and their autograd nodes in
    1.7 ms of GPU time: one process keeps up, eight processes on one host do not (measured on 8 x B200: 2.17 ms per
    step against 1.72 ms on one GPU, with the all-reduce itself exposing 0.02 ms).  Captured once, the step is two graph
    launches around the one NCCL call:

        graph A: render forward + backward   ->   all-reduce launched (beta's gradient is final)
        graph B: lift + pool forward + backward (runs under the all-reduce)   ->   wait

    The inputs are the tensors given at construction: a producer writes the next batch (features, volumes, prepared
    matrices) INTO them before each call, exactly like any graph-captured training loop; every kernel re-reads them on
    replay, so per-step matrices (training draws a new ida every step, base_exp.py:93-111) cost nothing extra.  The
    gradients live in ``.grad`` of the same tensors and are rewritten (not accumulated) by each replay."""

    def __init__(self, mod, depth, ctx, vols, mats_prep, cotangents, bucket: GradBucket, has_bda: bool = True,
                 plan=None, warmup: int = 3):
        from . import ops
        self.bucket = bucket
        den, sem, feat, rgb = vols
        self.beta = beta = mod.density.beta
        self.leaves = (depth, ctx, den, sem, feat, rgb, beta)
        cots = list(cotangents)

        def render_part():
            rend = ops.render_fwd(den, sem, rgb, feat, beta, mats_prep, None, mod.cfg_id, has_bda, 3)
            torch.autograd.backward(list(rend), cots[1:])
            return rend

        def lift_part():
            vox, _ = ops.lift_pool_fwd(depth, ctx, mats_prep, mod.cfg_id, has_bda, False, True, plan)
            torch.autograd.backward([vox], [cots[0]])
            return vox

        def clear():
            for t in self.leaves:
                t.grad = None

        side = torch.cuda.Stream(device=depth.device)
        side.wait_stream(torch.cuda.current_stream(depth.device))
        with torch.cuda.stream(side):                    # warm-up outside capture: lazy initialisation, allocator
            for _ in range(warmup):
                clear()
                render_part()
                lift_part()
        torch.cuda.current_stream(depth.device).wait_stream(side)
        clear()
        self.graph_render, self.graph_lift = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_render):
            self.rend = render_part()
        with torch.cuda.graph(self.graph_lift, pool=self.graph_render.pool()):
            self.vox = lift_part()
        self.grads = [t.grad for t in self.leaves]       # the graphs' own gradient buffers

    def __call__(self):
        for t, g in zip(self.leaves, self.grads):        # an eager backward in between may have re-pointed .grad
            t.grad = g
        self.graph_render.replay()
        self.bucket.allreduce_async([self.beta.grad])
        self.graph_lift.replay()
        self.bucket.wait()
        return self.vox, self.rend
