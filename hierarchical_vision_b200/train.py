"""Data-parallel training step around the SwinV2 hot path -- the stand-in for the reference's Composer loop.

The reference trains through ``composer.Trainer`` (main.py:104-131), which is not installable here.  This
module restates only what the measurement needs, following the reference's own protocol and recipe:

  * ``Model.forward(batch)`` / ``Model.loss(outputs, batch)``            reference models.py:121-152
  * multitask cross entropy ``dot(coeffs, CE_t)``                         reference hierarchy.py:65-94
  * optimizer parameter groups: no weight decay for 1-D params / biases   reference optim.py:48-58
  * SGD with momentum 0.875, weight decay 5e-4, gradient clipping 2.0     reference configs.py:46-48,
                                                                          configs/pretrain/inat21.yaml:44-47
  * per-rank batch = global batch / world size                            reference main.py:44-48
  * uint8 images normalised on the device                                 reference data.py:130-136, 154-164

One process per GPU (``torchrun``), gradients synchronised with NCCL all-reduce through
``DistributedDataParallel`` (gloo on CPU for the host-logic tests).  Nothing here knows about kernels: the
model is whatever ``nn.Module`` it is given (``hierarchical_vision_b200.swinv2.SwinTransformerV2`` in the
product, the oracle in tests).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

MULTITASK_COEFFS = (8.0, 5.65, 4.0, 2.82, 2.0, 1.41, 1.0)  # configs/pretrain/r50_multitask_base.yaml:3
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


@dataclass
class DistEnv:
    rank: int = 0
    local_rank: int = 0
    world_size: int = 1

    @property
    def is_main(self) -> bool:
        return self.rank == 0


def dist_env() -> DistEnv:
    return DistEnv(int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
                   int(os.environ.get("WORLD_SIZE", 1)))


def init_distributed(backend: str, env: Optional[DistEnv] = None) -> DistEnv:
    env = env or dist_env()
    if env.world_size > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=env.rank, world_size=env.world_size)
    return env


def per_rank_batch(global_batch: int, world_size: int) -> int:
    """main.py:44-48: the global batch must split evenly over ranks."""
    if global_batch % world_size != 0:
        raise ValueError(f"global batch {global_batch} not divisible by world size {world_size}")
    return global_batch // world_size


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of ``n`` samples owned by ``rank`` (weak scaling keeps n/world fixed)."""
    per = per_rank_batch(n, world_size)
    return rank * per, (rank + 1) * per


class NormalizeOnDevice(nn.Module):
    """uint8 (B,3,H,W) -> float, (x/255 - mean)/std, executed on the device (data.py:130-136)."""

    def __init__(self, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD):
        super().__init__()
        self.register_buffer("mean", torch.tensor(mean).view(1, -1, 1, 1) * 255.0)
        self.register_buffer("std", torch.tensor(std).view(1, -1, 1, 1) * 255.0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return (x.float() - self.mean) / self.std


def multitask_cross_entropy(logits: List[torch.Tensor], targets: torch.Tensor,
                            coeffs: Sequence[float] = MULTITASK_COEFFS) -> torch.Tensor:
    """targets (B, tiers); loss = sum_t coeff_t * CE(logits_t, targets[:, t])  (hierarchy.py:86-94)."""
    total = logits[0].new_zeros((), dtype=torch.float32)
    for t, lg in enumerate(logits):
        total = total + coeffs[t] * F.cross_entropy(lg.float(), targets[:, t])
    return total


class Model(nn.Module):
    """Composer-protocol wrapper (models.py:121-152): ``forward(batch) = module(batch[0])``,
    ``loss(outputs, batch)`` = cross entropy, or the multitask sum when the module returns a list."""

    def __init__(self, module: nn.Module, coeffs: Sequence[float] = MULTITASK_COEFFS, label_smoothing: float = 0.0):
        super().__init__()
        self.module = module
        self.coeffs = tuple(coeffs)
        self.label_smoothing = label_smoothing

    def forward(self, batch):
        inputs, _ = batch
        return self.module(inputs)

    def loss(self, outputs, batch):
        _, targets = batch
        if isinstance(outputs, (list, tuple)):
            return multitask_cross_entropy(list(outputs), targets, self.coeffs)
        return F.cross_entropy(outputs.float(), targets, label_smoothing=self.label_smoothing)


def build_optimizer(model: nn.Module, lr: float = 0.1, momentum: float = 0.875, weight_decay: float = 5e-4):
    """SGD with the reference's grouping: 1-D parameters, ``.bias`` and ``no_weight_decay()`` names get no
    decay (optim.py:5-58)."""
    skip = set()
    inner = model.module if hasattr(model, "module") else model
    if hasattr(inner, "no_weight_decay"):
        skip = set(inner.no_weight_decay())
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        short = name.split("module.")[-1]
        (no_decay if (p.dim() == 1 or name.endswith(".bias") or short in skip) else decay).append(p)
    groups = [{"params": decay}, {"params": no_decay, "weight_decay": 0.0}]
    # one fused multi-tensor kernel per group on CUDA instead of a foreach chain of ~25 launches
    fused = all(p.is_cuda for p in decay + no_decay) and len(decay) + len(no_decay) > 0
    return torch.optim.SGD(groups, lr=lr, momentum=momentum, weight_decay=weight_decay, fused=fused)


def wrap_ddp(model: nn.Module, env: DistEnv, device: torch.device) -> nn.Module:
    if env.world_size == 1:
        return model
    from torch.nn.parallel import DistributedDataParallel as DDP

    if device.type == "cuda":
        return DDP(model, device_ids=[device.index], gradient_as_bucket_view=True, static_graph=False)
    return DDP(model)


def train_step(model: nn.Module, optimizer: torch.optim.Optimizer, batch, *, autocast_dtype=None,
               clip_norm: Optional[float] = 2.0) -> torch.Tensor:
    """One optimisation step; returns the (detached) loss tensor on the device.  ``model`` is a
    :class:`Model` or its DDP wrapper; gradient all-reduce happens inside ``backward`` (DDP buckets)."""
    inner = model.module if hasattr(model, "module") and isinstance(model.module, Model) else model
    optimizer.zero_grad(set_to_none=True)
    device_type = batch[0].device.type
    with torch.autocast(device_type, dtype=autocast_dtype, enabled=autocast_dtype is not None):
        outputs = model(batch)
        loss = inner.loss(outputs, batch)
    loss.backward()
    if clip_norm is not None:
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip_norm)
    optimizer.step()
    return loss.detach()


class GraphedTrainStep:
    """The whole training step as ONE CUDA graph: uint8 normalisation, forward, loss, backward, gradient
    all-reduce, clipping and the SGD update are captured once and replayed, so a step costs one launch on the
    host instead of ~1000 (the eager step is launch-bound below ~200 images per GPU, and DistributedDataParallel's
    per-bucket hooks add host work on top).

    Gradients live in one flat fp32 buffer (every ``p.grad`` is a view into it, like DDP's
    ``gradient_as_bucket_view``); with world_size > 1 the only collective is one NCCL all-reduce(avg) of that
    buffer, captured inside the graph (parameter-gradient sync only, nothing to overlap it with is
    lost: 141 MB over NVLink is < 2 % of the step).  Inputs are copied into static buffers before each replay.
    """

    def __init__(self, model: "Model", optimizer: torch.optim.Optimizer, env: DistEnv, example_batch, *,
                 transform: Optional[nn.Module] = None, autocast_dtype=torch.bfloat16,
                 clip_norm: Optional[float] = 2.0, warmup: int = 3):
        img, lab = example_batch
        if not img.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors")
        self.model, self.optimizer, self.env = model, optimizer, env
        self.transform, self.autocast_dtype, self.clip_norm = transform, autocast_dtype, clip_norm
        self.static_img = torch.empty_like(img)
        self.static_lab = torch.empty_like(lab)
        params = [p for p in model.parameters() if p.requires_grad]
        self.flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=img.device)
        off = 0
        for p in params:
            if p.dtype != torch.float32:
                raise RuntimeError("GraphedTrainStep keeps fp32 master weights")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.params = params
        self.launches_per_step = 0
        self.graph = None
        self.static_loss = None
        self.static_img.copy_(img)
        self.static_lab.copy_(lab)
        self._warmup = warmup
        # input pipeline: host batches are copied on a side stream into one of two device staging buffers while the
        # previous step is still running; the step then starts with a device-to-device copy into the static buffers
        self._copy_stream = torch.cuda.Stream(device=img.device)
        self._stage = [(torch.empty_like(img), torch.empty_like(lab), torch.cuda.Event()) for _ in range(2)]
        self._consumed = [torch.cuda.Event() for _ in range(2)]  # staging slot copied into the static buffers
        self._staged = 0

    def eager(self, img: torch.Tensor, lab: torch.Tensor) -> torch.Tensor:
        """The same step without the graph (warm-up, per-kernel instrumentation, debugging)."""
        self.static_img.copy_(img, non_blocking=True)
        self.static_lab.copy_(lab, non_blocking=True)
        return self._body()

    def capture(self) -> "GraphedTrainStep":
        dev = self.static_img.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import functional as hvf

        n0 = hvf.LAUNCH_COUNT
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._body()
        self.launches_per_step = hvf.LAUNCH_COUNT - n0
        return self

    def _body(self) -> torch.Tensor:
        self.flat.zero_()
        x = self.transform(self.static_img) if self.transform is not None else self.static_img
        batch = (x, self.static_lab)
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            outputs = self.model(batch)
            loss = self.model.loss(outputs, batch)
        loss.backward()
        if self.env.world_size > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        if self.clip_norm is not None:
            # torch.nn.utils.clip_grad_norm_ on the flat view: one norm, one scale
            total = torch.linalg.vector_norm(self.flat)
            self.flat.mul_(torch.clamp(self.clip_norm / (total + 1e-6), max=1.0))
        self.optimizer.step()
        return loss.detach()

    def prefetch(self, img: torch.Tensor, lab: torch.Tensor) -> None:
        """Start the host->device copy of the NEXT batch (pinned host tensors) on the copy stream; it overlaps with
        whatever the compute stream is doing.  Consume it with :meth:`step_prefetched`."""
        slot = self._staged ^ 1
        dimg, dlab, ev = self._stage[slot]
        self._copy_stream.wait_event(self._consumed[slot])  # the slot's previous content has been copied out
        with torch.cuda.stream(self._copy_stream):
            dimg.copy_(img, non_blocking=True)
            dlab.copy_(lab, non_blocking=True)
            ev.record(self._copy_stream)
        self._staged = slot

    def step_prefetched(self) -> torch.Tensor:
        """Run one step on the batch handed to the last :meth:`prefetch`."""
        if self.graph is None:
            self.capture()
        dimg, dlab, ev = self._stage[self._staged]
        cur = torch.cuda.current_stream(dimg.device)
        cur.wait_event(ev)
        self.static_img.copy_(dimg, non_blocking=True)
        self.static_lab.copy_(dlab, non_blocking=True)
        self._consumed[self._staged].record(cur)
        self.graph.replay()
        return self.static_loss

    def __call__(self, img: torch.Tensor, lab: torch.Tensor) -> torch.Tensor:
        """Copy the batch (device or pinned host memory) into the static buffers, replay, return the loss tensor."""
        if self.graph is None:
            self.capture()
        self.static_img.copy_(img, non_blocking=True)
        self.static_lab.copy_(lab, non_blocking=True)
        self.graph.replay()
        return self.static_loss


def max_over_ranks(value: float, env: DistEnv, device: torch.device) -> float:
    """Timing rule: a multi-GPU duration is the max over ranks."""
    if env.world_size == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(env: DistEnv) -> None:
    if env.world_size > 1:
        dist.barrier()
