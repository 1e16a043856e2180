"""Data-parallel training step around the SwinV2 hot path -- the stand-in for the reference's Composer loop.

The reference trains through ``composer.Trainer`` (main.py:104-131), which is not installable here.  This
module restates only what the measurement needs, following the reference's own protocol and recipe:

  * ``Model.forward(batch)`` / ``Model.loss(outputs, batch)``            reference models.py:121-152
  * multitask cross entropy ``dot(coeffs, CE_t)``                         reference hierarchy.py:65-94
  * optimizer parameter groups: no weight decay for 1-D params / biases   reference optim.py:48-58
  * DecoupledSGDW (the reference default) / SGD with nesterov momentum    reference optim.py:16-44, configs.py:44-48
    0.875, weight decay 5e-4, gradient clipping 2.0                       configs/pretrain/inat21.yaml:44-47
  * cosine annealing with linear warm-up                                  reference configs.py:52-54, main.py:62-64
  * label smoothing, also per tier of a multitask head                    reference algorithmic.py:88-119, 160-164
  * per-rank batch = global batch / world size                            reference main.py:44-48
  * uint8 images normalised on the device                                 reference data.py:130-136, 154-164

One process per GPU (``torchrun``), gradients synchronised with NCCL all-reduce through
``DistributedDataParallel`` (gloo on CPU for the host-logic tests).  Nothing here knows about kernels: the
model is whatever ``nn.Module`` it is given (``hierarchical_vision_b200.swinv2.SwinTransformerV2`` in the
product, the oracle in tests).
"""
from __future__ import annotations

import math
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


def _fused_ce_ok(logits: torch.Tensor, target: torch.Tensor) -> bool:
    return (logits.is_cuda and logits.dim() == 2 and logits.dtype in (torch.float32, torch.bfloat16)
            and target.dtype == torch.int64 and target.dim() == 1)


def hvf_cross_entropy(logits, target, label_smoothing, weight=1.0):
    from . import functional as hvf  # the kernels only exist on the device; the CPU host-logic tests never get here
    return hvf.cross_entropy(logits, target, label_smoothing, weight)


def smooth_labels(logits: torch.Tensor, target: torch.Tensor, smoothing: float = 0.1) -> torch.Tensor:
    """(B,) class indices -> (B, classes) soft targets ``onehot * (1 - smoothing) + smoothing / classes``
    (algorithmic.py:160-164, itself Composer's ``smooth_labels``)."""
    n_classes = logits.shape[1]
    onehot = F.one_hot(target, n_classes).to(torch.float32)
    return onehot * (1.0 - smoothing) + smoothing / n_classes


def multitask_cross_entropy(logits: List[torch.Tensor], targets, coeffs: Sequence[float] = MULTITASK_COEFFS,
                            label_smoothing: float = 0.0) -> torch.Tensor:
    """targets (B, tiers) class indices, or a list of per-tier (soft) targets as the reference's patched
    LabelSmoothing produces (algorithmic.py:100-112); loss = sum_t coeff_t * CE(logits_t, targets_t)
    (hierarchy.py:65-94).  ``label_smoothing`` > 0 smooths every tier like that algorithm does."""
    if not isinstance(targets, (list, tuple)):
        targets = list(targets.unbind(dim=1))
    if not (len(logits) == len(targets) == len(coeffs)):
        raise ValueError(f"{len(logits)} != {len(targets)} != {len(coeffs)}")
    total = logits[0].new_zeros((), dtype=torch.float32)
    for t, lg in enumerate(logits):
        tg = targets[t]
        if _fused_ce_ok(lg, tg):  # class-index targets on the device: loss and gradient from one kernel per tier
            total = total + hvf_cross_entropy(lg, tg, label_smoothing, coeffs[t])
            continue
        if label_smoothing > 0.0 and not tg.is_floating_point():
            tg = smooth_labels(lg, tg, label_smoothing)
        total = total + coeffs[t] * F.cross_entropy(lg.float(), tg)
    return total


class Model(nn.Module):
    """Composer-protocol wrapper (models.py:121-152): ``forward(batch) = module(batch[0])``,
    ``loss(outputs, batch)`` = cross entropy, or the multitask sum when the module returns a list.
    ``label_smoothing``: what the reference's LabelSmoothing algorithm does to the targets before the loss."""

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
            return multitask_cross_entropy(list(outputs), targets, self.coeffs, self.label_smoothing)
        if _fused_ce_ok(outputs, targets):
            return hvf_cross_entropy(outputs, targets, self.label_smoothing)
        return F.cross_entropy(outputs.float(), targets, label_smoothing=self.label_smoothing)


def cosine_warmup_factor(step: int, warmup_steps: int, total_steps: int, alpha_f: float = 0.0) -> float:
    """LR multiplier of Composer's CosineAnnealingWithWarmupScheduler(t_warmup, alpha_f) (configs.py:52-54): linear
    from 0 to 1 over the warm-up, then ``alpha_f + (1 - alpha_f) * (1 + cos(pi * frac)) / 2`` over the remaining steps.
    Composer is not vendored here; this restates its documented schedule."""
    if warmup_steps > 0 and step < warmup_steps:
        return step / warmup_steps
    span = max(total_steps - warmup_steps, 1)
    frac = min(max(step - warmup_steps, 0) / span, 1.0)
    return alpha_f + (1.0 - alpha_f) * 0.5 * (1.0 + math.cos(math.pi * frac))


class FlatSGD(torch.optim.Optimizer):
    """The reference's two SGD flavours over a CUDA-graph-friendly layout (optim.py:16-44):

      * ``decoupled=True``  Composer's DecoupledSGDW, the reference default (configs.py:45): momentum buffer
        ``buf = momentum * buf + grad``, update ``p = p * (1 - (lr / initial_lr) * weight_decay) - lr * buf`` (restated
        from Composer 0.13's documented rule; the package itself is not installable here);
      * ``decoupled=False`` ``torch.optim.SGD(..., nesterov=True)`` of the reference's "sgd" branch:
        ``g = grad + wd * p; buf = momentum * buf + g; p -= lr * (g + momentum * buf)``.

    The learning rate lives in a device tensor (``set_lr``), so a captured CUDA graph follows a schedule instead of
    baking the value of capture time into its kernels; all tensors of a group are updated by a handful of
    multi-tensor kernels.  Parameter groups follow the reference (``build_optimizer``): a group's ``weight_decay``
    0 disables decay for it."""

    def __init__(self, params, lr: float, momentum: float = 0.875, weight_decay: float = 5e-4, decoupled: bool = True,
                 nesterov: bool = True):
        defaults = dict(lr=lr, momentum=momentum, weight_decay=weight_decay, decoupled=decoupled, nesterov=nesterov,
                        initial_lr=lr)
        super().__init__(params, defaults)
        dev = self.param_groups[0]["params"][0].device
        self.lr_t = torch.tensor(float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)

    def set_lr(self, lr: float) -> None:
        """New learning rate for every group, effective at the next step (also inside a replayed CUDA graph)."""
        if lr != self._lr_host:
            self.lr_t.fill_(float(lr))
            self._lr_host = float(lr)
        for g in self.param_groups:
            g["lr"] = float(lr)

    def prepare_fused(self, flat: torch.Tensor) -> bool:
        """Build the tables of the one-pass kernel (hv_sgdw_step) for gradients that are views into ``flat`` (the layout of
        GraphedTrainStep).  Returns False (and leaves the multi-tensor path in place) for the nesterov-SGD flavour, CPU
        tensors or gradients that are not views of ``flat``."""
        import struct

        self._fused = None
        if not flat.is_cuda or flat.dtype != torch.float32 or not all(g["decoupled"] for g in self.param_groups):
            return False
        recs, chunks = [], []
        base, esz, moms = flat.data_ptr(), flat.element_size(), {g["momentum"] for g in self.param_groups}
        if len(moms) != 1:
            return False
        for g in self.param_groups:
            wd_scale = (g["weight_decay"] / g["initial_lr"]) if (g["weight_decay"] != 0.0 and g["initial_lr"] != 0.0) else 0.0
            for p in g["params"]:
                if p.grad is None or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    return False
                goff = (p.grad.data_ptr() - base) // esz
                if goff < 0 or goff + p.numel() > flat.numel() or (p.grad.data_ptr() - base) % esz:
                    return False
                st = self.state[p]
                if "momentum_buffer" not in st:
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                buf = st["momentum_buffer"]
                if not buf.is_contiguous():
                    return False
                idx = len(recs)
                recs.append(struct.pack("<QQqif", p.data_ptr(), buf.data_ptr(), goff, p.numel(), wd_scale))
                chunks.extend((idx, off) for off in range(0, p.numel(), 4096))
        dev = flat.device
        table = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(dev)
        chunk_t = torch.tensor(chunks, dtype=torch.int32).reshape(-1, 2).to(dev)
        self._fused = (table, chunk_t, len(chunks), float(next(iter(moms))), flat)
        return True

    @torch.no_grad()
    def step_fused(self, clip_coef: Optional[torch.Tensor] = None) -> None:
        """The step of every group in ONE kernel (after :meth:`prepare_fused`); ``clip_coef``: device scalar the gradients
        are multiplied by on the fly (gradient clipping), or None.  The parameters are written through raw pointers: their
        autograd version counters do not move, so callers refresh weight shadows with ``force=True`` (GraphedTrainStep does)."""
        from . import _lib
        from . import functional as hvf

        table, chunk_t, n, mom, flat = self._fused
        if self.param_groups[0]["lr"] != self._lr_host:
            self.set_lr(self.param_groups[0]["lr"])
        lib = _lib.load()
        with torch.cuda.device(flat.device):
            rc = lib.hv_sgdw_step(table.data_ptr(), chunk_t.data_ptr(), n, flat.data_ptr(), self.lr_t.data_ptr(),
                                  clip_coef.data_ptr() if clip_coef is not None else None, mom,
                                  torch.cuda.current_stream(flat.device).cuda_stream)
        _lib.check(rc, "hv_sgdw_step")
        hvf.LAUNCH_COUNT += 1

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for g in self.param_groups:
            params = [p for p in g["params"] if p.grad is not None]
            if not params:
                continue
            if g["lr"] != self._lr_host:  # a scheduler wrote param_group["lr"] directly
                self.set_lr(g["lr"])
            grads = [p.grad for p in params]
            mom, wd = g["momentum"], g["weight_decay"]
            bufs = []
            for p in params:
                st = self.state[p]
                if "momentum_buffer" not in st:
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                bufs.append(st["momentum_buffer"])
            if g["decoupled"]:
                torch._foreach_mul_(bufs, mom)
                torch._foreach_add_(bufs, grads)
                if wd != 0.0 and g["initial_lr"] != 0.0:
                    torch._foreach_mul_(params, 1.0 - self.lr_t * (wd / g["initial_lr"]))
                upd = torch._foreach_mul(bufs, -self.lr_t)
            else:
                if wd != 0.0:
                    grads = torch._foreach_add(grads, params, alpha=wd)
                torch._foreach_mul_(bufs, mom)
                torch._foreach_add_(bufs, grads)
                if g["nesterov"]:
                    grads = torch._foreach_add(grads, bufs, alpha=mom)
                else:
                    grads = bufs
                upd = torch._foreach_mul(grads, -self.lr_t)
            torch._foreach_add_(params, upd)
        return loss


def build_optimizer(model: nn.Module, lr: float = 0.1, momentum: float = 0.875, weight_decay: float = 5e-4,
                    name: str = "decoupledsgdw"):
    """The reference's optimizer factory (optim.py:5-58) for its two SGD branches: ``"decoupledsgdw"`` (default,
    configs.py:45) and ``"sgd"`` (nesterov).  Grouping as in ``set_weight_decay``: 1-D parameters, ``.bias`` and
    ``no_weight_decay()`` names get no decay.  Like the reference, the skip list comes from the object handed in
    (its Composer ``Model`` has no ``no_weight_decay``, so ``logit_scale`` and ``cpb_mlp`` weights ARE decayed there;
    pass the bare backbone to honour ``SwinTransformerV2.no_weight_decay``)."""
    skip = set(model.no_weight_decay()) if hasattr(model, "no_weight_decay") else set()
    decay, no_decay = [], []
    for pname, p in model.named_parameters():
        if not p.requires_grad:
            continue
        short = pname.removeprefix("module.")  # DistributedDataParallel's wrapper prefix
        (no_decay if (p.dim() == 1 or pname.endswith(".bias") or short in skip) else decay).append(p)
    groups = [{"params": decay}, {"params": no_decay, "weight_decay": 0.0}]
    groups = [g for g in groups if g["params"]]
    key = name.lower()
    if key not in ("decoupledsgdw", "sgd"):
        raise ValueError(f"optimizer '{name}': this training stand-in covers the reference's SGD branches (decoupledsgdw, sgd)")
    return FlatSGD(groups, lr=lr, momentum=momentum, weight_decay=weight_decay, decoupled=key == "decoupledsgdw")


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
    if device_type == "cuda":
        from . import functional as hvf
        hvf.refresh_weight_shadows()  # the optimizer moved the master weights: one multi-tensor copy for all shadows
    with torch.autocast(device_type, dtype=autocast_dtype, enabled=autocast_dtype is not None):
        outputs = model(batch)
        loss = inner.loss(outputs, batch)
    loss.backward()
    if clip_norm is not None:
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip_norm)
    optimizer.step()
    return loss.detach()


class BucketedGradSync:
    """Collects the gradients of a backward pass into a flat buffer, bucket by bucket, and (world size > 1) all-reduces
    each bucket as soon as its last gradient exists -- late layers first, overlapping with the rest of the backward
    pass (the NCCL stream runs beside the compute stream; inside a CUDA graph this becomes a fork / join of the
    captured graph).  This is what DistributedDataParallel's bucket hooks do (reference main.py trains through
    Composer's DDP), without the per-step host work: the hooks only run while the graph is being captured.

    Gradients are not accumulated into the flat buffer parameter by parameter (autograd would launch one `grad += g`
    kernel per parameter and the buffer would have to be zeroed first): every ``p.grad`` starts the backward pass as
    ``None``, autograd hands the freshly computed gradient over without a kernel, and a finished bucket is moved into
    its slice of the flat buffer by one multi-tensor copy, after which ``p.grad`` is the view into the buffer.

    ``params``: parameters in flat-buffer order.  ``bounds``: bucket boundaries as parameter indices (ascending, first 0,
    last len(params)).  ``reduce``: all-reduce(avg) the buckets over the default process group."""

    def __init__(self, params: Sequence[torch.nn.Parameter], flat: torch.Tensor, bounds: Sequence[int], reduce: bool = True):
        self.flat = flat
        self.params = list(params)
        self.bounds = list(bounds)
        self.nb = len(self.bounds) - 1
        self.reduce = reduce
        offs = [0]
        for p in self.params:
            offs.append(offs[-1] + p.numel())
        self.pviews = [flat[offs[i]:offs[i + 1]].view_as(p) for i, p in enumerate(self.params)]
        self.views = [flat[offs[self.bounds[b]]:offs[self.bounds[b + 1]]] for b in range(self.nb)]
        self.sizes = [self.bounds[b + 1] - self.bounds[b] for b in range(self.nb)]
        self._count = [0] * self.nb
        self._works = []
        self._started = [False] * self.nb
        self.enabled = True
        # NCCL averages inside the collective; gloo (CPU tests of this logic) only sums
        self._avg = reduce and dist.is_initialized() and dist.get_backend() == "nccl"
        for b in range(self.nb):
            for p in self.params[self.bounds[b]:self.bounds[b + 1]]:
                p.register_post_accumulate_grad_hook(self._make_hook(b))

    @staticmethod
    def stage_bounds(named_params: Sequence[Tuple[str, torch.nn.Parameter]], min_bucket_numel: int = 1 << 20) -> List[int]:
        """One bucket per top-level stage of the model (``module.layers.<i>`` and whatever sits before / after),
        merged upwards until a bucket holds at least ``min_bucket_numel`` elements."""
        def key(name):
            parts = name.removeprefix("module.").split(".")
            return ".".join(parts[:2]) if parts[0] == "layers" and len(parts) > 1 else parts[0]
        bounds, sizes, last = [0], [], None
        for i, (name, p) in enumerate(named_params):
            k = key(name)
            if last is not None and k != last and sum(sizes) >= min_bucket_numel:
                bounds.append(i)
                sizes = []
            sizes.append(p.numel())
            last = k
        bounds.append(len(named_params))
        return bounds

    def _make_hook(self, b: int):
        def hook(_param):
            if not self.enabled:
                return
            self._count[b] += 1
            if self._count[b] == self.sizes[b]:
                self._start(b)
        return hook

    @torch.no_grad()
    def _start(self, b: int) -> None:
        self._started[b] = True
        src, dst = [], []
        for i in range(self.bounds[b], self.bounds[b + 1]):
            p, v = self.params[i], self.pviews[i]
            g = p.grad
            if g is None:
                v.zero_()  # unused parameter: its slice must not keep the previous step's gradient
            elif g.data_ptr() != v.data_ptr():
                src.append(g)
                dst.append(v)
            p.grad = v
        if src:
            torch._foreach_copy_(dst, src)
        if self.reduce:
            op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
            self._works.append(dist.all_reduce(self.views[b], op=op, async_op=True))

    def begin(self) -> None:
        """Call before backward."""
        self._count = [0] * self.nb
        self._started = [False] * self.nb
        self._works = []
        for p in self.params:
            p.grad = None

    def finish(self) -> None:
        """Call after backward: collect (and reduce) buckets whose hooks did not all fire (unused parameters), then make
        the current stream wait for every bucket."""
        for b in range(self.nb):
            if not self._started[b]:
                self._start(b)
        for w in self._works:
            w.wait()
        self._works = []
        if self.reduce and not self._avg:
            self.flat.div_(dist.get_world_size())


class GraphedTrainStep:
    """The whole training step as ONE CUDA graph: uint8 normalisation, forward, loss, backward, gradient
    all-reduce, clipping and the SGD update are captured once and replayed, so a step costs one launch on the
    host instead of ~1000 (the eager step is launch-bound below ~200 images per GPU, and DistributedDataParallel's
    per-bucket hooks add host work on top).

    Gradients live in one flat fp32 buffer (every ``p.grad`` is a view into it, like DDP's
    ``gradient_as_bucket_view``); with world_size > 1 the only collective is the NCCL all-reduce(avg) of that buffer,
    issued per stage bucket from inside the backward pass (:class:`BucketedGradSync`) and captured in the graph.
    Inputs are copied into static buffers before each replay.  The learning rate is a device tensor of the optimizer
    (:class:`FlatSGD`): ``set_lr`` before a replay changes the step the graph takes.  Capturing does not train: the
    warm-up iterations it needs run on a snapshot of the parameters and optimizer state, which is restored.
    """

    def __init__(self, model: "Model", optimizer: torch.optim.Optimizer, env: DistEnv, example_batch, *,
                 transform: Optional[nn.Module] = None, autocast_dtype=torch.bfloat16,
                 clip_norm: Optional[float] = 2.0, warmup: int = 3, overlap_allreduce: bool = True,
                 min_bucket_numel: int = 1 << 20, fused_optimizer: bool = True):
        img, lab = example_batch
        if not img.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors")
        self.model, self.optimizer, self.env = model, optimizer, env
        self.transform, self.autocast_dtype, self.clip_norm = transform, autocast_dtype, clip_norm
        self.static_img = torch.empty_like(img)
        self.static_lab = torch.empty_like(lab)
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        params = [p for _, p in named]
        self.flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=img.device)
        off = 0
        for p in params:
            if p.dtype != torch.float32:
                raise RuntimeError("GraphedTrainStep keeps fp32 master weights")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.params = params
        # gradient collection (and, with world size > 1, the overlapped all-reduce) per stage bucket
        self.sync = BucketedGradSync(params, self.flat, BucketedGradSync.stage_bounds(named, min_bucket_numel),
                                     reduce=env.world_size > 1 and overlap_allreduce)
        self._trailing_allreduce = env.world_size > 1 and not overlap_allreduce
        # DecoupledSGDW as one pass over (p, buf, g) with the clip coefficient applied on the fly (hv_sgdw_step)
        self._fused_opt = bool(fused_optimizer and hasattr(optimizer, "prepare_fused") and optimizer.prepare_fused(self.flat))
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

    def set_lr(self, lr: float) -> None:
        """Learning rate of the next step (replayed graphs included).  Needs an optimizer with a device-side learning
        rate (:class:`FlatSGD`); a plain torch optimizer bakes its lr into the captured kernels."""
        if not hasattr(self.optimizer, "set_lr"):
            raise RuntimeError("the optimizer has no device-side learning rate: its lr at capture time is baked into "
                               "the CUDA graph; use train.build_optimizer / FlatSGD")
        self.optimizer.set_lr(lr)

    @staticmethod
    def _after_step() -> None:
        """The step has changed the master weights without moving their version counters (graph replay / one-pass optimizer
        kernel): the bf16 shadows are stale for any forward outside this class (evaluation, checkpoint export) until
        refreshed; the next step refreshes them itself."""
        from . import functional as hvf
        hvf.mark_weight_shadows_stale()

    def eager(self, img: torch.Tensor, lab: torch.Tensor) -> torch.Tensor:
        """The same step without the graph (warm-up, per-kernel instrumentation, debugging)."""
        self.static_img.copy_(img, non_blocking=True)
        self.static_lab.copy_(lab, non_blocking=True)
        loss = self._body()
        self._after_step()
        return loss

    def capture(self) -> "GraphedTrainStep":
        dev = self.static_img.device
        # the warm-up steps (allocator / NCCL / cuBLAS initialisation before capture) must not train the model
        snap_p = [p.detach().clone() for p in self.params]
        snap_o = {k: (v.clone() if torch.is_tensor(v) else v) for p in self.params
                  for k, v in ((id(p), dict(self.optimizer.state.get(p, {}))),)}
        snap_o = {pid: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()} for pid, st in snap_o.items()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            for p, q in zip(self.params, snap_p):
                p.copy_(q)
                st = self.optimizer.state.get(p, None)
                if st is None:
                    continue
                old = snap_o[id(p)]
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()  # state created by the warm-up (momentum buffer): back to its initial value
        from . import functional as hvf

        n0 = hvf.LAUNCH_COUNT
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._body()
        self.launches_per_step = hvf.LAUNCH_COUNT - n0
        return self

    def _body(self) -> torch.Tensor:
        from . import functional as hvf
        hvf.refresh_weight_shadows(force=True)  # bf16 shadows of the fp32 master weights: one multi-tensor copy per step
        x = self.transform(self.static_img) if self.transform is not None else self.static_img
        batch = (x, self.static_lab)
        with torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            outputs = self.model(batch)
            loss = self.model.loss(outputs, batch)
        self.sync.begin()
        loss.backward()
        self.sync.finish()
        if self._trailing_allreduce:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        if self._fused_opt:
            # clip coefficient on the device, applied to the gradients inside the one-pass optimizer kernel
            coef = None
            if self.clip_norm is not None:
                coef = torch.clamp(self.clip_norm / (torch.linalg.vector_norm(self.flat) + 1e-6), max=1.0)
            self.optimizer.step_fused(coef)
            return loss.detach()
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
        self._after_step()
        return self.static_loss

    def __call__(self, img: torch.Tensor, lab: torch.Tensor) -> torch.Tensor:
        """Copy the batch (device or pinned host memory) into the static buffers, replay, return the loss tensor."""
        if self.graph is None:
            self.capture()
        self.static_img.copy_(img, non_blocking=True)
        self.static_lab.copy_(lab, non_blocking=True)
        self.graph.replay()
        self._after_step()
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
