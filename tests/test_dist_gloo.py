"""CPU, world_size 2, gloo: the host-side data-parallel logic of hierarchical_vision_b200.train
(sharding, DDP gradient averaging == single-process gradient on the concatenated batch, max-over-ranks
timing, optimizer grouping).  The model here is a small torch module: the kernels have no CPU path, so
the GPU side of data parallelism is exercised by bench.py under torchrun on the GPU box."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp
import torch.nn as nn

from hierarchical_vision_b200 import train as T


class TinyBackbone(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(12, 16)
        self.norm = nn.LayerNorm(16)
        self.head = nn.Linear(16, 5)

    def no_weight_decay(self):
        return {"fc1.weight"}

    def forward(self, x):
        return self.head(self.norm(torch.relu(self.fc1(x.flatten(1)))))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    env = T.init_distributed("gloo")
    torch.manual_seed(0)
    model = T.Model(TinyBackbone())
    ddp = T.wrap_ddp(model, env, torch.device("cpu"))
    opt = T.build_optimizer(ddp, lr=0.1)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 3, 2, 2, generator=g)
    y = torch.randint(0, 5, (8,), generator=g)
    lo, hi = T.shard_range(8, env.rank, env.world_size)
    loss = T.train_step(ddp, opt, (x[lo:hi], y[lo:hi]), clip_norm=None)
    grads = {k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()}  # plain arrays: no fd passing
    slow = T.max_over_ranks(1.0 + rank, env, torch.device("cpu"))
    T.barrier(env)
    q.put((rank, float(loss), grads, slow, (lo, hi)))
    torch.distributed.destroy_process_group()


def test_ddp_gloo_two_ranks_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=120) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the concatenated batch
    torch.manual_seed(0)
    model = T.Model(TinyBackbone())
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 3, 2, 2, generator=g)
    y = torch.randint(0, 5, (8,), generator=g)
    loss = model.loss(model((x, y)), (x, y))
    loss.backward()
    assert results[0][4] == (0, 4) and results[1][4] == (4, 8)
    assert results[0][3] == results[1][3] == 2.0  # max over ranks of (1.0, 2.0)
    assert abs(0.5 * (results[0][1] + results[1][1]) - float(loss)) < 1e-6
    for k, p in model.named_parameters():
        for r in results:  # DDP averages gradients: every rank holds the full-batch gradient
            assert torch.allclose(torch.from_numpy(r[2][k]), p.grad, atol=1e-6), k


def test_optimizer_groups_and_sharding_rules():
    model = T.Model(TinyBackbone())
    opt = T.build_optimizer(model, lr=0.1, momentum=0.875, weight_decay=5e-4)
    decay, no_decay = opt.param_groups
    names = {id(p): n for n, p in model.named_parameters()}
    # like the reference's Composer Model (models.py:121-152) the wrapper has no no_weight_decay(): only 1-D parameters and
    # biases escape the decay (optim.py:48-58) ...
    assert sorted(names[id(p)] for p in decay["params"]) == ["module.fc1.weight", "module.head.weight"]
    assert no_decay["weight_decay"] == 0.0 and decay["weight_decay"] == 5e-4
    # ... while the bare backbone's skip list is honoured
    bare = TinyBackbone()
    d2, nd2 = T.build_optimizer(bare, lr=0.1).param_groups
    n2 = {id(p): n for n, p in bare.named_parameters()}
    assert sorted(n2[id(p)] for p in d2["params"]) == ["head.weight"] and "fc1.weight" in {n2[id(p)] for p in nd2["params"]}
    assert T.per_rank_batch(2048, 8) == 256
    with pytest.raises(ValueError):
        T.per_rank_batch(10, 4)
    u8 = torch.full((1, 3, 2, 2), 255, dtype=torch.uint8)
    out = T.NormalizeOnDevice()(u8)
    assert torch.allclose(out[0, :, 0, 0], (1 - torch.tensor(T.IMAGENET_MEAN)) / torch.tensor(T.IMAGENET_STD), atol=1e-5)
    logits = [torch.zeros(2, 3), torch.zeros(2, 4)]
    tgt = torch.zeros(2, 2, dtype=torch.long)
    want = 8.0 * torch.log(torch.tensor(3.0)) + 5.65 * torch.log(torch.tensor(4.0))
    assert torch.allclose(T.multitask_cross_entropy(logits, tgt, (8.0, 5.65)), want, atol=1e-5)
    with pytest.raises(ValueError):  # hierarchy.py:80-82: one coefficient per tier
        T.multitask_cross_entropy(logits, tgt)


def _two_models():
    torch.manual_seed(0)
    a = TinyBackbone()
    torch.manual_seed(0)
    b = TinyBackbone()
    return a, b


def test_fused_optimizer_tables_need_cuda_gradient_views_and_shadows_can_be_marked_stale():
    """Host logic around the one-pass optimizer kernel: FlatSGD.prepare_fused declines (multi-tensor path stays) for CPU tensors
    and for the nesterov-SGD flavour; mark_weight_shadows_stale makes every registered shadow refresh at its next use."""
    from hierarchical_vision_b200 import functional as hvf

    a, _ = _two_models()
    ps = [p for p in a.parameters()]
    flat = torch.zeros(sum(p.numel() for p in ps))
    off = 0
    for p in ps:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    assert T.FlatSGD(ps, lr=0.1).prepare_fused(flat) is False                    # CPU tensors
    assert T.FlatSGD(ps, lr=0.1, decoupled=False).prepare_fused(flat) is False   # torch-SGD flavour
    w = torch.nn.Parameter(torch.randn(4, 3))
    sh = hvf.weight_shadow(w, torch.bfloat16)
    assert torch.equal(sh, w.detach().to(torch.bfloat16))
    w.data.mul_(2.0)  # a write that does not move the version counter (what a raw-pointer kernel or a graph replay does)
    assert torch.equal(hvf.weight_shadow(w, torch.bfloat16), sh)                  # stale: the version did not move
    hvf.mark_weight_shadows_stale()
    assert torch.equal(hvf.weight_shadow(w, torch.bfloat16), w.detach().to(torch.bfloat16))


def test_flat_sgd_matches_torch_nesterov_and_decoupled_rule():
    """reference optim.py:16-23 ("sgd": torch SGD with nesterov) and optim.py:37-44 (DecoupledSGDW), with the learning
    rate changed mid-run through the device-side lr (what a scheduler does, also under a CUDA graph)."""
    a, b = _two_models()
    oa = T.build_optimizer(a, lr=0.1, name="sgd")
    decay = [p for n, p in b.named_parameters() if p.dim() > 1 and n != "fc1.weight"]
    nd = [p for n, p in b.named_parameters() if p.dim() == 1 or n == "fc1.weight"]
    ob = torch.optim.SGD([{"params": decay}, {"params": nd, "weight_decay": 0.0}], lr=0.1, momentum=0.875, nesterov=True,
                         weight_decay=5e-4)
    for i in range(4):
        x = torch.randn(5, 3, 2, 2)
        for m, o in ((a, oa), (b, ob)):
            o.zero_grad()
            m(x).square().mean().backward()
            o.step()
        if i == 1:
            oa.set_lr(0.05)
            for g in ob.param_groups:
                g["lr"] = 0.05
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)
    # decoupled: p <- p (1 - lr / lr0 * wd) - lr * buf, buf <- momentum * buf + grad
    a, b = _two_models()
    oa = T.build_optimizer(a, lr=0.2, momentum=0.5, weight_decay=0.1, name="decoupledsgdw")
    bufs = {n: torch.zeros_like(p) for n, p in b.named_parameters()}
    lr = 0.2
    for i in range(3):
        x = torch.randn(5, 3, 2, 2)
        oa.zero_grad()
        a(x).square().mean().backward()
        oa.step()
        b.zero_grad()
        b(x).square().mean().backward()
        with torch.no_grad():
            for n, p in b.named_parameters():
                bufs[n] = 0.5 * bufs[n] + p.grad
                if p.dim() > 1 and n != "fc1.weight":
                    p.mul_(1 - lr / 0.2 * 0.1)
                p.add_(bufs[n], alpha=-lr)
        if i == 0:
            lr = 0.1
            oa.set_lr(lr)
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        T.build_optimizer(a, name="adamw")


def test_cosine_warmup_and_label_smoothing():
    f = [T.cosine_warmup_factor(s, 4, 20) for s in (0, 2, 4, 12, 20, 25)]
    assert f[0] == 0.0 and f[1] == 0.5 and f[2] == 1.0 and abs(f[3] - 0.5) < 1e-12 and abs(f[4]) < 1e-12 and abs(f[5]) < 1e-12
    assert abs(T.cosine_warmup_factor(20, 0, 20, alpha_f=0.1) - 0.1) < 1e-12
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(6, 9, generator=g)
    tgt = torch.randint(0, 9, (6,), generator=g)
    soft = T.smooth_labels(logits, tgt, 0.1)
    assert torch.allclose(soft.sum(1), torch.ones(6)) and abs(float(soft[0, tgt[0]]) - (0.9 + 0.1 / 9)) < 1e-6
    # algorithmic.py:88-119: smoothing applied per tier, loss on the soft targets == torch's label_smoothing
    lg = [torch.randn(6, 3, generator=g), torch.randn(6, 4, generator=g)]
    tg = torch.stack([torch.randint(0, 3, (6,), generator=g), torch.randint(0, 4, (6,), generator=g)], dim=1)
    want = 2.0 * torch.nn.functional.cross_entropy(lg[0], tg[:, 0], label_smoothing=0.1) + \
        0.5 * torch.nn.functional.cross_entropy(lg[1], tg[:, 1], label_smoothing=0.1)
    got = T.multitask_cross_entropy(lg, tg, (2.0, 0.5), label_smoothing=0.1)
    assert torch.allclose(got, want, atol=1e-6)
    m = T.Model(torch.nn.Identity(), coeffs=(2.0, 0.5), label_smoothing=0.1)
    assert torch.allclose(m.loss(lg, (None, tg)), want, atol=1e-6)


def _bucket_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    env = T.init_distributed("gloo")
    torch.manual_seed(0)
    model = TinyBackbone()
    named = list(model.named_parameters())
    params = [p for _, p in named]
    flat = torch.zeros(sum(p.numel() for p in params))
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    bounds = T.BucketedGradSync.stage_bounds(named, min_bucket_numel=1)  # one bucket per top-level module
    sync = T.BucketedGradSync(params, flat, bounds)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 3, 2, 2, generator=g)
    y = torch.randint(0, 5, (8,), generator=g)
    lo, hi = T.shard_range(8, env.rank, env.world_size)
    for _ in range(2):  # twice: the per-step bookkeeping resets
        flat.zero_()
        sync.begin()
        torch.nn.functional.cross_entropy(model(x[lo:hi]), y[lo:hi]).backward()
        sync.finish()
    q.put((rank, flat.numpy().copy(), bounds))
    T.barrier(env)
    torch.distributed.destroy_process_group()


def test_bucketed_grad_sync_two_ranks_matches_single_process():
    """The overlapped per-stage all-reduce of GraphedTrainStep (train.BucketedGradSync) leaves every rank with the
    gradient of the concatenated batch (reference main.py:44-48: per-rank batch = global / world)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=120) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    model = TinyBackbone()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 3, 2, 2, generator=g)
    y = torch.randint(0, 5, (8,), generator=g)
    torch.nn.functional.cross_entropy(model(x), y).backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert results[0][2] == [0, 2, 4, 6]  # fc1 | norm | head
    for r in results:
        assert torch.allclose(torch.from_numpy(r[1]), want, atol=1e-6)
