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
    assert sorted(names[id(p)] for p in decay["params"]) == ["module.head.weight"]
    assert no_decay["weight_decay"] == 0.0 and decay["weight_decay"] == 5e-4
    assert "module.fc1.weight" in {names[id(p)] for p in no_decay["params"]}  # no_weight_decay() honoured
    assert T.per_rank_batch(2048, 8) == 256
    with pytest.raises(ValueError):
        T.per_rank_batch(10, 4)
    u8 = torch.full((1, 3, 2, 2), 255, dtype=torch.uint8)
    out = T.NormalizeOnDevice()(u8)
    assert torch.allclose(out[0, :, 0, 0], (1 - torch.tensor(T.IMAGENET_MEAN)) / torch.tensor(T.IMAGENET_STD), atol=1e-5)
    logits = [torch.zeros(2, 3), torch.zeros(2, 4)]
    tgt = torch.zeros(2, 2, dtype=torch.long)
    want = 8.0 * torch.log(torch.tensor(3.0)) + 5.65 * torch.log(torch.tensor(4.0))
    assert torch.allclose(T.multitask_cross_entropy(logits, tgt), want, atol=1e-5)
