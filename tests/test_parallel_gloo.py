"""N > 1 host-side logic on CPU: world_size-2 gloo process group (rendezvous on 127.0.0.1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from ts_asr_whisper_b200 import parallel
    parallel.init_process_group("gloo")
    assert dist.get_world_size() == world
    # utterance sharding: the shards partition the batch
    mine = list(parallel.shard_range(11, rank, world))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # timing reduction = slowest rank
    mx = parallel.max_over_ranks([10.0 + rank, 5.0 - rank])
    sm = parallel.sum_over_ranks([float(len(mine))])
    # gradient exchange: mean over ranks, several buckets
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(7, 3)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2))]
    params[2].requires_grad_(False)
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    n = parallel.allreduce_gradients(params, bucket_bytes=32)
    # overlapped exchange: per-layer flat buckets reported while "the backward" is still running, waited for at the end
    ex = parallel.GradientExchange()
    flats = [torch.full((40,), float(rank + 1) * (k + 1)) for k in range(3)]
    for fl in flats:
        ex.bucket_ready(fl)
    ex.finish()
    assert ex.active and ex.n_collectives == 3 and ex.bytes == 3 * 40 * 4
    for k, fl in enumerate(flats):
        assert torch.allclose(fl, torch.full((40,), 1.5 * (k + 1)))
    parallel.barrier()
    q.put((rank, gathered, mx, sm, [p.grad.clone() for p in params], n))
    dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gathered, mx, sm, grads, n in res:
        assert sorted(sum(gathered, [])) == list(range(11)) and abs(len(gathered[0]) - len(gathered[1])) <= 1
        assert mx == [11.0, 5.0] and sm == [11.0]
        assert n >= 2  # 84-byte and 20-byte gradients with a 32-byte bucket: two collectives
        assert torch.allclose(grads[0], torch.full((7, 3), 1.5)) and torch.allclose(grads[1], torch.full((5,), 3.0))
        assert torch.allclose(grads[2], torch.full((2, 2), 3.0 * (rank + 1)))  # frozen parameter: left alone


def _ddp_worker(rank: int, world: int, port: int, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import dataclasses

    from oracle import synth
    from ts_asr_whisper_b200 import parallel, training
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=False, scb_layers=0)
    torch.manual_seed(100 + rank)  # DIFFERENT initial weights per rank: the broadcast has to repair that
    model = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
    assert model._ddp_params_and_buffers_to_ignore == []  # no process group yet: nothing changes
    parallel.init_process_group("gloo")
    for n, p in model.named_parameters():
        p.requires_grad_("fddt" in n)  # the recipe's warm-up phase: FDDT tables only
    ddp = torch.nn.parallel.DistributedDataParallel(model)  # what accelerate does for HF Trainer
    names = [n for n, _ in model.named_parameters()]
    managed = [n for n in names if n not in ddp.parameters_to_ignore]
    ex = training.gradient_exchange
    before = model.model.decoder.layers[0].fc1.weight.detach().clone()
    ex.ensure_synced(model)
    ex.ensure_synced(model)  # idempotent
    after = model.model.decoder.layers[0].fc1.weight.detach().clone()
    gathered = [None] * world
    dist.all_gather_object(gathered, after.sum().item())
    parallel.barrier()
    q.put((rank, managed, isinstance(ex, parallel.GradientExchange) and ex.active, bool((before != after).any()), gathered,
           [n for n, p in model.named_parameters() if p.requires_grad and n in managed]))
    dist.destroy_process_group()


def test_ddp_wrapper_leaves_the_gradients_to_the_overlapped_exchange():
    """torch DDP around the model (HF Trainer / accelerate, the reference's launch) is told to ignore every parameter but
    one small trainable one; the GradientExchange is installed for the hand-scheduled backward, and the parameters DDP no
    longer broadcasts are synchronised from rank 0 by the exchange itself"""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, managed, active, changed, sums, managed_trainable in res:
        assert len(managed) == 1 and "fddt" in managed[0] and managed_trainable == managed
        assert active
        assert changed == (rank != 0)        # rank 1 had different weights before the broadcast
        assert sums[0] == sums[1]            # ... and the same afterwards


def test_shard_range_partitions():
    from ts_asr_whisper_b200.parallel import shard_range
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            parts = [list(shard_range(n, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_ensure_synced_skips_host_tables_under_nccl(monkeypatch):
    """NCCL cannot broadcast host memory: a host-resident constant buffer (the soft-label smoothing table of a tokenizer attached
    after the model was moved to the GPU) is skipped, everything else is broadcast once.  (bench.py under torchrun failed on
    exactly this before the check existed.)"""
    import torch
    import torch.distributed as dist
    from ts_asr_whisper_b200 import parallel

    class FakeCuda(torch.Tensor):
        is_cuda = True  # stands in for a device tensor on this CPU-only box

    m = torch.nn.Module()
    m.w = torch.nn.Parameter(torch.zeros(3).as_subclass(FakeCuda))
    m.register_buffer("table", torch.ones(4))  # host-resident constant
    sent = []
    monkeypatch.setattr(dist, "get_backend", lambda group=None: "nccl")
    monkeypatch.setattr(dist, "broadcast", lambda t, src=0, group=None: sent.append(tuple(t.shape)))
    ex = parallel.GradientExchange()
    ex._active = True
    ex.ensure_synced(m)
    ex.ensure_synced(m)
    assert sent == [(3,)]
    monkeypatch.setattr(dist, "get_backend", lambda group=None: "gloo")
    ex2 = parallel.GradientExchange()
    ex2._active = True
    ex2.ensure_synced(m)
    assert sent == [(3,), (3,), (4,)]
