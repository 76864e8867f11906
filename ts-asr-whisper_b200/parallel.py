"""Data-parallel plumbing over torch.distributed (one process per GPU, NCCL on the GPU box, gloo in CPU tests).

The hot path shards over utterances (30 s windows are independent through mel, encoder, decoder and loss: SURVEY.md
section 8e): forward / decode run N replicas with NO data-path collective; training has one exchange step, the gradient
all-reduce (reference: torch DDP via accelerate, scripts/submit_slurm.sh:34, configs/base.yaml:73)."""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def rank_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend: str | None = None, device: torch.device | None = None) -> None:
    """Join the job's process group when launched under torchrun (WORLD_SIZE > 1); no-op otherwise."""
    _, world, _ = rank_world()
    if world <= 1 or dist.is_initialized():
        return
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
    dist.init_process_group(backend, **kw)


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of ``n_items`` utterances for ``rank`` (sizes differ by at most one; the shards
    partition range(n_items)).  Matches what a DistributedSampler-style split gives without padding duplicates."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def max_over_ranks(values: Sequence[float], device: torch.device | str = "cpu") -> List[float]:
    """Element-wise MAX of per-rank timings (ms): multi-GPU numbers are the slowest rank's."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def sum_over_ranks(values: Sequence[float], device: torch.device | str = "cpu") -> List[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20) -> int:
    """Average the gradients of ``params`` over the data-parallel group (the one exchange step of training).
    Gradients are packed into flat buckets of about ``bucket_bytes`` (sized for launch latency over NVLink/NVSwitch, not
    link count) and all-reduced with SUM then divided by the world size, like torch DDP.  Every parameter with
    requires_grad takes part -- the reference's "unfreeze after DDP wrap" hazard is not
    reproduced.  Returns the number of collectives issued."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    grads = [p.grad for p in params if p.requires_grad and p.grad is not None]
    n_coll, bucket, size = 0, [], 0

    def flush():
        nonlocal n_coll, bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_coll += 1
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return n_coll


class GradientExchange:
    """The one exchange step of training, overlapped with the backward pass.

    The hand-scheduled backward (training.py) finishes the gradients of one encoder layer at a time and reports each
    finished group through ``bucket_ready`` as ONE flat fp32 buffer (the parameters' gradients are views into it, so
    nothing is packed or copied).  The buffer is all-reduced (AVG) on a dedicated communication stream that waits on an
    event recorded on the compute stream, i.e. the NCCL kernels of layer i run over NVLink/NVSwitch while the tensor cores
    work on layer i-1; ``finish`` makes the compute stream wait for the last collective before the gradients are handed to
    autograd / the optimizer.  Bucket = one layer (79 MB for large-v3-turbo): far above NCCL's latency-bound regime, small
    enough that the last, exposed bucket costs ~0.2 ms.  Single-process runs (world size 1) make every call a no-op.

    Reference: torch DDP's bucketed all-reduce driven by autograd hooks (accelerate -> DistributedDataParallel,
    scripts/submit_slurm.sh:34, configs/base.yaml:73 ddp_find_unused_parameters)."""

    def __init__(self, group=None):
        self.group = group
        self._active: Optional[bool] = None  # decided at first use: the process group may be created after this object
        self.stream: Optional[torch.cuda.Stream] = None
        self.n_collectives = 0
        self.bytes = 0
        self._pending: List = []
        self._synced: set = set()

    @property
    def active(self) -> bool:
        if self._active is None:
            if not (dist.is_available() and dist.is_initialized()):
                return False  # not decided yet
            self._active = dist.get_world_size(self.group) > 1
        return self._active

    def ensure_synced(self, module: torch.nn.Module) -> None:
        """Once per module: broadcast every parameter and buffer from rank 0.  A DistributedDataParallel wrapper does this
        for the tensors it manages; the parameters this exchange takes off its hands (ddp_ignore_list) are not among them."""
        if not self.active or id(module) in self._synced:
            return
        self._synced.add(id(module))
        nccl = dist.get_backend(self.group) == "nccl"
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                if nccl and not t.is_cuda:
                    # host-resident constant tables (the soft-label smoothing weights of a tokenizer attached after the model
                    # was moved): derived identically on every rank, and NCCL cannot move host memory
                    continue
                dist.broadcast(t.data, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                               group=self.group)

    def bucket_ready(self, flat: torch.Tensor) -> None:
        if not self.active or flat.numel() == 0:
            return
        if flat.is_cuda:
            if self.stream is None:
                self.stream = torch.cuda.Stream(device=flat.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(flat.device))
            self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
            flat.record_stream(self.stream)
        else:  # gloo (CPU tests): no AVG, no streams
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((work, flat))
        self.n_collectives += 1
        self.bytes += flat.numel() * flat.element_size()

    def finish(self) -> None:
        if not self.active:
            return
        if self.stream is not None:
            torch.cuda.current_stream(self.stream.device).wait_stream(self.stream)
        world = dist.get_world_size(self.group)
        for work, flat in self._pending:
            work.wait()
            flat.div_(world)
        self._pending = []


def ddp_overlap_enabled() -> bool:
    return os.environ.get("DICOW_DDP_OVERLAP", "1") != "0"


def ddp_ignore_list(model: torch.nn.Module) -> List[str]:
    """What ``DiCoWForConditionalGeneration._ddp_params_and_buffers_to_ignore`` answers when torch
    DistributedDataParallel (HF Trainer -> accelerate, the reference's launch: scripts/submit_slurm.sh:34) wraps the model.

    The training step is ONE autograd node (training.DiCoWTrainStepFn): all gradients reach autograd together at the end of
    its backward, so DDP's bucketed all-reduce would start only then -- 2.9 GB of exposed communication per fine-tune step.
    Instead the hand-scheduled backward hands each finished layer's flat gradient bucket to a GradientExchange while the
    earlier layers are still being differentiated (installed here as ``training.gradient_exchange``), and DDP is told to
    leave those parameters alone.  DDP insists on managing at least one trainable parameter: the smallest trainable one
    stays with it (it is averaged twice -- the same value).  Single-process runs and DICOW_DDP_OVERLAP=0 return [] and
    change nothing."""
    if not ddp_overlap_enabled() or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() <= 1:
        return []
    from . import training
    if training.gradient_exchange is None:
        training.gradient_exchange = GradientExchange()
    named = list(model.named_parameters())
    trainable = [(p.numel(), n) for n, p in named if p.requires_grad]
    if not trainable:
        return []
    keep = min(trainable)[1]
    return [n for n, _ in named if n != keep]
