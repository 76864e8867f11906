"""AdamW on the C-ABI multi-tensor kernel (csrc/optimizer.cu).

The reference's ``get_optimizer`` (src/models/containers.py:100-114) returns ``torch.optim.AdamW`` over two parameter groups;
:class:`AdamW` here is that optimizer -- same constructor arguments, same update rule, same ``state_dict`` layout (``step``,
``exp_avg``, ``exp_avg_sq`` per parameter, so checkpoints move both ways) -- whose ``step()`` updates every fp32 CUDA parameter of
the model in ONE launch instead of torch's 43 multi-tensor launches for the 546 tensors of the fine-tune step.  Parameters
that are not dense fp32 CUDA tensors (none in the recipes) go through ``torch.optim.AdamW``'s own functional path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List

import numpy as np
import torch

from . import lib as _lib
from . import ops

_TABLE_DTYPE = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8"), ("lr", "<f4"), ("wd", "<f4"),
                         ("bc1", "<f4"), ("bc2s", "<f4")])
assert _TABLE_DTYPE.itemsize == C.sizeof(_lib.AdamwTensorArgs)


class AdamW(torch.optim.AdamW):
    """``torch.optim.AdamW`` with a single-launch ``step()`` for fp32 CUDA parameters (amsgrad / maximize / capturable /
    differentiable are refused: the recipes do not use them)."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2, **kw):
        for k in ("amsgrad", "maximize", "capturable", "differentiable"):
            if kw.get(k):
                raise NotImplementedError(f"AdamW({k}=True) is not implemented by the B200 optimizer kernel")
        kw.pop("fused", None), kw.pop("foreach", None)
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, foreach=False, fused=False,
                         **{k: v for k, v in kw.items() if k not in ("amsgrad", "maximize", "capturable", "differentiable")})
        self._chunk_cache: Dict[tuple, torch.Tensor] = {}
        self._table_cache: Dict[tuple, tuple] = {}

    def _chunks(self, sizes: tuple, dev: torch.device) -> torch.Tensor:
        key = (sizes, dev.index)
        t = self._chunk_cache.get(key)
        if t is None:
            ce = _lib.load_library().dicow_adamw_chunk_elems()
            idx = []
            for i, n in enumerate(sizes):
                idx.extend((i, c) for c in range(-(-n // ce)))
            t = torch.tensor(idx, dtype=torch.int32).reshape(-1, 2).contiguous().to(dev)
            self._chunk_cache = {key: t}  # one entry: the parameter list of a run does not change
        return t

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        by_dev: Dict[torch.device, List[tuple]] = {}
        others = []
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                fast = (p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32 and not p.grad.is_sparse
                        and p.is_contiguous() and p.grad.is_contiguous() and st["exp_avg"].is_contiguous())
                if not fast:
                    others.append((group, p))
                    continue
                st["step"] += 1
                k = float(st["step"])
                by_dev.setdefault(p.device, []).append((p, p.grad, st["exp_avg"], st["exp_avg_sq"], group["lr"], group["weight_decay"],
                                                        1.0 - b1 ** k, math.sqrt(1.0 - b2 ** k), b1, b2, group["eps"]))
        for dev, items in by_dev.items():
            # one launch per distinct (beta1, beta2, eps): the recipes use one setting for both groups
            settings: Dict[tuple, list] = {}
            for it in items:
                settings.setdefault(it[8:], []).append(it)
            for (b1, b2, eps), its in settings.items():
                # the p / m / v pointers and sizes of a run do not change: template cached, per step only g and the scalars
                key = (dev.index, b1, b2, eps, len(its), its[0][0].data_ptr(), its[-1][0].data_ptr())
                tpl = self._table_cache.get(key)
                if tpl is None or any(t is not it[0] for t, it in zip(tpl[1], its)):
                    tab = np.empty(len(its), dtype=_TABLE_DTYPE)
                    tab["p"] = [it[0].data_ptr() for it in its]
                    tab["m"] = [it[2].data_ptr() for it in its]
                    tab["v"] = [it[3].data_ptr() for it in its]
                    tab["n"] = [it[0].numel() for it in its]
                    tpl = (tab, [it[0] for it in its])
                    self._table_cache = {key: tpl}
                tab = tpl[0]
                tab["g"] = [it[1].data_ptr() for it in its]
                tab["lr"] = [it[4] for it in its]
                tab["wd"] = [it[5] for it in its]
                tab["bc1"] = [it[6] for it in its]
                tab["bc2s"] = [it[7] for it in its]
                table = torch.from_numpy(tab.view(np.uint8).copy()).to(dev, non_blocking=True)
                chunks = self._chunks(tuple(int(p.numel()) for p, *_ in its), dev)
                h = _lib.handle(dev.index or 0)
                with torch.cuda.device(dev):
                    rc = _lib.load_library().dicow_adamw_step(h, table.data_ptr(), chunks.data_ptr(), chunks.shape[0],
                                                              float(b1), float(b2), float(eps), ops._stream(dev))
                _lib.check(rc, h, "dicow_adamw_step")
                ops.launch_count += 1
                table.record_stream(torch.cuda.current_stream(dev))
        if others:  # not dense fp32 CUDA tensors: torch's own single-tensor path
            from torch.optim.adamw import adamw as _functional
            for group, p in others:
                st = self.state[p]
                b1, b2 = group["betas"]
                _functional([p], [p.grad], [st["exp_avg"]], [st["exp_avg_sq"]], [], [st["step"]], amsgrad=False, beta1=b1,
                            beta2=b2, lr=group["lr"], weight_decay=group["weight_decay"], eps=group["eps"], maximize=False,
                            foreach=False, capturable=False, differentiable=False, fused=False)
        return loss
