"""Error metrics shared by the GPU parity tests.

``max|err| / max|ref|`` over a whole tensor (the north_star bound) is blind to errors in low-magnitude rows / channels, so
every hidden-state / logits comparison also reports per-ROW figures (a row = one frame's or token's feature vector):
the largest relative RMS error of a row and the smallest cosine between a row and its reference."""
from __future__ import annotations

import torch


def rel_err(out: torch.Tensor, ref: torch.Tensor) -> float:
    out, ref = out.detach().float().cpu(), ref.detach().float().cpu()
    return ((out - ref).abs().max() / ref.abs().max().clamp(min=1e-12)).item()


def row_metrics(out: torch.Tensor, ref: torch.Tensor):
    """(max over rows of ||out - ref||_2 / ||ref||_2, min over rows of cos(out, ref)); rows = vectors along the last axis.
    Rows whose reference norm is below 1e-3 of the largest row norm are measured against that floor."""
    o = out.detach().double().cpu().reshape(-1, out.shape[-1])
    r = ref.detach().double().cpu().reshape(-1, ref.shape[-1])
    rn = r.norm(dim=1)
    floor = 1e-3 * rn.max().clamp(min=1e-30)
    rms = ((o - r).norm(dim=1) / torch.maximum(rn, floor)).max().item()
    cos = torch.nn.functional.cosine_similarity(o, r, dim=1)
    cos = torch.where(rn > floor, cos, torch.ones_like(cos)).min().item()
    return rms, cos


def assert_close(out, ref, *, tol: float, row_rms: float, row_cos: float, what: str = "") -> dict:
    e = rel_err(out, ref)
    rms, cos = row_metrics(out, ref)
    msg = f"{what}: max-norm rel err {e:.3e} (tol {tol:g}), worst row rel RMS {rms:.3e} (tol {row_rms:g}), min row cos {cos:.6f} (>= {row_cos})"
    print(msg)
    assert e < tol and rms < row_rms and cos > row_cos, msg
    return {"rel_err": e, "row_rms": rms, "row_cos": cos}
