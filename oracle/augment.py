"""CPU oracle (test infrastructure, NOT product code) for the training-time augmentations of the collator -- SURVEY.md 8(f).2.

Restates, element by element (numpy float32, Python loops over segments / output frames), what the reference's collator does
to a padded batch with tensor code (src/data/collators.py:144-222):

  * ``segment_augment``  -- DataCollator.soft_segment_augmentation        (collators.py:77-136)
  * ``noise_rescale``    -- DataCollator.add_gaussian_noise_and_rescale   (collators.py:50-75)
  * ``spec_augment``     -- SpecAug on [mel || STNO repeated x2] and the pair-mean back to STNO frames (collators.py:205-210;
                           src/data/augmentations.py: time_warp :70-98, mask_along_axis :20-66, SpecAug.forward :414-434).
                           The time warp is torch.nn.functional.interpolate(mode="bicubic", align_corners=False) -- third-party
                           PyTorch (ATen/native/UpSample.h: area_pixel_compute_source_index, guard_index_and_lambda,
                           get_cubic_upsample_coefficients with A = -0.75; taps clamped to the slice) -- restated in ``_bicubic_rows``.

Every random number of the reference is drawn from torch's global CPU generator in a data-independent order, so the draws
are separated from the arithmetic: ``draw_plan`` consumes the generator exactly as DataCollator.__call__ does (same seed ->
same plan), the three functions above apply a plan.

Pinned: tests/test_oracle_golden.py replays tests/golden/augment.npz, which tests/golden/make_golden_augment.py produced by
seeding torch and calling the REFERENCE's own DataCollator methods / SpecAug (imported from /root/reference/src).

Reference behaviours kept on purpose:
  * SpecAug masks channels ``[:128]`` of the concatenated input whatever the mel size (augmentations.py:425-431): with 80 mel
    bins the four STNO channels are inside the masked range, with 128 they are not;
  * one (center, warped) pair warps the whole batch; both halves are interpolated from the unwarped input;
  * ``(1 - softness)`` is formed in double precision and rounded to float32 once (Python scalar times float32 tensor).
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional, Tuple

import numpy as np
import torch

f32 = np.float32


@dataclasses.dataclass
class AugmentConfig:
    """the collator fields that steer the augmentations (collators.py:20-28) and the SpecAug parameters (:31-48)"""
    conv_subsample_factor: int = 2
    stno_gaussian_noise_var: Optional[float] = None
    stno_gaussian_noise_prob: Optional[float] = None
    stno_segment_augment_prob: Optional[float] = 0.3
    stno_segment_change_prob: float = 0.1
    stno_min_segment_length: int = 5
    stno_max_segment_length: int = 50
    spec_aug_prob: float = 0.3
    time_warp_window: int = 5
    freq_mask_width_range: Tuple[int, int] = (0, 27)
    num_freq_mask: int = 2
    time_mask_width_ratio_range: Tuple[float, float] = (0.0, 0.05)
    num_time_mask: int = 5
    mask_channels: int = 128


@dataclasses.dataclass
class Plan:
    segments: List[Tuple[int, int, int, int, float]]   # (batch row, start, end, index into the 3 other classes, softness)
    noise_rows: Optional[np.ndarray]                    # int64 [n]
    noise: Optional[np.ndarray]                         # float32 [n, C, T], already scaled by sqrt(variance)
    warp: Optional[Tuple[int, int]]                     # (center, warped)
    freq_masks: Optional[np.ndarray]                    # int64 [B, n, 2] (pos, length)
    time_masks: Optional[np.ndarray]                    # int64 [B, n, 2]
    spec: bool = False


def draw_plan(B: int, C: int, T: int, n_mels: int, T_feat: int, cfg: AugmentConfig) -> Plan:
    """consume torch's global CPU generator in the order of DataCollator.__call__ (collators.py:184-210)"""
    plan = Plan([], None, None, None, None, None)
    p = cfg.stno_segment_augment_prob
    if p is not None and p > 0 and torch.rand(1).item() < p:
        for b in range(B):                                                    # collators.py:95-134
            pos = 0
            while pos < T:
                seg_len = torch.randint(cfg.stno_min_segment_length, cfg.stno_max_segment_length + 1, (1,)).item()
                end = min(pos + seg_len, T)
                if torch.rand(1).item() < cfg.stno_segment_change_prob:
                    which = torch.randint(0, C - 1, (1,)).item()
                    softness = torch.rand(1).item()
                    plan.segments.append((b, pos, end, which, softness))
                pos = end
    if cfg.stno_gaussian_noise_var is not None and cfg.stno_gaussian_noise_var > 0:
        n = int(B * cfg.stno_gaussian_noise_prob)                             # collators.py:53-61
        if n > 0:
            plan.noise_rows = torch.randperm(B)[:n].numpy().copy()
            plan.noise = (torch.randn((n, C, T)) * (cfg.stno_gaussian_noise_var ** 0.5)).numpy().copy()
    if torch.rand(1).item() < cfg.spec_aug_prob:
        plan.spec = True
        w = cfg.time_warp_window
        if not (T_feat - w <= w):                                             # augmentations.py:83-87
            center = torch.randint(w, T_feat - w, (1,))[0]
            warped = torch.randint(center - w, center + w, (1,))[0] + 1
            plan.warp = (int(center), int(warped))
        D = min(cfg.mask_channels, n_mels + C)

        def masks(lo, hi, dim, num):                                          # augmentations.py:38-45
            length = torch.randint(lo, hi, (B, num))
            pos = torch.randint(0, max(1, dim - int(length.max())), (B, num))
            return torch.stack([pos, length], dim=-1).numpy().copy()

        plan.freq_masks = masks(cfg.freq_mask_width_range[0], cfg.freq_mask_width_range[1], D, cfg.num_freq_mask)
        lo = max(0, math.floor(T_feat * cfg.time_mask_width_ratio_range[0]))  # augmentations.py:313-317
        hi = min(T_feat, math.floor(T_feat * cfg.time_mask_width_ratio_range[1]))
        if hi > lo:
            plan.time_masks = masks(lo, hi, T_feat, cfg.num_time_mask)
    return plan


def segment_augment(stno: np.ndarray, plan: Plan) -> np.ndarray:
    """stno float32 [B, C, T]; collators.py:106-132 for the segments the plan changes"""
    out = stno.astype(f32).copy()
    C = out.shape[1]
    for b, start, end, which, softness in plan.segments:
        seg = out[b, :, start:end]
        means = [f32(seg[c].sum(dtype=f32) / f32(end - start)) for c in range(C)]
        dominant = int(np.argmax(np.array(means, dtype=f32)))
        target = [c for c in range(C) if c != dominant][which]
        keep, soft = f32(1.0 - softness), f32(softness)
        new = np.empty_like(seg)
        for c in range(C):
            new[c] = keep * seg[c] + soft * f32(1.0 if c == target else 0.0)
        tot = new[0].copy()
        for c in range(1, C):
            tot = tot + new[c]
        out[b, :, start:end] = new / tot
    return out


def noise_rescale(stno: np.ndarray, plan: Plan) -> np.ndarray:
    """collators.py:63-75"""
    out = stno.astype(f32).copy()
    if plan.noise_rows is None:
        return out
    C = out.shape[1]
    for j, b in enumerate(plan.noise_rows):
        x = out[b] + plan.noise[j]
        lo = np.minimum(x.min(axis=0), f32(0.0))
        x = x - lo
        tot = x[0].copy()
        for c in range(1, C):
            tot = tot + x[c]
        out[b] = x / tot
    return out


def _fma(a, b, c):
    """round(a * b + c) once: the product of two float32 is exact in float64"""
    return (np.asarray(a, dtype=np.float64) * np.asarray(b, dtype=np.float64) + np.asarray(c, dtype=np.float64)).astype(f32)


def _cubic_coefficients(t: np.float32):
    """ATen get_cubic_upsample_coefficients, A = -0.75, float32 arithmetic with the multiply-adds fused (see _bicubic_rows)"""
    A = f32(-0.75)

    def conv1(x):  # ((A + 2) x - (A + 3)) x x + 1
        return _fma(f32(_fma(f32(A + f32(2)), x, -f32(A + f32(3))) * x), x, f32(1))

    def conv2(x):  # ((A x - 5 A) x + 8 A) x - 4 A
        return _fma(_fma(_fma(A, x, -f32(f32(5) * A)), x, f32(f32(8) * A)), x, -f32(f32(4) * A))

    x2 = f32(f32(1.0) - t)
    return conv2(f32(t + f32(1.0))), conv1(t), conv1(x2), conv2(f32(x2 + f32(1.0)))


def _bicubic_rows(x: np.ndarray, out_len: int) -> np.ndarray:
    """x float32 [n, L] -> [n, out_len]: 1-D bicubic resampling along the last axis (align_corners=False).

    Rounding: the expressions are ATen's; WHERE a multiply-add is fused is the choice of the compiler that built the
    reference's PyTorch.  Against the x86 build in this image, fusing the source index ``scale * (i + 0.5) - 0.5``, the
    coefficient polynomials and the 4-tap accumulation reproduces torch to <= 5e-7 on unit-variance input (unfused: 8e-6,
    because the source index near 100 carries an ulp of 7.6e-6 into the weights) -- so this restatement fuses them, the
    CUDA kernel uses the same explicit fused operations, and the goldens are compared at 2e-6 * max(1, |x|)."""
    L = x.shape[1]
    out = np.empty((x.shape[0], out_len), dtype=f32)
    scale = f32(f32(L) / f32(out_len))
    for i in range(out_len):
        real = f32(_fma(scale, f32(f32(i) + f32(0.5)), f32(-0.5)))
        idx = min(int(math.floor(real)), L - 1)
        lam = f32(min(max(f32(real - f32(idx)), f32(0.0)), f32(1.0)))
        w = _cubic_coefficients(lam)
        acc = None
        for j in range(4):
            src = x[:, max(min(idx - 1 + j, L - 1), 0)]
            acc = src * f32(w[j]) if acc is None else _fma(src, f32(w[j]), acc)
        out[:, i] = acc
    return out


def spec_augment(feats: np.ndarray, stno: np.ndarray, plan: Plan, cfg: AugmentConfig) -> Tuple[np.ndarray, np.ndarray]:
    """feats float32 [B, M, Tf], stno float32 [B, C, Ts] with Tf = factor * Ts -> (feats, stno) after SpecAug (collators.py:205-210)"""
    feats, stno = feats.astype(f32), stno.astype(f32)
    if not plan.spec:
        return feats.copy(), stno.copy()
    B, M, Tf = feats.shape
    C, k = stno.shape[1], cfg.conv_subsample_factor
    x = np.concatenate([feats, np.repeat(stno, k, axis=2)], axis=1)          # [B, M + C, Tf], channel-major like the input
    if plan.warp is not None:
        center, warped = plan.warp
        y = np.empty_like(x)
        for b in range(B):
            y[b, :, :warped] = _bicubic_rows(x[b, :, :center], warped)
            y[b, :, warped:] = _bicubic_rows(x[b, :, center:], Tf - warped)
        x = y
    D = min(cfg.mask_channels, M + C)
    for b in range(B):
        for pos, length in plan.freq_masks[b]:
            x[b, pos:min(pos + length, D), :] = 0.0
    if plan.time_masks is not None:
        for b in range(B):
            for pos, length in plan.time_masks[b]:
                x[b, :D, pos:pos + length] = 0.0
    new_stno = x[:, M:, :].reshape(B, C, Tf // k, k)
    acc = new_stno[..., 0].copy()
    for j in range(1, k):
        acc = acc + new_stno[..., j]
    return x[:, :M, :].copy(), (acc / f32(k)).astype(f32)


def augment(feats: np.ndarray, stno: np.ndarray, plan: Plan, cfg: AugmentConfig) -> Tuple[np.ndarray, np.ndarray]:
    """the training branch of DataCollator.__call__ after padding (collators.py:184-210)"""
    stno = segment_augment(stno, plan)
    stno = noise_rescale(stno, plan)
    return spec_augment(feats, stno, plan, cfg)


__all__ = ["AugmentConfig", "Plan", "draw_plan", "segment_augment", "noise_rescale", "spec_augment", "augment"]
