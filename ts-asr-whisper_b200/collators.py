"""Collator with the reference's surface (src/data/collators.py:14-222, ``DataCollator``) whose tensor work runs on the
GPU: padding copies land in one pinned staging buffer and go up in one transfer, and the training-time augmentations
(soft STNO segment changes, Gaussian STNO noise, SpecAug on [mel || STNO]) are applied by libdicow_b200.so
(ops.augment_batch -> dicow_augment_batch) instead of ~8 tensor passes in a DataLoader worker.

Randomness: the reference draws every random number from torch's global CPU generator, and none of the draws depends on
the data.  ``draw_plan`` consumes the generator in exactly the reference's order, so a run seeded like the reference
augments the same samples the same way (pinned against the reference's own collator output in tests/golden/augment.npz).
The Gaussian noise itself is drawn on the host too (``torch.randn`` of [n, 4, T'] -- 24 kB per sample) for the same reason.

Labels (tokenizer call, language slot, -100 padding, upper-cased variant) are host work on a few hundred integers and are
restated from the reference as is.  No CPU fallback for the tensor part.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
from transformers import BatchFeature

from . import ops


@dataclasses.dataclass
class AugmentPlan:
    """host-drawn randomness of one collator call, as int32 / fp32 tables ready for dicow_augment_batch"""
    seg: Optional[torch.Tensor] = None          # int32 [n, 4]
    seg_soft: Optional[torch.Tensor] = None     # fp32 [n, 2]
    noise_rows: Optional[torch.Tensor] = None   # int32 [n]
    noise: Optional[torch.Tensor] = None        # fp32 [n, C, T]
    spec: bool = False
    warp: Optional[Tuple[int, int]] = None
    freq_masks: Optional[torch.Tensor] = None   # int32 [B, n, 2]
    time_masks: Optional[torch.Tensor] = None   # int32 [B, n, 2]

    def to(self, device) -> "AugmentPlan":
        mv = lambda t: None if t is None else t.to(device, non_blocking=True)  # noqa: E731
        return dataclasses.replace(self, seg=mv(self.seg), seg_soft=mv(self.seg_soft), noise_rows=mv(self.noise_rows),
                                   noise=mv(self.noise), freq_masks=mv(self.freq_masks), time_masks=mv(self.time_masks))


def _mask_table(lo: int, hi: int, dim: int, B: int, num: int) -> torch.Tensor:
    """mask_along_axis's two draws (src/data/augmentations.py:38-45): widths, then positions bounded by the widest mask"""
    length = torch.randint(lo, hi, (B, num))
    pos = torch.randint(0, max(1, dim - int(length.max())), (B, num))
    return torch.stack([pos, length], dim=-1).to(torch.int32).contiguous()


@dataclasses.dataclass
class DataCollator:
    feature_extractor: Any
    tokenizer: Any
    bos_token_id: Any
    max_length: int
    conv_subsample_factor: int = 2
    stno_gaussian_noise_var: float = None
    stno_gaussian_noise_prob: float = None
    stno_segment_augment_prob: float = 0.3
    stno_segment_change_prob: float = 0.1
    stno_min_segment_length: int = 5
    stno_max_segment_length: int = 50
    spec_aug_prob: float = 0.3
    use_enrollments: bool = False
    device: Union[str, torch.device] = "cuda"
    # SpecAug parameters the reference hard-codes in __post_init__ (collators.py:31-48)
    time_warp_window: int = 5
    freq_mask_width_range: Tuple[int, int] = (0, 27)
    num_freq_mask: int = 2
    time_mask_width_ratio_range: Tuple[float, float] = (0.0, 0.05)
    num_time_mask: int = 5
    mask_channels: int = 128  # SpecAug.forward masks x[:, :, :128] whatever the mel size (augmentations.py:425-431)

    # ------------------------------------------------------------------------------------------------------------
    def draw_plan(self, B: int, C: int, T: int, n_mels: int, T_feat: int) -> AugmentPlan:
        """the draws of the training branch of __call__ (collators.py:184-210), in order, from torch's global generator"""
        plan = AugmentPlan()
        p = self.stno_segment_augment_prob
        if p is not None and p > 0 and torch.rand(1).item() < p:
            seg, soft = [], []
            for b in range(B):                                      # soft_segment_augmentation, collators.py:95-134
                pos = 0
                while pos < T:
                    n = torch.randint(self.stno_min_segment_length, self.stno_max_segment_length + 1, (1,)).item()
                    end = min(pos + n, T)
                    if torch.rand(1).item() < self.stno_segment_change_prob:
                        which = torch.randint(0, C - 1, (1,)).item()
                        s = torch.rand(1).item()
                        seg.append((b, pos, end, which))
                        soft.append((s, 1.0 - s))                   # 1 - softness in double, rounded once below
                    pos = end
            if seg:
                plan.seg = torch.tensor(seg, dtype=torch.int32)
                plan.seg_soft = torch.tensor(soft, dtype=torch.float64).to(torch.float32)
        if self.stno_gaussian_noise_var is not None and self.stno_gaussian_noise_var > 0:
            n = int(B * self.stno_gaussian_noise_prob)              # add_gaussian_noise_and_rescale, collators.py:53-61
            if n > 0:
                plan.noise_rows = torch.randperm(B)[:n].to(torch.int32)
                plan.noise = torch.randn((n, C, T)) * (self.stno_gaussian_noise_var ** 0.5)
        if torch.rand(1).item() < self.spec_aug_prob:
            plan.spec = True
            w = self.time_warp_window
            if not (T_feat - w <= w):                               # time_warp, augmentations.py:83-87
                center = torch.randint(w, T_feat - w, (1,))[0]
                warped = torch.randint(center - w, center + w, (1,))[0] + 1
                plan.warp = (int(center), int(warped))
            D = min(self.mask_channels, n_mels + C)
            plan.freq_masks = _mask_table(self.freq_mask_width_range[0], self.freq_mask_width_range[1], D, B, self.num_freq_mask)
            lo = max(0, math.floor(T_feat * self.time_mask_width_ratio_range[0]))   # augmentations.py:313-317
            hi = min(T_feat, math.floor(T_feat * self.time_mask_width_ratio_range[1]))
            if hi > lo:
                plan.time_masks = _mask_table(lo, hi, T_feat, B, self.num_time_mask)
        return plan

    def apply_plan(self, feats: torch.Tensor, stno: torch.Tensor, plan: AugmentPlan) -> Tuple[torch.Tensor, torch.Tensor]:
        """feats fp32 [B, M, Tf], stno fp32 [B, C, Ts] on the GPU (stno is changed in place) -> augmented (feats, stno)"""
        if plan.seg is None and plan.noise_rows is None and not plan.spec:
            return feats, stno
        plan = plan.to(feats.device)
        return ops.augment_batch(stno, seg=plan.seg, seg_soft=plan.seg_soft, noise_rows=plan.noise_rows, noise=plan.noise,
                                 feats=feats, factor=self.conv_subsample_factor, spec=plan.spec, warp=plan.warp,
                                 freq_masks=plan.freq_masks, time_masks=plan.time_masks, mask_channels=self.mask_channels)

    def augment(self, batch):
        """Augment an already collated batch: ``input_features`` [B, M, Tf] / ``stno_mask`` [B, 4, Tf / 2] as the reference's
        collator returns them with its own augmentations switched off (stno_segment_augment_prob=0, stno_gaussian_noise_var
        =None, spec_aug_prob=0).  DataLoader workers keep padding / tokenising on the CPU; the training process moves the
        batch to the GPU and calls this -- same draws, same result as ``__call__`` on the samples."""
        dev = torch.device(self.device)
        feats = batch["input_features"].to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        stno = batch["stno_mask"].to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        if stno is batch["stno_mask"]:
            stno = stno.clone()  # the segment / noise steps work in place
        B, C, Ts = stno.shape
        plan = self.draw_plan(B, C, Ts, feats.shape[1], feats.shape[2])
        batch["input_features"], batch["stno_mask"] = self.apply_plan(feats, stno, plan)
        return batch

    # ------------------------------------------------------------------------------------------------------------
    def _pad_to_device(self, inputs) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """pad_sequence of features / attention masks / STNO masks (collators.py:153-163) into ONE pinned staging buffer,
        one host->device copy; padded STNO frames are silence (class 0 = 1)."""
        dev = torch.device(self.device)
        B = len(inputs)
        f0 = inputs[0]["input_features"]
        M = f0.shape[0]
        if f0.is_cuda:  # samples produced on the GPU (input_pipeline.DeviceInputPipeline): pad in place, nothing crosses PCIe
            return self._pad_on_device(inputs, f0.device)
        Tf = max(s["input_features"].shape[-1] for s in inputs)
        Ts = max(s["stno_mask"].shape[0] for s in inputs)
        C = inputs[0]["stno_mask"].shape[1]
        Ta = max(s["attention_mask"].shape[0] for s in inputs)
        n_f, n_s = B * M * Tf, B * C * Ts
        stage = torch.zeros(n_f + n_s + B * Ta, dtype=torch.float32, pin_memory=dev.type == "cuda")
        feats, stno = stage[:n_f].view(B, M, Tf), stage[n_f:n_f + n_s].view(B, C, Ts)
        att = stage[n_f + n_s:].view(B, Ta)
        for b, s in enumerate(inputs):
            f, m, a = s["input_features"], s["stno_mask"], s["attention_mask"]
            feats[b, :, :f.shape[-1]] = f
            stno[b, :, :m.shape[0]] = m.T
            stno[b, 0, m.shape[0]:] = 1.0
            att[b, :a.shape[0]] = a
        on = stage.to(dev, non_blocking=True)
        return (on[:n_f].view(B, M, Tf), on[n_f:n_f + n_s].view(B, C, Ts),
                on[n_f + n_s:].view(B, Ta).to(inputs[0]["attention_mask"].dtype))

    @staticmethod
    def _pad_on_device(inputs, dev) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        B, M, C = len(inputs), inputs[0]["input_features"].shape[0], inputs[0]["stno_mask"].shape[1]
        Tf = max(s["input_features"].shape[-1] for s in inputs)
        Ts = max(s["stno_mask"].shape[0] for s in inputs)
        Ta = max(s["attention_mask"].shape[0] for s in inputs)
        feats = torch.zeros(B, M, Tf, dtype=torch.float32, device=dev)
        stno = torch.zeros(B, C, Ts, dtype=torch.float32, device=dev)
        att = torch.zeros(B, Ta, dtype=inputs[0]["attention_mask"].dtype, device=dev)
        for b, s in enumerate(inputs):
            f, m, a = s["input_features"], s["stno_mask"], s["attention_mask"]
            feats[b, :, :f.shape[-1]] = f
            stno[b, :, :m.shape[0]] = m.T
            stno[b, 0, m.shape[0]:] = 1.0
            att[b, :a.shape[0]] = a
        return feats, stno, att

    def __call__(self, inputs: List[Dict[str, Union[List[int], torch.Tensor]]], nested: bool = False) -> BatchFeature:
        longform = [sample["is_long_form"] for sample in inputs]
        if len(set(longform)) != 1:
            raise ValueError("Some inputs are longform and some are not")
        in_longform = longform[0]
        labels = self.tokenizer([sample["transcript"] for sample in inputs], padding="longest", max_length=self.max_length,
                                return_tensors="pt")
        feats, stno, masks = self._pad_to_device(inputs)
        batch = BatchFeature({"input_features": feats, "attention_mask": masks, "stno_mask": stno})

        languages = [sample.get("language") for sample in inputs]
        if all(languages):
            langs = self.tokenizer.convert_tokens_to_ids([f"<|{lang}|>" for lang in languages])
            if in_longform:   # generation with given languages (collators.py:171-175)
                pre = self.tokenizer.prefix_tokens
                batch["forced_decoder_ids"] = torch.tensor([[pre[0], lang, pre[2]] for lang in langs])
            else:             # training: the language slot of the labels (collators.py:176-178)
                labels["input_ids"][:, 1] = torch.tensor(langs)
        elif any(languages):
            raise ValueError("Some inputs have language and some not. Please unify it if you want to condition by language.")

        ids = labels["input_ids"].masked_fill(labels.attention_mask.ne(1), -100)
        if (ids[:, 0] == self.bos_token_id).all().item():
            ids = ids[:, 1:]
        batch["labels"] = ids
        upper = getattr(self.tokenizer, "upper_cased_tokens", {})
        batch["upp_labels"] = ids.clone().apply_(lambda x: upper.get(int(x), x))

        if not in_longform and not nested:   # training-time augmentations (collators.py:184-210)
            B, C, Ts = stno.shape
            plan = self.draw_plan(B, C, Ts, feats.shape[1], feats.shape[2])
            batch["input_features"], batch["stno_mask"] = self.apply_plan(feats, stno, plan)

        if self.use_enrollments and not nested:
            enrollments = self([sample["enrollment"] for sample in inputs], nested=True)
            enrollments.pop("labels")
            enrollments.pop("upp_labels")
            batch["enrollments"] = enrollments
        return batch


@dataclasses.dataclass
class DataCollatorForPretraining(DataCollator):
    """src/data/collators.py:225-243 (CTC pre-training of the encoder, src/pretrain_encoder.py): features and attention masks
    padded to the longest sample, labels with -100 padding and the leading bos stripped; no STNO mask, no augmentation."""
    use_timestamps: bool = False

    def __call__(self, inputs: List[Dict[str, Union[List[int], torch.Tensor]]]) -> BatchFeature:  # type: ignore[override]
        dev = torch.device(self.device)
        labels = self.tokenizer([sample["transcript"] for sample in inputs], padding="longest", max_length=self.max_length,
                                return_tensors="pt")
        feats = [s["input_features"].squeeze() for s in inputs]          # [M, frames]
        masks = [s["attention_mask"].reshape(-1) for s in inputs]
        B, M = len(inputs), feats[0].shape[0]
        Tf, Ta = max(f.shape[-1] for f in feats), max(m.shape[0] for m in masks)
        stage = torch.zeros(B * M * Tf + B * Ta, dtype=torch.float32, pin_memory=dev.type == "cuda")
        fv, mv = stage[:B * M * Tf].view(B, M, Tf), stage[B * M * Tf:].view(B, Ta)
        for b, (f, m) in enumerate(zip(feats, masks)):
            fv[b, :, :f.shape[-1]] = f
            mv[b, :m.shape[0]] = m
        on = stage.to(dev, non_blocking=True)
        batch = BatchFeature({"input_features": on[:B * M * Tf].view(B, M, Tf),
                              "attention_mask": on[B * M * Tf:].view(B, Ta).to(masks[0].dtype)})
        ids = labels["input_ids"].masked_fill(labels.attention_mask.ne(1), -100)
        if (ids[:, 0] == self.bos_token_id).all().item():
            ids = ids[:, 1:]
        batch["labels"] = ids
        return batch


__all__ = ["AugmentPlan", "DataCollator", "DataCollatorForPretraining"]
