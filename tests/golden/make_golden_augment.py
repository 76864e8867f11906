"""Golden augmented batches from the REFERENCE's own collator (src/data/collators.py DataCollator.__call__, :144-222, with
SpecAug from src/data/augmentations.py), run in the build container with a stub tokenizer (the labels are not under test).

Inputs are regenerated from the stored numpy seed by ``make_inputs``; the file keeps the torch seed set right before the
collator call, the collator fields, and the collator's outputs.

    python tests/golden/make_golden_augment.py   ->  tests/golden/augment.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

# name: (numpy seed, torch seed, n_mels, frame counts of the samples (mel frames, even), collator fields)
CASES = {
    "v3_all": (11, 101, 128, (240, 200, 236), dict(stno_gaussian_noise_var=0.05, stno_gaussian_noise_prob=0.7,
                                                     stno_segment_augment_prob=1.0, stno_segment_change_prob=0.5,
                                                     stno_min_segment_length=3, stno_max_segment_length=12, spec_aug_prob=1.0)),
    "mel80_all": (12, 102, 80, (200, 200), dict(stno_gaussian_noise_var=0.1, stno_gaussian_noise_prob=0.5,
                                                stno_segment_augment_prob=1.0, stno_segment_change_prob=0.3,
                                                stno_min_segment_length=5, stno_max_segment_length=50, spec_aug_prob=1.0)),
    "recipe_probs_a": (13, 102, 128, (120, 96), dict(stno_gaussian_noise_var=0.002, stno_gaussian_noise_prob=1.0,
                                                     spec_aug_prob=0.3)),
    "recipe_probs_b": (14, 117, 128, (120, 96), dict(stno_gaussian_noise_var=0.002, stno_gaussian_noise_prob=1.0,
                                                     spec_aug_prob=0.3)),
    "short_no_warp": (15, 105, 80, (10, 8), dict(stno_segment_augment_prob=0.0, spec_aug_prob=1.0)),
}


def make_inputs(np_seed: int, n_mels: int, frames):
    """seeded samples: log-mel-like features [M, Tf] and soft STNO masks [Tf / 2, 4] (rows sum to 1)"""
    rng = np.random.default_rng(np_seed)
    samples = []
    for tf in frames:
        feats = rng.standard_normal((n_mels, tf)).astype(np.float32)
        raw = rng.random((tf // 2, 4)).astype(np.float32) ** 3 + np.float32(1e-3)
        stno = (raw / raw.sum(axis=1, keepdims=True)).astype(np.float32)
        samples.append((feats, stno))
    return samples


class _Tok:
    """just enough tokenizer for DataCollator.__call__ (collators.py:151-152, 181-186)"""
    upper_cased_tokens = {}

    def __call__(self, texts, padding=None, max_length=None, return_tensors=None):
        return _Enc(len(texts))


class _Enc(dict):
    def __init__(self, n):
        super().__init__(input_ids=torch.arange(3).repeat(n, 1) + 5)
        self.attention_mask = torch.ones(n, 3, dtype=torch.long)


if __name__ == "__main__":
    sys.path.insert(0, "/root/reference/src")
    from data.collators import DataCollator  # the reference

    out = {}
    for name, (np_seed, torch_seed, n_mels, frames, fields) in CASES.items():
        col = DataCollator(feature_extractor=None, tokenizer=_Tok(), bos_token_id=0, max_length=16, **fields)
        batch_in = [{"is_long_form": False, "transcript": "x", "input_features": torch.from_numpy(f),
                     "attention_mask": torch.ones(f.shape[1], dtype=torch.long), "stno_mask": torch.from_numpy(s)}
                    for f, s in make_inputs(np_seed, n_mels, frames)]
        torch.manual_seed(torch_seed)
        batch = col(batch_in)
        out[name + "/input_features"] = batch["input_features"].numpy().astype(np.float32)
        out[name + "/stno_mask"] = batch["stno_mask"].numpy().astype(np.float32)
        print(name, out[name + "/input_features"].shape, out[name + "/stno_mask"].shape,
              float(np.abs(out[name + "/input_features"]).sum()), float(out[name + "/stno_mask"].sum()))
    np.savez_compressed(os.path.join(HERE, "augment.npz"), **out)
    print(os.path.getsize(os.path.join(HERE, "augment.npz")) / 1e3, "kB")
