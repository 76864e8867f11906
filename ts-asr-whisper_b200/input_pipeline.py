"""GPU-side input pipeline: DataLoader workers ship raw audio and sample-level speaker activity; the
log-mel, the STNO mask, the batch padding and the training-time augmentations all run on the GPU.

Mirrors what the reference does per sample on the CPU in ``TS_ASR_DatasetSuperclass.cut_to_sample``
(src/data/local_datasets.py: get_features :198-214 -> WhisperFeatureExtractor with padding="longest",
pad_to_multiple_of=n_samples, return_attention_mask=True; get_stno_mask :162-184 -> _create_stno_masks :186-196) and per
batch in ``DataCollator.__call__`` (src/data/collators.py:144-222).  The sample dictionaries carry the reference's keys
(input_features [M, frames], attention_mask [frames], stno_mask [frames / 2, 4], transcript, is_long_form, language), so
either half can be swapped for the reference's own code.  Kernels: dicow_logmel, dicow_stno_mask, dicow_augment_batch --
no CPU fallback.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence

from .collators import DataCollator
from .feature_extraction import DiCoWFeatureExtractor


class DeviceInputPipeline:
    def __init__(self, feature_extractor: DiCoWFeatureExtractor, collator: DataCollator,
                 model_features_subsample_factor: int = 2):
        self.feature_extractor = feature_extractor
        self.collator = collator
        self.model_features_subsample_factor = model_features_subsample_factor

    def sample(self, audio, speaker_activity, speaker_index: int, transcript: str = "", is_long_form: bool = False,
               language: Optional[str] = None) -> Dict[str, Any]:
        """one recording / cut -> the reference's sample dictionary with GPU tensors.  ``audio``: 16 kHz mono samples;
        ``speaker_activity``: 0/1 [n_speakers, n_samples] (``cut.speakers_audio_mask``); ``speaker_index``: the target's
        row, -1 for none (local_datasets.py:176-180)."""
        fe = self.feature_extractor
        batch = fe(audio, return_tensors="pt", sampling_rate=fe.sampling_rate, return_attention_mask=True, truncation=False,
                   padding="longest", pad_to_multiple_of=fe.n_samples)
        stno = fe.stno_mask(speaker_activity, speaker_index, self.model_features_subsample_factor)
        return {"input_features": batch["input_features"][0], "attention_mask": batch["attention_mask"][0],
                "stno_mask": stno, "transcript": transcript, "is_long_form": is_long_form, "language": language}

    def __call__(self, raw: Sequence[Dict[str, Any]]):
        """raw: dictionaries with audio, speaker_activity, speaker_index and optionally transcript / is_long_form /
        language / enrollment (itself such a dictionary, for SE-DiCoW).  Returns the collated, augmented batch."""
        def one(r):
            s = self.sample(r["audio"], r["speaker_activity"], r["speaker_index"], r.get("transcript", ""),
                            r.get("is_long_form", False), r.get("language"))
            if "enrollment" in r:
                s["enrollment"] = one(r["enrollment"])
            return s
        return self.collator([one(r) for r in raw])


__all__ = ["DeviceInputPipeline"]
