"""GPU log-mel feature extractor with the call surface of HF ``WhisperFeatureExtractor`` as the reference uses it
(src/data/local_datasets.py:168,173,208-214; src/models/containers.py:54): ``feature_extractor(audio, return_tensors,
sampling_rate, return_attention_mask, truncation, padding, pad_to_multiple_of)`` -> ``BatchFeature(input_features,
attention_mask)``, attributes ``n_samples / hop_length / sampling_rate / feature_size / chunk_length / nb_max_frames /
mel_filters``.  The arithmetic (STFT, power, mel, log, floor, scale) is one call into libdicow_b200.so
(ops.logmel -> dicow_logmel); padding and the BatchFeature plumbing stay on the host like in HF
(HF:models/whisper/feature_extraction_whisper.py:189-342).  No CPU fallback."""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
import torch
from transformers.audio_utils import mel_filter_bank
from transformers.feature_extraction_utils import BatchFeature

from . import ops


class DiCoWFeatureExtractor:
    model_input_names = ["input_features"]

    def __init__(self, feature_size: int = 80, sampling_rate: int = 16000, hop_length: int = 160, chunk_length: int = 30,
                 n_fft: int = 400, padding_value: float = 0.0, device: Union[str, torch.device] = "cuda", **_):
        if n_fft != 400 or hop_length != 160:
            raise NotImplementedError("the B200 log-mel kernel implements Whisper's n_fft=400 / hop_length=160 only")
        self.feature_size = feature_size
        self.sampling_rate = sampling_rate
        self.hop_length = hop_length
        self.chunk_length = chunk_length
        self.n_fft = n_fft
        self.padding_value = padding_value
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length
        # same table as HF (feature_extraction_whisper.py:95-103): float64 numpy -> float32
        self.mel_filters = mel_filter_bank(num_frequency_bins=1 + n_fft // 2, num_mel_filters=feature_size,
                                           min_frequency=0.0, max_frequency=8000.0, sampling_rate=sampling_rate,
                                           norm="slaney", mel_scale="slaney")
        self.device = torch.device(device)
        self._filters_dev: Optional[torch.Tensor] = None

    @classmethod
    def from_pretrained(cls, name_or_path: str, **kwargs):
        """Reads preprocessor_config.json through HF when available; large-v3 family has 128 mel bins."""
        try:
            from transformers import WhisperFeatureExtractor
            hf = WhisperFeatureExtractor.from_pretrained(name_or_path)
            kwargs.setdefault("feature_size", hf.feature_size)
            kwargs.setdefault("chunk_length", hf.chunk_length)
        except Exception:  # offline without a cached config: infer from the model name
            kwargs.setdefault("feature_size", 128 if "large-v3" in str(name_or_path) else 80)
        return cls(**kwargs)

    def _filters(self, dev: torch.device) -> torch.Tensor:
        if self._filters_dev is None or self._filters_dev.device != dev:
            self._filters_dev = torch.from_numpy(np.ascontiguousarray(self.mel_filters, dtype=np.float32)).to(dev)
        return self._filters_dev

    def __call__(self, raw_speech, truncation: bool = True, pad_to_multiple_of: Optional[int] = None,
                 return_tensors: Optional[str] = None, return_attention_mask: Optional[bool] = None,
                 padding: Optional[str] = "max_length", max_length: Optional[int] = None,
                 sampling_rate: Optional[int] = None, device: Optional[Union[str, torch.device]] = None, **_):
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(f"The model corresponding to this feature extractor was trained using a sampling rate of "
                             f"{self.sampling_rate}. Make sure the provided `raw_speech` input was sampled with "
                             f"{self.sampling_rate} and not {sampling_rate}.")
        dev = torch.device(device) if device is not None else self.device
        if dev.type != "cuda":
            raise ops.DicowError("DiCoWFeatureExtractor runs on an sm_100 CUDA device only (no CPU fallback)")
        if isinstance(raw_speech, torch.Tensor):
            waves: List[torch.Tensor] = [raw_speech] if raw_speech.dim() == 1 else list(raw_speech)
        elif isinstance(raw_speech, np.ndarray):
            waves = [torch.from_numpy(raw_speech)] if raw_speech.ndim == 1 else [torch.from_numpy(w) for w in raw_speech]
        elif len(raw_speech) and isinstance(raw_speech[0], (np.ndarray, list, tuple, torch.Tensor)):
            waves = [torch.as_tensor(np.asarray(w) if not isinstance(w, torch.Tensor) else w) for w in raw_speech]
        else:
            waves = [torch.as_tensor(np.asarray(raw_speech))]
        waves = [w.to(torch.float32).flatten() for w in waves]
        lengths = [int(w.numel()) for w in waves]
        # HF SequenceFeatureExtractor.pad semantics for the two modes the reference uses
        max_length = max_length if max_length is not None else self.n_samples
        if padding == "longest":
            target = max(lengths)
            if truncation:
                target = min(target, max_length)
        else:  # "max_length"
            target = max_length
        if truncation:
            lengths = [min(n, target) for n in lengths]
        else:
            target = max(target, max(lengths))
        if pad_to_multiple_of:
            target = -(-target // pad_to_multiple_of) * pad_to_multiple_of
        if target % self.hop_length:
            # HF emits floor(target / hop) frames with the STFT's reflect padding taken around the unaligned last sample;
            # the B200 kernel frames hop-aligned audio only.  The reference recipe pads to a multiple of 30 s
            # (src/data/local_datasets.py:208-214), so this only refuses callers that would otherwise get one frame more
            # than HF (and an STNO alignment shifted by it).
            raise ValueError(f"padded audio length {target} is not a multiple of hop_length={self.hop_length}: pass "
                             f"pad_to_multiple_of (the reference uses n_samples={self.n_samples})")
        batch = torch.full((len(waves), target), self.padding_value, dtype=torch.float32)
        for i, (w, n) in enumerate(zip(waves, lengths)):
            batch[i, :n] = w[:n].cpu()
        audio = batch.to(dev, non_blocking=True)
        len_dev = torch.tensor(lengths, dtype=torch.int64).to(dev, non_blocking=True)
        want_mask = bool(return_attention_mask)
        res = ops.logmel(audio, self._filters(dev), len_dev, return_attention_mask=want_mask)
        data = {"input_features": res[0] if want_mask else res}
        if want_mask:
            data["attention_mask"] = res[1]
        if return_tensors == "np":
            data = {k: v.cpu().numpy() for k, v in data.items()}
        return BatchFeature(data)

    def stno_mask(self, speaker_activity, speaker_index: int, model_features_subsample_factor: int = 2,
                  device: Optional[Union[str, torch.device]] = None) -> torch.Tensor:
        """STNO mask of one recording on the GPU -- what the reference's dataset computes with numpy per sample
        (src/data/local_datasets.py:162-196: ``get_stno_mask`` / ``_create_stno_masks``): ``speaker_activity`` is the
        sample-level 0/1 matrix [n_speakers, n_samples] (``cut.speakers_audio_mask``), ``speaker_index`` the target's row
        (-1: none).  Returns fp32 [frames, 4] (silence, target, non-target, overlap), frames = padded samples / 320."""
        dev = torch.device(device) if device is not None else self.device
        if dev.type != "cuda":
            raise ops.DicowError("DiCoWFeatureExtractor runs on an sm_100 CUDA device only (no CPU fallback)")
        act = torch.as_tensor(np.asarray(speaker_activity) if not isinstance(speaker_activity, torch.Tensor) else speaker_activity)
        return ops.stno_mask(act.to(dev), int(speaker_index), window_samples=self.n_samples,
                             frame_samples=model_features_subsample_factor * self.hop_length)
