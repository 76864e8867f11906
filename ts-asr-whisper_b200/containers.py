"""Mirror of src/models/containers.py (WhisperContainer, get_optimizer) over the B200 model classes -- the object
src/train.py / src/pretrain_encoder.py build their model, feature extractor and tokenizer from.

Same constructor arguments, attributes (.model, .feature_extractor, .tokenizer) and methods (freeze_except); differences:
  * use_flash_attention is accepted and ignored (attention always runs on the tcgen05 flash kernels);
  * use_lora attaches the reference's adapter configuration (r = 16, lora_alpha = 32, decoder q/k/v/out_proj/fc1/fc2,
    src/models/containers.py:69-78) through DiCoWForConditionalGeneration.add_lora -- peft is not needed (the kernels multiply
    by the merged weight, the backward projects onto A / B); the model is NOT wrapped in a PeftModel, parameter names keep
    their reference paths with ".lora_A" / ".lora_B" appended (lora_state_dict() exports peft's adapter keys);
  * the feature extractor is the GPU log-mel front-end with the reference's call signature;
  * ``tokenizer`` may be passed in (tests, offline use); by default WhisperTokenizerFast.from_pretrained as the reference.
"""
from __future__ import annotations

import os

import torch

from .feature_extraction import DiCoWFeatureExtractor
from .modeling_dicow import DiCoWForConditionalGeneration
from .optim import AdamW


# DiCoWConfig overrides and where the reference takes them from (src/models/containers.py:27-46); None = keep the checkpoint's
_FROM_MODEL_ARGS = ("ctc_weight", "fddt_is_diagonal", "fddt_bias_only", "fddt_use_silence", "fddt_use_target",
                    "fddt_use_overlap", "fddt_use_non_target", "apply_fddt_to_n_layers", "fddt_init", "non_target_fddt_value",
                    "use_pre_pos_fddt", "pre_ctc_sub_sample", "additional_layer", "additional_self_attention_layer",
                    "scb_layers")
_FROM_DATA_ARGS = ("use_enrollments",)


def _config_overrides(model_args, data_args, **direct) -> dict:
    picked = {f: getattr(model_args, f) for f in _FROM_MODEL_ARGS}
    picked.update({f: getattr(data_args, f) for f in _FROM_DATA_ARGS})
    picked.update(direct)
    return {k: v for k, v in picked.items() if v is not None}


def _set_trainable(model, decide) -> None:
    for name, param in model.named_parameters():
        param.requires_grad = bool(decide(name))


class WhisperContainer:
    def __init__(self, use_flash_attention=False, params_to_keep_frozen_keywords=None, remove_timestamps_from_ctc=False,
                 model_args=None, data_args=None, use_fddt=False, use_lora=False, tokenizer=None, feature_extractor=None):
        del use_flash_attention  # one attention implementation here
        name = self.model_type = model_args.whisper_model
        timestamps = data_args.use_timestamps
        self.model = DiCoWForConditionalGeneration.from_pretrained(
            name, **_config_overrides(model_args, data_args, use_fddt=use_fddt,
                                      remove_timestamps_from_ctc=remove_timestamps_from_ctc))
        self.model.post_init()
        self.feature_extractor = feature_extractor if feature_extractor is not None else DiCoWFeatureExtractor.from_pretrained(name)
        if tokenizer is None:
            from transformers.models.whisper import WhisperTokenizerFast
            tokenizer = WhisperTokenizerFast.from_pretrained(name, predict_timestamps=timestamps)
        self.tokenizer = tokenizer
        multilingual = ".en" not in name
        if multilingual:  # src/models/containers.py:57-65: the label prefix is <|lang|><|transcribe|>
            gen = self.model.generation_config
            gen.language, gen.task = None, "transcribe"
            tokenizer.set_prefix_tokens(predict_timestamps=timestamps, task="transcribe", language=data_args.global_lang_id)
        else:
            tokenizer.set_prefix_tokens(predict_timestamps=timestamps)
        self.model.set_tokenizer(tokenizer)
        self.model.config.forced_decoder_ids = None
        if use_lora:  # src/models/containers.py:69-78
            self.model.add_lora(r=16, lora_alpha=32)
        if params_to_keep_frozen_keywords is not None:  # src/models/containers.py:80-90 (adapters always train)
            frozen = tuple(params_to_keep_frozen_keywords)
            _set_trainable(self.model, lambda n: "lora_" in n or not any(k in n for k in frozen))

    def freeze_except(self, prefixes_to_preheat):
        _set_trainable(self.model, lambda n: n.startswith(tuple(prefixes_to_preheat)))


def get_optimizer(model, training_args, prefixes_with_higher_lr=None):
    """src/models/containers.py:100-114: AdamW with a second group (higher learning rate, no weight decay) for the
    parameters whose names start with one of ``prefixes_with_higher_lr``; None unless ``use_custom_optimizer``."""
    if not training_args.use_custom_optimizer:
        return None
    fast_prefixes = tuple(prefixes_with_higher_lr or ())
    groups = {False: [], True: []}
    for name, param in model.named_parameters():
        groups[bool(fast_prefixes) and name.startswith(fast_prefixes)].append(param)
    lr = training_args.learning_rate
    # the same optimizer (a torch.optim.AdamW subclass: update rule, param groups and state_dict unchanged); its step() updates
    # all fp32 CUDA parameters in one launch (optim.py).  DICOW_TORCH_ADAMW=1 returns the stock class.
    cls = torch.optim.AdamW if os.environ.get("DICOW_TORCH_ADAMW") == "1" else AdamW
    return cls(
        [dict(params=groups[False]), dict(params=groups[True], lr=training_args.fddt_lr_multiplier * lr, weight_decay=0.0)],
        lr=lr, weight_decay=training_args.weight_decay)
