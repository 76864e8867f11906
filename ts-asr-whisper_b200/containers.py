"""Mirror of src/models/containers.py (WhisperContainer, get_optimizer) over the B200 model classes -- the object
src/train.py / src/pretrain_encoder.py build their model, feature extractor and tokenizer from.

Same constructor arguments, attributes (.model, .feature_extractor, .tokenizer) and methods (freeze_except); differences:
  * use_flash_attention is accepted and ignored (attention always runs on the tcgen05 flash kernels);
  * use_lora raises: the decoder's projections are parameter containers driven by the C-ABI kernels, not nn.Linear
    forwards peft could wrap (not built);
  * the feature extractor is the GPU log-mel front-end with the reference's call signature;
  * ``tokenizer`` may be passed in (tests, offline use); by default WhisperTokenizerFast.from_pretrained as the reference.
"""
from __future__ import annotations

import torch

from .feature_extraction import DiCoWFeatureExtractor
from .modeling_dicow import DiCoWForConditionalGeneration


class WhisperContainer:
    def __init__(self, use_flash_attention=False, params_to_keep_frozen_keywords=None, remove_timestamps_from_ctc=False,
                 model_args=None, data_args=None, use_fddt=False, use_lora=False, tokenizer=None, feature_extractor=None):
        del use_flash_attention
        self.model_type = model_args.whisper_model
        predict_timestamps = data_args.use_timestamps
        global_lang_id = data_args.global_lang_id
        overwrite_args = {  # src/models/containers.py:27-46
            "ctc_weight": model_args.ctc_weight,
            "fddt_is_diagonal": model_args.fddt_is_diagonal,
            "fddt_bias_only": model_args.fddt_bias_only,
            "fddt_use_silence": model_args.fddt_use_silence,
            "fddt_use_target": model_args.fddt_use_target,
            "fddt_use_overlap": model_args.fddt_use_overlap,
            "fddt_use_non_target": model_args.fddt_use_non_target,
            "remove_timestamps_from_ctc": remove_timestamps_from_ctc,
            "apply_fddt_to_n_layers": model_args.apply_fddt_to_n_layers,
            "use_fddt": use_fddt,
            "fddt_init": model_args.fddt_init,
            "non_target_fddt_value": model_args.non_target_fddt_value,
            "use_pre_pos_fddt": model_args.use_pre_pos_fddt,
            "use_enrollments": data_args.use_enrollments,
            "pre_ctc_sub_sample": model_args.pre_ctc_sub_sample,
            "additional_layer": model_args.additional_layer,
            "additional_self_attention_layer": model_args.additional_self_attention_layer,
            "scb_layers": model_args.scb_layers,
        }
        clean_kwargs = {k: v for k, v in overwrite_args.items() if v is not None}
        self.model = DiCoWForConditionalGeneration.from_pretrained(self.model_type, **clean_kwargs)
        self.model.post_init()
        self.feature_extractor = feature_extractor or DiCoWFeatureExtractor.from_pretrained(self.model_type)
        if tokenizer is None:
            from transformers.models.whisper import WhisperTokenizerFast
            tokenizer = WhisperTokenizerFast.from_pretrained(self.model_type, predict_timestamps=predict_timestamps)
        self.tokenizer = tokenizer
        if ".en" not in self.model_type:
            self.model.generation_config.language = None
            self.model.generation_config.task = "transcribe"
            self.tokenizer.set_prefix_tokens(predict_timestamps=predict_timestamps, task="transcribe", language=global_lang_id)
        else:
            self.tokenizer.set_prefix_tokens(predict_timestamps=predict_timestamps)
        self.model.set_tokenizer(self.tokenizer)
        self.model.config.forced_decoder_ids = None
        if use_lora:
            raise NotImplementedError("use_lora: the B200 decoder is not built from nn.Linear forwards peft could wrap")
        if params_to_keep_frozen_keywords is not None:  # src/models/containers.py:80-90
            for name, param in self.model.named_parameters():
                param.requires_grad = not any(keyword in name for keyword in params_to_keep_frozen_keywords)

    def freeze_except(self, prefixes_to_preheat):
        for name, param in self.model.named_parameters():
            param.requires_grad = any(name.startswith(prefix) for prefix in prefixes_to_preheat)


def get_optimizer(model, training_args, prefixes_with_higher_lr=None):
    """src/models/containers.py:100-114: AdamW with a second group (higher learning rate, no weight decay) for the
    parameters whose names start with one of ``prefixes_with_higher_lr``"""
    prefixes = list(prefixes_with_higher_lr or [])
    if not training_args.use_custom_optimizer:
        return None
    base = [p for n, p in model.named_parameters() if not any(n.startswith(x) for x in prefixes)]
    new = [p for n, p in model.named_parameters() if any(n.startswith(x) for x in prefixes)]
    return torch.optim.AdamW([{"params": base},
                              {"params": new, "lr": training_args.fddt_lr_multiplier * training_args.learning_rate,
                               "weight_decay": 0.0}],
                             lr=training_args.learning_rate, weight_decay=training_args.weight_decay)
