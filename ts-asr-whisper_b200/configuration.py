"""DiCoWConfig -- same field names / defaults / model_type as the reference's config so checkpoints and recipes carry
over (reference: src/models/dicow/config.py:6-59; fields enumerated in SURVEY.md section 8a row A3)."""
from __future__ import annotations

from transformers import WhisperConfig

# field -> default, in the reference's order
_DICOW_FIELDS = {
    "ctc_loss_reduction": "mean",
    "final_dropout": 0.0,
    "ctc_zero_infinity": False,
    "ctc_weight": 0.0,
    "blank_token_id": None,
    "additional_layer": False,
    "additional_self_attention_layer": False,
    "pre_ctc_sub_sample": False,
    "use_fddt": True,
    "fddt_is_diagonal": True,
    "fddt_bias_only": False,
    "fddt_use_silence": True,
    "fddt_use_target": True,
    "fddt_use_overlap": True,
    "fddt_use_non_target": True,
    "remove_timestamps_from_ctc": False,
    "apply_fddt_to_n_layers": -1,
    "fddt_init": "suppressive",
    "non_target_fddt_value": 0.0,
    "use_enrollments": False,
    "scb_layers": None,
    "use_pre_pos_fddt": False,
}


class DiCoWConfig(WhisperConfig):
    model_type = "DiCoW"

    def __init__(self, **kwargs):
        extra = {k: kwargs.pop(k, default) for k, default in _DICOW_FIELDS.items()}
        super().__init__(**kwargs)
        for k, v in extra.items():
            setattr(self, k, v)

    # what the CUDA path supports today; anything else raises at model construction instead of silently diverging
    def check_supported(self) -> None:
        unsupported = []
        if self.d_model // self.encoder_attention_heads != 64 or self.d_model // self.decoder_attention_heads != 64:
            unsupported.append("head_dim != 64")
        if self.activation_function != "gelu":
            unsupported.append(f"activation_function={self.activation_function}")
        if unsupported:
            raise NotImplementedError("DiCoW B200 path does not implement: " + "; ".join(unsupported))
