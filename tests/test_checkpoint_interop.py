"""Checkpoint / hub interop of the mirrored model classes (SURVEY.md section 8b "parameter naming/shape contract", 8(f).4),
on CPU (parameter containers only; no kernel is launched):

  * save_pretrained -> from_pretrained round-trips every tensor exactly and keeps proj_out tied to the decoder embedding
    (src/train.py:109-113) -- from_pretrained() runs _init_weights over EVERY module after loading, so an unguarded init
    there silently replaces the checkpoint;
  * from_pretrained(dir, **DiCoWConfig overrides) as WhisperContainer does (src/models/containers.py:24-52): modules the
    checkpoint does not have (SCB, CTC head, FDDT) get the reference's own inits, everything else is untouched;
  * the reference's re-initialisation calls (src/train.py:102-113, src/pretrain_encoder.py:39-40) work on the state_dict
    names: encoder-only load without the FDDT tables, whole-model load with the re-tied proj_out.
"""
import tempfile

import pytest
import torch

from oracle import synth


@pytest.fixture(scope="module")
def saved():
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    dm = synth.Dims(**{**synth.GOLDEN_MINI.__dict__, "use_enrollments": False, "scb_layers": 0})
    torch.manual_seed(0)
    m = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.3)  # "trained": nothing sits at its init value any more
    with tempfile.TemporaryDirectory() as d:
        m.save_pretrained(d)
        yield d, {k: v.clone() for k, v in m.state_dict().items()}, DiCoWForConditionalGeneration, dm


def _max_diff(a, b):
    return max((a[k].float() - b[k].float()).abs().max().item() for k in a)


def test_round_trip_exact_and_tied(saved):
    d, sd0, cls, _ = saved
    m = cls.from_pretrained(d)
    sd = m.state_dict()
    assert set(sd) == set(sd0)
    assert _max_diff(sd0, sd) == 0.0
    assert m.proj_out.weight.data_ptr() == m.model.decoder.embed_tokens.weight.data_ptr()
    assert m.config.model_type == "DiCoW" and m.config.use_fddt


def test_from_pretrained_with_overrides_adds_reference_inits(saved):
    d, sd0, cls, dm = saved
    m = cls.from_pretrained(d, use_enrollments=True, scb_layers=2)
    sd = m.state_dict()
    assert _max_diff(sd0, sd) == 0.0, "loaded tensors must not be re-initialised"
    cae = m.model.encoder.ca_enrolls[1].cae
    eye = torch.eye(dm.d)
    assert (cae.ffn[0].weight[:dm.d, :dm.d] - eye).abs().max() < 0.05   # layers.py:95-110: copy-through start
    assert (cae.ffn[3].weight[:, :dm.d] - eye).abs().max() < 0.05
    assert float(cae.cross_gate.gate.detach()) == 0.0                            # layers.py:79-93
    assert 0.01 < float(cae.cross_attn.q_proj.weight.std()) < 0.03      # HF init_std
    m2 = cls.from_pretrained(d, fddt_bias_only=True)                     # FDDT.py:10: bias vectors start at zero
    assert all(float(v.abs().max()) == 0.0 for k, v in m2.state_dict().items() if "fddt" in k)
    # full-matrix FDDT over a diagonal checkpoint: the [d] tables cannot be loaded into [d, d] (the reference raises the
    # same size mismatch); with ignore_mismatched_sizes they start from layers.py:37-44's suppressive identities
    m3 = cls.from_pretrained(d, fddt_is_diagonal=False, ignore_mismatched_sizes=True)
    w = m3.model.encoder.fddts[0].target_linear.weight
    assert w.shape == (dm.d, dm.d) and torch.allclose(w, torch.eye(dm.d))
    w = m3.model.encoder.initial_fddt.non_target_linear.weight  # encoder.py:66-73: the pre-positional FDDT suppresses
    assert torch.allclose(w, m3.config.non_target_fddt_value * torch.eye(dm.d))


def test_reference_reinit_calls(saved):
    d, sd0, cls, dm = saved
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    fresh = cls(DiCoWConfig(**dm.hf_kwargs()))
    # src/train.py:102-106: encoder weights without the FDDT tables
    enc_sd = {k[len("model.encoder."):]: v for k, v in sd0.items() if k.startswith("model.encoder.")}
    res = fresh.get_encoder().load_state_dict({k: v for k, v in enc_sd.items() if "fddt" not in k}, strict=False)
    assert res.unexpected_keys == [] and all("fddt" in k for k in res.missing_keys)
    assert torch.equal(fresh.state_dict()["model.encoder.layers.0.fc1.weight"], sd0["model.encoder.layers.0.fc1.weight"])
    # src/train.py:108-113: whole model, proj_out re-tied from the embedding
    state = {k: v for k, v in sd0.items() if k != "proj_out.weight"}
    state["proj_out.weight"] = state["model.decoder.embed_tokens.weight"]
    res = fresh.load_state_dict(state, strict=False)
    assert res.unexpected_keys == [] and res.missing_keys == []
    assert _max_diff(sd0, fresh.state_dict()) == 0.0
    # src/models/containers.py:80-97 / src/pretrain_encoder.py:42-51: name-keyword freezing works on these names
    for n, p in fresh.get_encoder().named_parameters():
        p.requires_grad = n.startswith(("additional_self_attention_layer", "lm_head", "subsample_conv"))
    n_train = sum(p.requires_grad for p in fresh.get_encoder().parameters())
    assert n_train == 4 + 3 + 2 + 1  # extra attention (q, k, v, out weights + 3 biases) + 2 sub-sampling convs + lm_head


def test_whisper_container_and_optimizer(saved):
    """src/models/containers.py:19-114 over a local checkpoint directory (no hub, no tokenizer files offline: a stub
    tokenizer with the attributes the model reads is passed in)"""
    import types
    d, sd0, cls, dm = saved
    from ts_asr_whisper_b200.containers import WhisperContainer, get_optimizer

    class Tok:
        prefix_tokens = [258, 259, 260]
        pad_token_id = 257

        def set_prefix_tokens(self, **kw):
            self.prefix_kw = kw

        def get_vocab(self):
            return {f"<|{0.02 * i:.2f}|>": 262 + i for i in range(38)}

    margs = types.SimpleNamespace(whisper_model=d, ctc_weight=0.3, fddt_is_diagonal=True, fddt_bias_only=False,
                                  fddt_use_silence=True, fddt_use_target=True, fddt_use_overlap=True, fddt_use_non_target=True,
                                  apply_fddt_to_n_layers=-1, fddt_init="suppressive", non_target_fddt_value=0.5,
                                  use_pre_pos_fddt=True, pre_ctc_sub_sample=True, additional_layer=False,
                                  additional_self_attention_layer=True, scb_layers=2)
    dargs = types.SimpleNamespace(use_timestamps=True, global_lang_id="en", use_enrollments=True)
    c = WhisperContainer(model_args=margs, data_args=dargs, use_fddt=True, params_to_keep_frozen_keywords=["decoder"],
                         tokenizer=Tok(), feature_extractor=object())
    assert c.tokenizer.prefix_kw == {"predict_timestamps": True, "task": "transcribe", "language": "en"}
    assert c.model.soft_label_creator is not None and c.model.config.forced_decoder_ids is None
    assert hasattr(c.model.get_encoder(), "ca_enrolls") and len(c.model.get_encoder().ca_enrolls) == 2
    assert all(p.requires_grad != ("decoder" in n) for n, p in c.model.named_parameters())
    c.freeze_except(["model.encoder.fddts", "model.encoder.initial_fddt", "model.encoder.ca_enrolls"])
    on = [n for n, p in c.model.named_parameters() if p.requires_grad]
    assert on and all(n.startswith(("model.encoder.fddts", "model.encoder.initial_fddt", "model.encoder.ca_enrolls")) for n in on)
    targs = types.SimpleNamespace(use_custom_optimizer=True, fddt_lr_multiplier=100.0, learning_rate=2e-6, weight_decay=0.01)
    opt = get_optimizer(c.model, targs, ["model.encoder.fddts", "model.encoder.ca_enrolls"])
    assert len(opt.param_groups) == 2 and abs(opt.param_groups[1]["lr"] - 2e-4) < 1e-12 and opt.param_groups[1]["weight_decay"] == 0.0
    assert get_optimizer(c.model, types.SimpleNamespace(use_custom_optimizer=False)) is None


def test_lora_adapters_like_the_reference_container(saved):
    """WhisperContainer(use_lora=True) (src/models/containers.py:69-90): rank-16 adapters on the decoder's q/k/v/out_proj/
    fc1/fc2, base model frozen, "lora_" parameters always trainable, B = 0 so the adapted model starts as the base model;
    adapter export in peft's key format round-trips; merging folds W + (alpha / r) B A into the base weight"""
    import types
    d, sd0, cls, dm = saved
    from ts_asr_whisper_b200.containers import WhisperContainer
    from ts_asr_whisper_b200.modeling_dicow import effective_weight

    class Tok:
        prefix_tokens = [258, 259, 260]
        pad_token_id = 257

        def set_prefix_tokens(self, **kw):
            pass

        def get_vocab(self):
            return {f"<|{0.02 * i:.2f}|>": 262 + i for i in range(38)}

    margs = types.SimpleNamespace(whisper_model=d, ctc_weight=0.3, fddt_is_diagonal=True, fddt_bias_only=False,
                                  fddt_use_silence=True, fddt_use_target=True, fddt_use_overlap=True, fddt_use_non_target=True,
                                  apply_fddt_to_n_layers=-1, fddt_init="suppressive", non_target_fddt_value=0.5,
                                  use_pre_pos_fddt=True, pre_ctc_sub_sample=True, additional_layer=False,
                                  additional_self_attention_layer=True, scb_layers=None)
    dargs = types.SimpleNamespace(use_timestamps=True, global_lang_id="en", use_enrollments=False)
    c = WhisperContainer(model_args=margs, data_args=dargs, use_fddt=True, use_lora=True,
                         params_to_keep_frozen_keywords=["decoder", "embed_positions"], tokenizer=Tok(), feature_extractor=object())
    m = c.model
    lora = [n for n, _ in m.named_parameters() if "lora_" in n]
    assert len(lora) == 2 * 10 * dm.dec_layers and all(".decoder." in n for n in lora)
    named = dict(m.named_parameters())
    assert all(named[n].requires_grad for n in lora)
    assert all(not p.requires_grad for n, p in named.items() if "decoder" in n and "lora_" not in n)
    assert any(p.requires_grad for n, p in named.items() if n.startswith("model.encoder.layers"))
    q = m.model.decoder.layers[0].self_attn.q_proj
    assert q.lora_A.shape == (16, dm.d) and q.lora_B.shape == (dm.d, 16) and q.lora_scale == 2.0
    assert float(q.lora_B.abs().max()) == 0.0 and torch.equal(effective_weight(q), q.weight.detach().float())
    with torch.no_grad():
        for n in lora:
            named[n].normal_(0.0, 0.05)
    sd = m.lora_state_dict()
    assert "base_model.model.model.decoder.layers.0.self_attn.q_proj.lora_A.weight" in sd and len(sd) == len(lora)
    want = effective_weight(q).clone()
    assert torch.allclose(want, q.weight.detach() + 2.0 * q.lora_B.detach() @ q.lora_A.detach(), atol=1e-6)
    m2 = cls.from_pretrained(d)
    m2.add_lora()
    m2.load_lora_state_dict(sd)
    assert torch.equal(effective_weight(m2.model.decoder.layers[0].self_attn.q_proj), want)
    m2.merge_lora()
    assert not [n for n, _ in m2.named_parameters() if "lora_" in n]
    assert torch.equal(m2.model.decoder.layers[0].self_attn.q_proj.weight.detach(), want)


def test_python_surface_of_survey_8b_exists():
    """the members src/train.py, src/pretrain_encoder.py, src/utils/trainers.py and utils/export_dicow.py touch"""
    import dataclasses
    from oracle import synth
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=False, scb_layers=0)
    model = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
    for name in ("from_pretrained", "post_init", "set_tokenizer", "get_encoder", "get_enc_logits", "generate", "forward",
                 "named_parameters", "state_dict", "load_state_dict", "register_for_auto_class", "save_pretrained"):
        assert callable(getattr(model, name)), name
    assert model.main_input_name == "input_features" and model.get_encoder().main_input_name == "input_features"
    model.config.forced_decoder_ids = None  # src/models/containers.py:68 (not a config field any more in transformers 5.x)
    for name in ("forced_decoder_ids", "decoder_start_token_id", "pad_token_id", "eos_token_id", "vocab_size"):
        assert hasattr(model.config, name), name
    unused = model.generation_config.update(max_new_tokens=7, num_beams=1, ctc_weight=0.2, length_penalty=0.1,
                                            return_timestamps=True, begin_suppress_tokens=None, forced_decoder_ids=None)
    assert model.generation_config.max_new_tokens == 7 and isinstance(unused, dict)
    enc = model.get_encoder()
    for name in ("forward", "get_loss", "get_max_len", "possibly_update_last_hidden_states"):
        assert callable(getattr(enc, name)), name
    assert enc.get_max_len() == 2 * dm.T
    DiCoWConfig.register_for_auto_class()
    DiCoWForConditionalGeneration.register_for_auto_class("AutoModelForSpeechSeq2Seq")  # utils/export_dicow.py:22-23


def test_reference_checkpoint_with_smoothing_matrix_buffer_loads_quietly(tmp_path, caplog):
    """a checkpoint saved by the reference after set_tokenizer() contains soft_label_creator.ts_smoothing_matrix
    (persistent buffer, modeling_dicow.py:33): from_pretrained accepts it without an unexpected-key report"""
    import dataclasses
    import logging
    import os
    from safetensors.torch import load_file, save_file
    from oracle import synth
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=False, scb_layers=0)
    model = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
    model.save_pretrained(tmp_path)
    path = os.path.join(tmp_path, "model.safetensors")
    sd = load_file(path)
    sd["soft_label_creator.ts_smoothing_matrix"] = torch.zeros(38, dm.vocab)
    save_file(sd, path, metadata={"format": "pt"})
    with caplog.at_level(logging.WARNING):
        again, info = DiCoWForConditionalGeneration.from_pretrained(tmp_path, output_loading_info=True)
    assert not info["unexpected_keys"], info["unexpected_keys"]
    for (n, a), (_, b) in zip(model.state_dict().items(), again.state_dict().items()):
        assert torch.equal(a, b), n


def test_training_refuses_dropout_it_would_not_apply():
    """training.trainable() is the gate both forward()s consult: zero dropout / LayerDrop (every Whisper checkpoint, the
    reference's defaults) passes; a config that asks for stochastic regularisation is refused in train mode only"""
    import dataclasses
    import pytest
    from oracle import synth
    from ts_asr_whisper_b200 import training
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=False, scb_layers=0)
    model = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs())).train()
    assert training.trainable(model) is True and training.trainable(model.get_encoder()) is True
    with torch.no_grad():
        assert training.trainable(model) is False
    for key, value in (("dropout", 0.1), ("attention_dropout", 0.1), ("final_dropout", 0.05), ("encoder_layerdrop", 0.1),
                       ("apply_spec_augment", True)):
        old = getattr(model.config, key)
        setattr(model.config, key, value)
        with pytest.raises(NotImplementedError, match=key):
            training.trainable(model)
        assert training.trainable(model.eval()) is True  # evaluation is deterministic either way
        model.train()
        setattr(model.config, key, old)
    assert training.trainable(model) is True


def test_forward_refuses_outputs_the_fused_path_does_not_produce():
    import dataclasses
    import pytest
    from oracle import synth
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=False, scb_layers=0)
    model = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
    feats = torch.zeros(1, dm.n_mels, 2 * dm.T)
    ids = torch.tensor([[258, 259]])
    for kw in (dict(output_attentions=True), dict(output_hidden_states=True), dict(head_mask=torch.ones(3, 2)),
               dict(decoder_head_mask=torch.ones(2, 2)), dict(past_key_values=()), dict(decoder_inputs_embeds=torch.zeros(1, 2, dm.d))):
        with pytest.raises(NotImplementedError):
            model(feats, stno_mask=torch.zeros(1, 4, dm.T), decoder_input_ids=ids, **kw)
