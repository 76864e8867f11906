"""GPU parity of the B200 DiCoWEncoder (C-ABI kernels) against the CPU oracle on identical synthetic weights/inputs.
Tolerance: north_star's bf16 bar -- max |err| <= 2e-2 x max |ref| (GEMM/attention operands are bf16, fp32 accumulate,
fp32 residual stream; the oracle is fp32 throughout)."""
import dataclasses

import pytest
import torch

from oracle import dicow_oracle as orc
from oracle import synth
from parity import row_metrics

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2
ROW_RMS_TOL, ROW_COS_MIN = 3e-2, 0.9995


def build_encoder(dm: synth.Dims, dev):
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling import DiCoWEncoder
    cfg = DiCoWConfig(**dm.hf_kwargs())
    enc = DiCoWEncoder(cfg)
    params = synth.make_params(dm, decoder=False)
    sd = {k[len("model.encoder."):]: torch.from_numpy(v) for k, v in params.items()}
    enc.load_state_dict(sd, strict=True)
    return enc.to(dev).eval(), orc.to_torch(params)


def rel_err(out, ref):
    return ((out.float().cpu() - ref).abs().max() / ref.abs().max()).item()


CASES = {
    # whisper-tiny + FDDT (BASELINE configs[0]) with the CTC head
    "tiny": (synth.WHISPER_TINY, 2, False),
    # large-v3-turbo dims, 2 layers (full-depth runs are covered by bench.py / smoke at B200 sizes)
    "turbo2": (dataclasses.replace(synth.LARGE_V3_TURBO, enc_layers=2, vocab=2047), 1, False),
    # BASELINE configs[1] at full depth: large-v3-turbo, 32 layers, 30 s windows (small vocabulary for the CTC head only)
    "turbo32": (dataclasses.replace(synth.LARGE_V3_TURBO, vocab=2047), 2, False),
    # SE-DiCoW: enrollment stream + 2 SCB layers, tiny dims
    "tiny_se": (dataclasses.replace(synth.WHISPER_TINY, use_enrollments=True, scb_layers=2, vocab=1000), 2, True),
    # the golden miniature (odd T, every feature on)
    "mini": (synth.GOLDEN_MINI, 2, True),
}


@pytest.mark.parametrize("name", list(CASES))
def test_encoder_matches_oracle(name):
    dm, B, use_enr = CASES[name]
    dev = torch.device("cuda:0")
    enc, p = build_encoder(dm, dev)
    feats = torch.from_numpy(synth.make_features(name, B, dm.n_mels, 2 * dm.T))
    stno = torch.from_numpy(synth.make_stno(name, B, dm.T, "soft", pad_tail=11))
    enr = None
    if use_enr:
        enr = {"input_features": torch.from_numpy(synth.make_features(name + "e", B, dm.n_mels, 2 * dm.T)),
               "stno_mask": torch.from_numpy(synth.make_stno(name + "e", B, dm.T, "hard"))}
    with torch.no_grad():
        ref = orc.encoder_forward(p, dm, feats, stno, enr)
        ref_logits = orc.ctc_logits(p, dm, ref)
    enr_d = {k: v.to(dev) for k, v in enr.items()} if enr else None
    out = enc(feats.to(dev), stno_mask=stno.to(dev), enrollments=enr_d)
    e1 = rel_err(out.last_hidden_state, ref)
    lo = enc(feats.to(dev), stno_mask=stno.to(dev), enrollments=enr_d, return_logits=True)
    e2 = rel_err(lo.logits, ref_logits)
    torch.cuda.synchronize()
    rms1, cos1 = row_metrics(out.last_hidden_state, ref)
    rms2, cos2 = row_metrics(lo.logits, ref_logits)
    print(f"{name}: hidden rel err {e1:.3e} (worst row rel RMS {rms1:.3e}, min row cos {cos1:.6f}), "
          f"ctc logits rel err {e2:.3e} (row RMS {rms2:.3e}, cos {cos2:.6f})")
    assert out.last_hidden_state.shape == ref.shape and lo.logits.shape == ref_logits.shape
    assert e1 < BF16_TOL and e2 < BF16_TOL
    # per frame (row): relative RMS error and cosine -- catches errors confined to low-magnitude rows / channels that the
    # max-norm over the whole tensor cannot see
    assert rms1 < ROW_RMS_TOL and cos1 > ROW_COS_MIN and rms2 < ROW_RMS_TOL and cos2 > ROW_COS_MIN


def test_golden_fixture_on_gpu():
    """CUDA path vs the committed reference outputs (tests/golden/mini_model.npz) directly."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mini_model.npz"))
    dm = synth.GOLDEN_MINI
    dev = torch.device("cuda:0")
    enc, _ = build_encoder(dm, dev)
    feats = torch.from_numpy(synth.make_features("g0", 2, dm.n_mels, 2 * dm.T)).to(dev)
    stno = torch.from_numpy(synth.make_stno("g0", 2, dm.T, "soft", pad_tail=7)).to(dev)
    enr = {"input_features": torch.from_numpy(synth.make_features("g0e", 2, dm.n_mels, 2 * dm.T)).to(dev),
           "stno_mask": torch.from_numpy(synth.make_stno("g0e", 2, dm.T, "hard")).to(dev)}
    out = enc(feats, stno_mask=stno, enrollments=enr).last_hidden_state
    assert rel_err(out, torch.from_numpy(g["enc_se"])) < BF16_TOL
    lg = enc(feats, stno_mask=stno, enrollments=enr, return_logits=True).logits
    assert rel_err(lg, torch.from_numpy(g["ctc_logits_se"])) < BF16_TOL


def test_no_cpu_fallback():
    from ts_asr_whisper_b200 import ops
    dm = synth.GOLDEN_MINI
    enc, _ = build_encoder(dm, torch.device("cuda:0"))
    with pytest.raises(ops.DicowError):
        enc.cpu()(torch.zeros(1, dm.n_mels, 2 * dm.T), stno_mask=torch.zeros(1, 4, dm.T))


@pytest.mark.parametrize("name,over", [("full_matrix", {"fddt_is_diagonal": False}), ("bias_only", {"fddt_bias_only": True}),
                                       ("additional_layer", {"additional_layer": True})])
def test_variants_match_reference_golden_and_oracle(name, over):
    """non-default variants of rows A4 / A5 / A9 (full-matrix FDDT, bias-only FDDT, additional encoder layer) against the
    reference's own outputs (tests/golden/variants.npz) and the oracle"""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "variants.npz"))
    base = {**synth.GOLDEN_MINI.__dict__, "use_enrollments": False, "scb_layers": 0}
    dm = synth.Dims(**{**base, **over})
    dev = torch.device("cuda:0")
    enc, p = build_encoder(dm, dev)
    feats = torch.from_numpy(synth.make_features("v0", 2, dm.n_mels, 2 * dm.T)).to(dev)
    stno = torch.from_numpy(synth.make_stno("v0", 2, dm.T, "soft", pad_tail=5)).to(dev)
    with torch.no_grad():
        out = enc(feats, stno_mask=stno).last_hidden_state
        lg = enc(feats, stno_mask=stno, return_logits=True).logits
    e1, e2 = rel_err(out, torch.from_numpy(g[name + "/enc"])), rel_err(lg, torch.from_numpy(g[name + "/ctc_logits"]))
    print(f"{name}: hidden rel err {e1:.3e}, ctc logits rel err {e2:.3e}")
    assert e1 < BF16_TOL and e2 < BF16_TOL
    # whisper-tiny widths as well (d = 384: other tile shapes), against the oracle
    dm2 = dataclasses.replace(synth.WHISPER_TINY, vocab=1000, enc_layers=2, T=200, **over)
    enc2, p2 = build_encoder(dm2, dev)
    f2 = torch.from_numpy(synth.make_features("v1", 2, dm2.n_mels, 2 * dm2.T))
    s2 = torch.from_numpy(synth.make_stno("v1", 2, dm2.T, "soft", pad_tail=9))
    with torch.no_grad():
        ref = orc.encoder_forward(p2, dm2, f2, s2)
        ref_l = orc.ctc_logits(p2, dm2, ref)
        o2 = enc2(f2.to(dev), stno_mask=s2.to(dev), return_logits=True)
    assert rel_err(o2.encoder_last_hidden_state, ref) < BF16_TOL and rel_err(o2.logits, ref_l) < BF16_TOL


@pytest.mark.parametrize("name,over", [("fddt_first_layer_only", {"apply_fddt_to_n_layers": 1}),
                                       ("no_pre_pos_fddt", {"use_pre_pos_fddt": False})])
def test_fddt_placement_variants_match_oracle(name, over):
    """apply_fddt_to_n_layers (encoder.py:45-47, 205) and use_pre_pos_fddt=False (encoder.py:62, 173-176): which layers carry
    an FDDT -- against the oracle, whose forward and gradients for these variants are pinned to the reference model live
    (tests/test_reference_live.py)"""
    base = {**synth.GOLDEN_MINI.__dict__, "use_enrollments": False, "scb_layers": 0}
    dev = torch.device("cuda:0")
    for dm, tag, B in ((synth.Dims(**{**base, **over}), "p0", 2),
                       (dataclasses.replace(synth.WHISPER_TINY, vocab=1000, enc_layers=3, T=200, **over), "p1", 2)):
        enc, p = build_encoder(dm, dev)
        f = torch.from_numpy(synth.make_features(tag, B, dm.n_mels, 2 * dm.T))
        s = torch.from_numpy(synth.make_stno(tag, B, dm.T, "soft", pad_tail=9))
        with torch.no_grad():
            ref = orc.encoder_forward(p, dm, f, s)
            ref_l = orc.ctc_logits(p, dm, ref)
            o = enc(f.to(dev), stno_mask=s.to(dev), return_logits=True)
        e1, e2 = rel_err(o.encoder_last_hidden_state, ref), rel_err(o.logits, ref_l)
        print(f"{name} d={dm.d}: hidden rel err {e1:.3e}, ctc logits rel err {e2:.3e}")
        assert e1 < BF16_TOL and e2 < BF16_TOL
