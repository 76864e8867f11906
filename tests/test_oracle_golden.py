"""Pin the CPU oracle (oracle/dicow_oracle.py) to outputs of the reference itself (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference's own modules).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import dicow_oracle as orc
from oracle import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
EOS, SOT, LANG, TASK, NOTS, TS_BEGIN, N_TS = 257, 258, 259, 260, 261, 262, 38
SUPPRESS = [1, 2, 7, 8, 9, 10, 14, 25, 258, 259, 260]


@pytest.fixture(scope="module")
def mini():
    g = np.load(os.path.join(GOLD, "mini_model.npz"))
    dm = synth.GOLDEN_MINI
    p = orc.to_torch(synth.make_params(dm))
    B = 2
    feats = torch.from_numpy(synth.make_features("g0", B, dm.n_mels, 2 * dm.T))
    stno = torch.from_numpy(synth.make_stno("g0", B, dm.T, "soft", pad_tail=7))
    enr = {"input_features": torch.from_numpy(synth.make_features("g0e", B, dm.n_mels, 2 * dm.T)),
           "stno_mask": torch.from_numpy(synth.make_stno("g0e", B, dm.T, "hard"))}
    dm_plain = synth.Dims(**{**dm.__dict__, "use_enrollments": False, "scb_layers": 0})
    return g, dm, dm_plain, p, feats, stno, enr


def _close(a, b, tol=2e-5):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else a
    err = np.abs(a - b).max()
    assert err <= tol * max(1.0, np.abs(b).max()), f"max abs err {err}"


def test_encoder_se_dicow(mini):
    g, dm, _, p, feats, stno, enr = mini
    with torch.no_grad():
        _close(orc.encoder_forward(p, dm, feats, stno, enr), g["enc_se"])
        _close(orc.encoder_forward(p, dm, feats, stno, enr, return_logits=True), g["ctc_logits_se"])


def test_encoder_plain(mini):
    g, _, dmp, p, feats, stno, _ = mini
    with torch.no_grad():
        _close(orc.encoder_forward(p, dmp, feats, stno), g["enc_plain"])


def test_stno_rows_sum_to_one(mini):
    _, dm, _, _, _, stno, enr = mini
    assert torch.allclose(stno.sum(1), torch.ones_like(stno.sum(1)), atol=1e-6)
    assert torch.allclose(enr["stno_mask"].sum(1), torch.ones(2, dm.T))


def test_forward_losses(mini):
    g, _, dmp, p, feats, stno, _ = mini
    labels = torch.from_numpy(g["labels"])
    upp = torch.from_numpy(g["upp_labels"])
    assert np.array_equal(g["labels"], synth.make_labels("g0", 2, 12, dmp.vocab, EOS, TS_BEGIN, prefix=(LANG, TASK)))
    with torch.no_grad():
        loss, logits, _ = orc.model_forward(p, dmp, feats, stno, labels, upp, ctc_prefix_tokens=(SOT, LANG, TASK))
        _close(logits, g["fwd_logits"], 5e-5)
        _close(loss, g["fwd_hard_loss"], 2e-5)
        loss2, _, _ = orc.model_forward(p, dmp, feats, stno, labels, upp, ctc_prefix_tokens=(SOT, LANG, TASK),
                                        ts_begin=TS_BEGIN, n_ts=N_TS)
        _close(loss2, g["fwd_soft_loss"], 2e-5)


def test_greedy_decode_token_identical(mini):
    g, _, dmp, p, feats, stno, _ = mini
    with torch.no_grad():
        enc = orc.encoder_forward(p, dmp, feats, stno)
        prompt = torch.tensor([[SOT, LANG, TASK]] * 2)
        ids, lgs = orc.greedy_decode(p, dmp, enc, prompt, 24, suppress=SUPPRESS, no_timestamps=NOTS, ts_begin=TS_BEGIN,
                                     return_logits=True)
    _close(lgs[0], g["greedy_first_logits"], 5e-5)
    assert ids.tolist() == g["greedy_ids"].tolist()


def test_fddt_identity_at_init():
    """Known-answer identity: per-layer FDDT with all weights 1, biases 0 is the identity when the
    STNO rows sum to 1 (reference init: src/models/dicow/encoder.py:49-61)."""
    x = torch.from_numpy(synth.gaussish("kat/x", (2, 9, 16)))
    stno = torch.from_numpy(synth.make_stno("kat", 2, 9, "soft"))
    y = orc.fddt(x, stno, torch.ones(4, 16), torch.zeros(4, 16))
    assert torch.allclose(x, y, atol=1e-6)
    # initial_fddt at init: x * (v m_S + m_T + v m_N + m_O), v = non_target_fddt_value (encoder.py:62-73)
    v = 0.5
    w = torch.tensor([v, 1.0, v, 1.0])[:, None].expand(4, 16)
    y = orc.fddt(x, stno, w, torch.zeros(4, 16))
    scale = v * stno[:, 0] + stno[:, 1] + v * stno[:, 2] + stno[:, 3]
    assert torch.allclose(y, x * scale[..., None], atol=1e-6)


@pytest.mark.parametrize("n_mels", [80, 128])
def test_log_mel(n_mels):
    g = np.load(os.path.join(GOLD, "mel.npz"))
    _close(orc.mel_filterbank(n_mels), g[f"filters{n_mels}"], 1e-6)
    wav = synth.make_audio(f"mel{n_mels}", 41777)
    feat, mask = orc.log_mel(wav, n_mels, chunk_samples=32000)
    assert feat.shape == g[f"feat{n_mels}"].shape
    _close(feat, g[f"feat{n_mels}"], 2e-5)
    assert np.array_equal(mask, g[f"mask{n_mels}"])
