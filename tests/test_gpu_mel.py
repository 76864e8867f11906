"""GPU parity of the fused log-mel kernel (dicow_logmel) against the committed reference outputs of the installed HF
WhisperFeatureExtractor (tests/golden/mel.npz, made by tests/golden/make_golden.py) and against the CPU oracle on
longer, multi-window recordings.  fp32 kernel: tolerance 1e-3 (north_star) -- measured error is ~1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import dicow_oracle as orc
from oracle import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3


@pytest.mark.parametrize("n_mels", [80, 128])
def test_logmel_matches_reference_golden(n_mels):
    from ts_asr_whisper_b200.feature_extraction import DiCoWFeatureExtractor
    g = np.load(os.path.join(GOLD, "mel.npz"))
    fe = DiCoWFeatureExtractor(feature_size=n_mels, chunk_length=2, device="cuda:0")
    np.testing.assert_allclose(fe.mel_filters.astype(np.float32), g[f"filters{n_mels}"], atol=1e-7)
    wav = synth.make_audio(f"mel{n_mels}", 41777)
    # exactly the reference call (src/data/local_datasets.py:208-214)
    f = fe(wav, return_tensors="pt", sampling_rate=16000, return_attention_mask=True, truncation=False,
           padding="longest", pad_to_multiple_of=fe.n_samples)
    feat = f.input_features[0].cpu().numpy()
    assert feat.shape == g[f"feat{n_mels}"].shape
    err = np.abs(feat - g[f"feat{n_mels}"]).max()
    print(f"log-mel M={n_mels}: max abs err vs reference golden {err:.3e}")
    assert err < TOL
    assert np.array_equal(f.attention_mask[0].cpu().numpy(), g[f"mask{n_mels}"])


@pytest.mark.parametrize("n_samples", [480000, 16000 * 75, 400, 161 * 160])
def test_logmel_matches_oracle(n_samples):
    """30 s window, a 75 s recording padded to 90 s (shared floor across windows), and tiny edge cases."""
    from ts_asr_whisper_b200 import ops
    dev = torch.device("cuda:0")
    n_mels = 128
    chunk = 480000 if n_samples >= 480000 else 160 * ((n_samples + 159) // 160)
    wavs = [synth.make_audio(f"melB{n_samples}_{i}", n_samples - 37 * i) for i in range(3)]
    refs = [orc.log_mel(w, n_mels, chunk_samples=chunk) for w in wavs]
    n_pad = refs[0][0].shape[1] * 160
    batch = torch.zeros(3, n_pad)
    for i, w in enumerate(wavs):
        batch[i, :len(w)] = torch.from_numpy(w)
    lengths = torch.tensor([len(w) for w in wavs], dtype=torch.int64, device=dev)
    filt = torch.from_numpy(orc.mel_filterbank(n_mels)).to(dev)
    out, mask = ops.logmel(batch.to(dev), filt, lengths, return_attention_mask=True)
    torch.cuda.synchronize()
    for i, (rf, rm) in enumerate(refs):
        assert out[i].shape == rf.shape
        err = np.abs(out[i].cpu().numpy() - rf).max()
        assert err < TOL, f"row {i}: max abs err {err}"
        assert np.array_equal(mask[i].cpu().numpy(), rm)


@pytest.mark.parametrize("name", ["three_spk", "one_spk", "no_target", "four_spk_first"])
def test_stno_mask_kernel_bit_exact(name):
    """A2 (src/data/local_datasets.py:162-196) on the GPU: bit-exact against the reference's own output and the oracle"""
    import os
    import numpy as np
    from ts_asr_whisper_b200 import ops
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stno_mask.npz"))
    n_spk, n_samples, target = [int(v) for v in g[name + "/meta"]]
    act = np.unpackbits(g[name + "/activity"], axis=1)[:, :n_samples].astype(bool)
    out = ops.stno_mask(torch.from_numpy(act).to("cuda:0"), target)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), g[name + "/stno"])
    assert np.array_equal(out.cpu().numpy(), orc.stno_mask(act, target))
    cf = ops.stno_mask(torch.from_numpy(act).to("cuda:0"), target, channels_first=True)
    assert torch.equal(cf.t().contiguous(), out)
