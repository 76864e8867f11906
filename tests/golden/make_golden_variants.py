"""Golden outputs of the REFERENCE's own modules for the non-default variants of rows A4 / A5 / A9:
full-matrix FDDT (CustomLinear, src/models/dicow/layers.py:7-47), bias-only FDDT (src/models/dicow/FDDT.py:43-51) and the
additional encoder layer in front of the CTC head (src/models/dicow/encoder.py:16-17,88-93).  Same set-up as
make_golden.py (reference imported from /root/reference/src + the 4.55 compatibility shim; seeded synthetic weights).

    python tests/golden/make_golden_variants.py   ->  tests/golden/variants.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the shim, imports the reference)
from oracle import synth  # noqa: E402

BASE = {**synth.GOLDEN_MINI.__dict__, "use_enrollments": False, "scb_layers": 0}
VARIANTS = {
    "full_matrix": synth.Dims(**{**BASE, "fddt_is_diagonal": False}),
    "bias_only": synth.Dims(**{**BASE, "fddt_bias_only": True}),
    "additional_layer": synth.Dims(**{**BASE, "additional_layer": True}),
}

if __name__ == "__main__":
    torch.manual_seed(0)
    out = {}
    B = 2
    for name, dm in VARIANTS.items():
        model = mg.build_reference(dm)
        feats = torch.from_numpy(synth.make_features("v0", B, dm.n_mels, 2 * dm.T))
        stno = torch.from_numpy(synth.make_stno("v0", B, dm.T, "soft", pad_tail=5))
        enc = model.get_encoder()
        with torch.no_grad():
            out[name + "/enc"] = enc(feats, stno_mask=stno).last_hidden_state.numpy()
            out[name + "/ctc_logits"] = enc(feats, stno_mask=stno, return_logits=True).logits.numpy()
        print(name, out[name + "/enc"].shape, float(np.abs(out[name + "/enc"]).max()), out[name + "/ctc_logits"].shape)
    np.savez_compressed(os.path.join(HERE, "variants.npz"), **out)
