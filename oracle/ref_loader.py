"""Import the reference's own DiCoW modules from oracle/_ref (see oracle/make_ref.py) -- TEST / BASELINE INFRASTRUCTURE.

Shim: the reference pins transformers 4.55, where WhisperEncoderLayer.forward returns a tuple; the
installed 5.x returns the bare tensor and the reference's ``layer_outputs[0]`` would silently drop the batch dimension.  The
shim wraps the return value; nothing of the reference is edited."""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "dicow", "encoder.py"))


_loaded = None


def load():
    """-> (DiCoWConfig, DiCoWEncoder) classes of the reference"""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise ImportError("oracle/_ref is not built (python oracle/make_ref.py in the build container)")
    import transformers.models.whisper.modeling_whisper as mw
    if not getattr(mw.WhisperEncoderLayer.forward, "_dicow_tuple_shim", False):
        orig = mw.WhisperEncoderLayer.forward

        def fwd(self, hidden_states, attention_mask=None, layer_head_mask=None, output_attentions=False, **kw):
            out = orig(self, hidden_states, attention_mask, **kw)
            return out if isinstance(out, tuple) else (out,)

        fwd._dicow_tuple_shim = True
        mw.WhisperEncoderLayer.forward = fwd
    sys.path.insert(0, REF_ROOT)
    try:
        from models.dicow.config import DiCoWConfig
        from models.dicow.encoder import DiCoWEncoder
    finally:
        sys.path.remove(REF_ROOT)
    _loaded = (DiCoWConfig, DiCoWEncoder)
    return _loaded
