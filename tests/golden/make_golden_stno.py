"""Golden STNO masks from the REFERENCE's own `_create_stno_masks` (src/data/local_datasets.py:186-196).  The module
imports lhotse (not installed), so the function's source is cut out of the file with `ast` and executed as is; the
down-sampling in front of it (local_datasets.py:167-175: pad to whole 30 s windows, mean over 320 samples) is restated here
line by line.

    python tests/golden/make_golden_stno.py   ->  tests/golden/stno_mask.npz
"""
import ast
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/data/local_datasets.py"


def reference_create_stno_masks():
    tree = ast.parse(open(SRC).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "_create_stno_masks":
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np}
            exec(compile(mod, SRC, "exec"), ns)
            return ns["_create_stno_masks"]
    raise RuntimeError("_create_stno_masks not found")


def activity(rng, n_spk, n_samples):
    a = np.zeros((n_spk, n_samples), dtype=bool)
    for s in range(n_spk):
        t = 0
        while t < n_samples:
            gap, dur = int(rng.integers(0, 60000)), int(rng.integers(2000, 90000))
            a[s, t + gap:t + gap + dur] = True
            t += gap + dur
    return a


if __name__ == "__main__":
    create = reference_create_stno_masks()
    rng = np.random.default_rng(5)
    out = {}
    for name, n_spk, n_samples, target in (("three_spk", 3, 16000 * 41 + 123, 1), ("one_spk", 1, 480000, 0),
                                           ("no_target", 2, 16000 * 12, -1), ("four_spk_first", 4, 16000 * 75, 0)):
        act = activity(rng, n_spk, n_samples)
        spk_mask = act
        pad_len = (480000 - spk_mask.shape[-1]) % 480000                      # local_datasets.py:168-169
        spk_mask = np.pad(spk_mask, ((0, 0), (0, pad_len)), mode="constant")
        spk_mask = spk_mask.astype(np.float32).reshape(spk_mask.shape[0], -1, 2 * 160).mean(axis=-1)   # :172-174
        if target == -1:                                                      # :176-178
            spk_mask = np.pad(spk_mask, ((0, 1), (0, 0)), mode="constant")
        out[name + "/activity"] = np.packbits(act, axis=1)
        out[name + "/meta"] = np.array([n_spk, n_samples, target])
        out[name + "/stno"] = create(spk_mask, target).astype(np.float32)
        print(name, out[name + "/stno"].shape, out[name + "/stno"].sum(0))
    np.savez_compressed(os.path.join(HERE, "stno_mask.npz"), **out)
