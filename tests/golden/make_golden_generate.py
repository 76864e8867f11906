"""Golden vectors for the host-side generate() logic, produced by calling the REFERENCE's own functions in the build
container (needs /root/reference; not available on the GPU box):

  * DiCoWGenerationMixin._retrieve_segment (staticmethod, src/models/dicow/generation.py:415-534) on crafted token
    sequences that reach every branch (timestamp pairs, single-timestamp ending, lone timestamp, rollback, no
    timestamps, empty);
  * DiCoWGenerationMixin._fix_timestamps_from_segmentation (generation.py:322-413) with a recording fake tokenizer:
    the golden is the text the reference asks the tokenizer to encode.

    python tests/golden/make_golden_generate.py   ->  tests/golden/generate_logic.json
"""
from __future__ import annotations

import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/src")

from models.dicow.generation import DiCoWGenerationMixin  # noqa: E402

TS = 50365  # <|0.00|> of the multilingual vocabulary (export_sources/generation_config.json)


def t(x):  # timestamp token for x seconds
    return TS + int(round(x / 0.02))


RETRIEVE_CASES = {
    "pairs_single_end": [t(0.0), 11, 12, t(2.5), t(2.5), 13, 14, 15, t(7.0)],
    "pairs_double_end": [t(0.0), 11, t(1.0), t(1.0), 12, 13, t(4.0), t(4.0)],
    "pair_then_open_text": [t(0.4), 11, 12, t(3.0), t(3.2), 13, 14],
    "one_ts_small": [t(1.0), 11, 12, 13],
    "one_ts_rollback": [t(9.0), 11, 12],
    "two_ts_no_pair": [t(0.5), 11, 12, t(6.0), 13],
    "no_ts": [11, 12, 13],
    "single_token_text": [11],
    "only_ts": [t(0.0)],
    "empty": [],
    "three_pairs": [t(0.0), 5, t(0.5), t(0.5), 6, t(1.5), t(1.5), 7, 8, t(29.98), t(29.98)],
}


def run_retrieve():
    out = {}
    for name, toks in RETRIEVE_CASES.items():
        for seek_frames, off in ((3000, 0.0), (1234, 60.0)):
            seq = torch.tensor(toks, dtype=torch.long)
            try:
                segs, offset = DiCoWGenerationMixin._retrieve_segment(
                    seek_sequence=seq, seek_outputs=[None, None], time_offset=torch.tensor([0.0, off], dtype=torch.float64),
                    timestamp_begin=TS, seek_num_frames=torch.tensor([0, seek_frames]), time_precision=0.02,
                    time_precision_features=0.01, input_stride=2, prev_idx=1, idx=1, return_token_timestamps=False,
                    decoder_input_ids=torch.zeros(2, 3, dtype=torch.long))
                res = {"offset": int(offset), "segments": [
                    {"start": float(s["start"]), "end": float(s["end"]), "tokens": [int(x) for x in s["tokens"]]}
                    for s in segs]}
            except Exception as ex:  # noqa: BLE001  (empty sequences raise inside the reference)
                res = {"error": type(ex).__name__}
            out[f"{name}/{seek_frames}"] = {"tokens": toks, "seek_num_frames": seek_frames, "time_offset": off, **res}
    return out


class RecordingTokenizer:
    """decode(tokens) -> 'w<id>' words; __call__(text) records the text and returns its bytes as ids"""
    pad_token_id = 50257

    def __init__(self):
        self.texts = []

    def get_vocab(self):
        return {"<|0.00|>": TS, "Ġ": 220}

    def decode(self, toks):
        return "".join(f" w{int(x)}" for x in toks)

    def __call__(self, text):
        self.texts.append(text)
        return {"input_ids": list(text.encode())}


FIX_CASES = {
    "within_one_block": [(0.0, 2.5, [11, 12]), (2.5, 7.0, [13])],
    "crossing_blocks": [(1.0, 4.0, [11]), (28.0, 31.5, [12, 13]), (31.5, 33.0, [14]), (95.0, 97.0, [15])],
    "starts_late": [(61.0, 62.0, [11])],
    "ends_on_boundary": [(25.0, 30.0, [11]), (30.0, 35.5, [12]), (59.0, 60.0, [13]), (60.0, 90.0, [14])],
    "thirty_second_segment": [(10.0, 40.0, [11]), (40.0, 41.0, [12])],
    "dummy_and_empty": [(0.0, 0.0, [TS]), (3.0, 3.5, []), (4.0, 5.0, [11])],
}


def run_fix():
    out = {}
    for name, segs in FIX_CASES.items():
        tok = RecordingTokenizer()
        self = types.SimpleNamespace(tokenizer=tok, round_to_nearest_0_02=DiCoWGenerationMixin.round_to_nearest_0_02)
        seq = {"sequences": torch.zeros(1, 1, dtype=torch.long),
               "segments": [[{"start": torch.tensor(a, dtype=torch.float64), "end": torch.tensor(b, dtype=torch.float64),
                              "tokens": torch.tensor(tk, dtype=torch.long)} for a, b, tk in segs]]}
        res = DiCoWGenerationMixin._fix_timestamps_from_segmentation(self, seq)
        out[name] = {"segments": [[a, b, tk] for a, b, tk in segs], "text": tok.texts[0], "ids": res[0].tolist()}
    return out


if __name__ == "__main__":
    data = {"timestamp_begin": TS, "retrieve_segment": run_retrieve(), "fix_timestamps": run_fix()}
    with open(os.path.join(HERE, "generate_logic.json"), "w") as f:
        json.dump(data, f, indent=1)
    print({k: len(v) if isinstance(v, dict) else v for k, v in data.items()})
    for k, v in data["retrieve_segment"].items():
        print(k, v.get("offset"), v.get("error"), len(v.get("segments", [])))
    for k, v in data["fix_timestamps"].items():
        print(k, v["text"])
