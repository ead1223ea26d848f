"""Tokenizer hook of the reference (mreserve/lowercase_encoder.py): special-token ids and `get_encoder()`.

The BPE vocabulary (`lowercase_encoder.json`, 0.9 MB) is a data file of the reference and is not redistributed
here; `get_encoder` loads it from, in this order: the `path` argument, $MRESERVE_TOKENIZER_JSON, a copy placed next
to this module, `~/.cache/merlotreserve/lowercase_encoder.json`.  Tokenisation itself is CPU-side text processing
outside the hot path (out of scope); `PretrainedMerlotReserve.get_label_space` only needs the ids."""
import os
from typing import Optional

os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")

PADDING, START, END, MASK, MASKAUDIO, AUDIOSPAN, LTOVPOOL, RESETCTX = 0, 1, 2, 3, 4, 5, 6, 9   # lowercase_encoder.py:9-16
SPECIAL_TOKENS = ["<|PAD|>", "<|START|>", "<|END|>", "<|MASK|>", "<|MASKAUDIO|>", "<|AUDIOSPAN|>", "<|LTOVPOOL|>"] + \
                 [f"<|unused{i}|>" for i in range(3)]                                          # lowercase_encoder.py:18


def get_encoder(path: Optional[str] = None):
    """lowercase_encoder.py:20-24: a `tokenizers.Tokenizer` built from the reference's vocabulary file."""
    from tokenizers import Tokenizer
    here = os.path.dirname(os.path.abspath(__file__))
    candidates = [path, os.environ.get("MRESERVE_TOKENIZER_JSON"), os.path.join(here, "lowercase_encoder.json"),
                  os.path.join(os.path.expanduser("~"), ".cache", "merlotreserve", "lowercase_encoder.json")]
    for fn in candidates:
        if fn and os.path.exists(fn):
            return Tokenizer.from_file(fn)
    raise FileNotFoundError("lowercase_encoder.json not found; pass its path or set MRESERVE_TOKENIZER_JSON "
                            "(the file ships with the reference: mreserve/lowercase_encoder.json)")
