"""Golden vectors for the radix detokenisation from the IMPORTED reference tokenizer class
(sparse_caption/tokenizer.py:550-725).  Run in the build container only:  python tests/golden/make_radix_golden.py

The reference's ``RadixTokenizer.__init__`` trains / loads a SentencePiece model; the radix arithmetic itself
(``_decode_radix_ids``, ``grouper``, ``base_to_decimal``, ``decimal_to_base``, the radix map) is pure Python, so an instance is
created without ``__init__`` and given the two attributes those methods read."""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")
from sparse_caption.tokenizer import RadixTokenizer  # noqa: E402
from sparse_caption.utils.config import Config  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
rng = np.random.RandomState(8888)
blob = {}
for name, base, vocab, L in (("b768", 768, 10000, 26), ("b256", 256, 9487, 30), ("b32", 32, 9000, 41)):
    tok = RadixTokenizer.__new__(RadixTokenizer)
    tok.config = Config(radix_base=base)
    tok.tokens_per_word = len(tok.decimal_to_base(vocab - 3, base))
    # the radix map of __init__ (:557-571), through the reference's own static method
    radix_map = {}
    for i in range(vocab - 3):
        r = tok.decimal_to_base(i, base)
        radix_map[i] = [1] * (tok.tokens_per_word - len(r)) + r
    radix_map[-4], radix_map[-3], radix_map[-2], radix_map[-1] = [0], radix_map[vocab - 4], [base + 1], [base + 2]
    tok.radix_map = radix_map
    n = 96
    seqs = np.zeros((n, L), dtype=np.int64)
    words_out = -np.ones((n, (L + tok.tokens_per_word - 1) // tok.tokens_per_word), dtype=np.int64)
    for k in range(n):
        n_words = rng.randint(0, L // tok.tokens_per_word + 1)
        words = rng.randint(4, vocab, size=n_words)
        ids = [x for w in words for x in radix_map[w - 4]]
        style = k % 4
        if style == 0:
            ids = ids + [tok.eos_token_id]
        elif style == 1:
            ids = ids[: max(0, len(ids) - 1)] + [tok.eos_token_id]      # ragged last group
        elif style == 2:
            ids = ids                                                   # no <eos>: the whole row is the caption
        else:
            ids = ids + [tok.eos_token_id] + list(rng.randint(1, base + 1, size=3))  # garbage after <eos>
        ids = (ids + [0] * L)[:L] if style != 2 else (ids + list(rng.randint(1, base + 1, size=L)))[:L]
        seqs[k] = ids
        got = tok._decode_radix_ids(list(int(x) for x in ids))          # <- the reference
        words_out[k, : len(got)] = got
    blob[f"{name}_seq"] = seqs
    blob[f"{name}_words"] = words_out
    blob[f"{name}_meta"] = np.array([base, tok.tokens_per_word, tok.eos_token_id, tok.bos_token_id])
np.savez_compressed(os.path.join(OUT, "radix_detok.npz"), **blob)
print({k: v.shape for k, v in blob.items()})
