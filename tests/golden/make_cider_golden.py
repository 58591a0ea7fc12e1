"""Golden CIDEr-D scores from the IMPORTED reference scorer (sparse_caption/scst/cider/pyciderevalcap/ciderD/ciderD_scorer.py)
on captions of random word ids rendered as strings.  Run in the build container only:  python tests/golden/make_cider_golden.py
Two document-frequency modes: "corpus" (df from the given references) and a cached table (what ``coco-train-words.p`` provides:
the reference then reads ``document_frequency`` / ``ref_len`` instead of computing them, ciderD_scorer.py:78-84, 214-220)."""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")
from sparse_caption.scst.cider.pyciderevalcap.ciderD.ciderD_scorer import CiderScorer  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
rng = np.random.RandomState(1234)
V, B, NREF, NS, L = 60, 24, 5, 6, 16     # small vocabulary: plenty of shared n-grams


def caption(lo=3, hi=L):
    n = rng.randint(lo, hi + 1)
    # a Zipf-ish unigram distribution so that n-grams repeat inside and across captions
    return list((rng.zipf(1.5, size=n) % (V - 4) + 4).astype(int))


refs = [[caption() for _ in range(rng.randint(3, NREF + 1))] for _ in range(B)]
hyps = [[caption(0, L) if rng.rand() > 0.3 else list(refs[i][rng.randint(len(refs[i]))][: rng.randint(1, L)]) for _ in range(NS)] for i in range(B)]
to_s = lambda ids: " ".join(f"w{i}" for i in ids)


def run(df_table=None, n_docs=None):
    sc = CiderScorer(df_mode="corpus")
    for i in range(B):
        for j in range(NS):
            sc += (to_s(hyps[i][j]), [to_s(r) for r in refs[i]])
    if df_table is not None:
        sc.df_mode = "cached"
        sc.document_frequency = df_table
        sc.ref_len = np.log(float(n_docs))
    _, scores = sc.compute_score()
    return np.asarray(scores, dtype=np.float64)


corpus_scores = run()
# a cached table: document frequencies of an independent "training corpus" of 500 images
from collections import defaultdict  # noqa: E402
from sparse_caption.scst.cider.pyciderevalcap.ciderD.ciderD_scorer import precook  # noqa: E402
train = [[caption() for _ in range(5)] for _ in range(500)]
table = defaultdict(float)
for rs in train:
    for ng in set(ng for r in rs for ng in precook(to_s(r), 4)):
        table[ng] += 1
cached_scores = run(table, 500)
keys = sorted(table)
pad = lambda seqs, w: np.array([list(s) + [-1] * (w - len(s)) for s in seqs], dtype=np.int64)
blob = {
    "refs": pad([r for rs in refs for r in rs], L), "ref_count": np.array([len(rs) for rs in refs]),
    "hyps": pad([h for hs in hyps for h in hs], L), "corpus_scores": corpus_scores, "cached_scores": cached_scores,
    "df_ngrams": pad([[int(w[1:]) for w in k] for k in keys], 4), "df_counts": np.array([table[k] for k in keys]), "df_docs": np.array(500),
}
np.savez_compressed(os.path.join(OUT, "ciderd.npz"), **blob)
print({k: v.shape for k, v in blob.items()}, float(corpus_scores.mean()), float(cached_scores.mean()))
