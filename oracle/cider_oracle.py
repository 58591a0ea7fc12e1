"""CPU oracle for the CIDEr-D reward of SCST.  TEST INFRASTRUCTURE ONLY (see ort_oracle.py).

Restates ``CiderScorer.compute_cider`` (sparse_caption/scst/cider/pyciderevalcap/ciderD/ciderD_scorer.py:133-212) and the
sample / baseline bookkeeping of ``CaptionScorer.__call__`` (sparse_caption/scst/scorers.py:47-114) on sequences of WORD IDS
instead of whitespace-split strings (an n-gram of words and the n-gram of their ids are the same object under a word-level
tokenizer).  Python floats / numpy float64 in the reference's own order of operations.
Pinned by tests/golden/ciderd.npz: scores of the imported reference ``CiderScorer`` on the same captions rendered as strings.
"""
import math
from collections import defaultdict
from typing import Dict, List, Sequence, Tuple

import numpy as np

Ngram = Tuple[int, ...]


def precook(words: Sequence[int], n: int = 4) -> Dict[Ngram, int]:
    """ciderD_scorer.py:19-35: n-gram counts, insertion order = (k ascending, position ascending)."""
    counts: Dict[Ngram, int] = defaultdict(int)
    for k in range(1, n + 1):
        for i in range(len(words) - k + 1):
            counts[tuple(words[i: i + k])] += 1
    return counts


def corpus_document_frequency(refs_per_image: Sequence[Sequence[Sequence[int]]], n: int = 4) -> Dict[Ngram, float]:
    """compute_doc_freq (:121-131): number of IMAGES whose references contain the n-gram ("corpus" mode)."""
    df: Dict[Ngram, float] = defaultdict(float)
    for refs in refs_per_image:
        for ng in set(ng for ref in refs for ng in precook(ref, n)):
            df[ng] += 1
    return df


def counts2vec(cnts: Dict[Ngram, int], df: Dict[Ngram, float], ref_len: float, n: int = 4):
    """:134-160.  NB the reference's ``length`` counts BIGRAMS (``if n == 1`` with n = len(ngram) - 1)."""
    vec = [dict() for _ in range(n)]
    length = 0
    norm = [0.0 for _ in range(n)]
    for ngram, tf in cnts.items():
        d = np.log(max(1.0, df.get(ngram, 0.0)))
        k = len(ngram) - 1
        vec[k][ngram] = float(tf) * (ref_len - d)
        norm[k] += pow(vec[k][ngram], 2)
        if k == 1:
            length += tf
    return vec, [np.sqrt(v) for v in norm], length


def sim(vec_h, vec_r, norm_h, norm_r, len_h, len_r, n: int = 4, sigma: float = 6.0):
    """:162-189: clipped cosine similarity per n-gram order with the Gaussian length penalty."""
    delta = float(len_h - len_r)
    val = np.array([0.0 for _ in range(n)])
    for k in range(n):
        for ngram in vec_h[k]:
            val[k] += min(vec_h[k][ngram], vec_r[k].get(ngram, 0.0)) * vec_r[k].get(ngram, 0.0)
        if norm_h[k] != 0 and norm_r[k] != 0:
            val[k] /= norm_h[k] * norm_r[k]
        val[k] *= np.e ** (-(delta ** 2) / (2 * sigma ** 2))
    return val


def ciderd(hyps: Sequence[Sequence[int]], refs: Sequence[Sequence[Sequence[int]]], df: Dict[Ngram, float], ref_len: float,
           n: int = 4, sigma: float = 6.0) -> np.ndarray:
    """One score per (hypothesis, reference set) pair (:191-212)."""
    scores = []
    for test, rs in zip(hyps, refs):
        vec, norm, length = counts2vec(precook(test, n), df, ref_len, n)
        score = np.array([0.0 for _ in range(n)])
        for ref in rs:
            vr, nr, lr = counts2vec(precook(ref, n), df, ref_len, n)
            score += sim(vec, vr, norm, nr, length, lr, n, sigma)
        s = np.mean(score)
        s /= len(rs)
        s *= 10.0
        scores.append(s)
    return np.array(scores)


def caption_scorer(refs, sample, baseline, df, ref_len, cider_weight: float = 1.0):
    """CaptionScorer.__call__ (scorers.py:47-114) with the CIDEr-D term: items = baselines first, then every sample of every
    image; returns (sc_sample [B * n], sc_baseline [B * n])."""
    nb = len(baseline) if baseline else 0
    ns = len(sample[0])
    hyps, rr = [], []
    for i in range(nb):
        hyps.append(baseline[i][0])
        rr.append(refs[i])
    for i in range(len(sample)):
        for j in range(ns):
            hyps.append(sample[i][j])
            rr.append(refs[i])
    scores = ciderd(hyps, rr, df, ref_len) * cider_weight
    sc_sample = scores[nb:]
    if baseline:
        sc_baseline = np.repeat(scores[:nb], ns)
    else:
        tot = sc_sample.reshape([-1, ns]).sum(-1)
        sc_baseline = (np.repeat(tot, ns) - sc_sample) / (ns - 1)
    return sc_sample, sc_baseline
