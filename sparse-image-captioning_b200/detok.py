"""Step after the path: radix token ids -> word ids for whole batches of decoded captions.

The reference detokenises caption by caption in Python (`RadixTokenizer._decode_radix_ids`,
sparse_caption/tokenizer.py:595-602: cut at the first <eos>, group `tokens_per_word` digits with fill value 1, then
`base_to_decimal` (:706-712: sum(max(d - 1, 0) * radix**i)) + 4).  Here the same arithmetic runs vectorised over the
[B, beam, L] int tensor the engine returns - on the device, or on the host after the D2H copy - so that the end-to-end
captions/s of the ACORT (radix) configuration is not bounded by a Python loop over 512 x beam captions.
"""
import torch


def radix_to_word_ids(seq: torch.Tensor, radix_base: int, tokens_per_word: int, eos_id: int = None):
    """seq: int tensor [..., L] of radix token ids (digits are 1-based: digit value = id - 1; ids above the base are the
    special tokens).  Returns (word_ids int64 [..., ceil(L / tokens_per_word)], n_words int64 [...]): caption i has
    n_words[i] valid entries, the rest are 0 (<pad>)."""
    if eos_id is None:
        eos_id = radix_base + 2  # tokenizer.py:662-666
    L = seq.shape[-1]
    s = seq.long()
    is_eos = s == eos_id
    # position of the first <eos> (L when absent)
    pos = torch.arange(L, device=s.device).expand_as(s)
    first = torch.where(is_eos, pos, torch.full_like(pos, L)).min(-1).values
    keep = pos < first.unsqueeze(-1)
    digit = torch.where(keep, (s - 1).clamp_min(0), torch.zeros_like(s))  # fill value 1 -> digit 0
    G = (L + tokens_per_word - 1) // tokens_per_word
    pad = G * tokens_per_word - L
    if pad:
        digit = torch.nn.functional.pad(digit, (0, pad))
    digit = digit.view(*s.shape[:-1], G, tokens_per_word)
    weights = torch.tensor([radix_base ** i for i in range(tokens_per_word - 1, -1, -1)], device=s.device, dtype=torch.long)
    words = (digit * weights).sum(-1) + 4
    n_words = (first + tokens_per_word - 1) // tokens_per_word
    valid = torch.arange(G, device=s.device).expand(*s.shape[:-1], G) < n_words.unsqueeze(-1)
    return torch.where(valid, words, torch.zeros_like(words)), n_words
