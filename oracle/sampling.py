"""CPU restatement of the reference's token sampling chain -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu legs).

Follows llm/src/Generate.cc in the order llm/src/nn_modules/cuda/LLaMAGenerate.cu:112-166 applies it:
  sample_repetition_penalty             Generate.cc:14-34
  sample_frequency_and_presence_penalties   Generate.cc:36-60
  temp <= 0: sample_token_greedy        Generate.cc:62-70
  else sample_top_k (min_keep 1)        Generate.cc:120-136
       sample_tail_free(z=1), sample_typical(p=1): identity at the reference's defaults (Generate.cc:203-206, 248-251 return early)
       sample_top_p (min_keep 1)        Generate.cc:304-327  (over sample_softmax, Generate.cc:81-101)
       sample_temperature               Generate.cc:72-76
       sample_token's sample_softmax    Generate.cc:103-118; the draw itself (std::discrete_distribution over std::mt19937) is replaced
                                        by an inverse-CDF lookup with a given uniform u, as the device sampler does.
Pinned against the compiled reference functions (oracle/_ref/libtce_ref_generate.so, tests/test_oracle_golden.py) and the committed
fixture tests/golden/sampling.npz.  float32 arithmetic in the reference's (sequential) order.
"""
import numpy as np

F = np.float32


def apply_penalties(logits, window, repeat_penalty=1.1, frequency_penalty=0.0, presence_penalty=0.0):
    out = np.array(logits, dtype=np.float32, copy=True)
    window = [int(t) for t in window]
    if len(window) == 0:
        return out
    counts = {}
    for t in window:
        counts[t] = counts.get(t, 0) + 1
    if F(repeat_penalty) != F(1.0):
        for t in counts:
            if 0 <= t < out.size:
                out[t] = out[t] * F(repeat_penalty) if out[t] <= 0 else out[t] / F(repeat_penalty)
    if F(frequency_penalty) != F(0.0) or F(presence_penalty) != F(0.0):
        for t, c in counts.items():
            if 0 <= t < out.size:
                out[t] = out[t] - (F(c) * F(frequency_penalty) + F(1.0) * F(presence_penalty))
    return out


def _softmax_sorted(l):
    """sample_softmax on logits already sorted descending: p = exp(l - l[0]) / sum, sequential float32 sum."""
    p = np.exp((l - l[0]).astype(np.float32)).astype(np.float32)
    s = F(0.0)
    for v in p:
        s = F(s + v)
    return (p / s).astype(np.float32)


def candidates(logits, window=(), top_k=40, top_p=0.95, temp=0.8, repeat_penalty=1.1, frequency_penalty=0.0, presence_penalty=0.0):
    """-> (ids, probs): the surviving candidates in logit-descending order (ties: ascending id, which the reference leaves unspecified)
    and the probabilities sample_token draws from.  temp <= 0: the greedy token with probability 1."""
    l = apply_penalties(logits, window, repeat_penalty, frequency_penalty, presence_penalty)
    n = l.size
    if temp <= 0:
        return np.array([int(np.argmax(l))], dtype=np.int32), np.array([1.0], dtype=np.float32)
    k = n if top_k <= 0 else min(max(int(top_k), 1), n)
    order = np.lexsort((np.arange(n), -l.astype(np.float64)))[:k]  # logit descending, id ascending
    ids = order.astype(np.int32)
    ll = l[ids]
    if top_p < 1.0:
        p = _softmax_sorted(ll)
        cum = F(0.0)
        last = ids.size
        for i in range(ids.size):
            cum = F(cum + p[i])
            if cum > F(top_p) and i >= 1:
                last = i
                break
        ids, ll = ids[:last], ll[:last]
    ll = (ll / F(temp)).astype(np.float32)
    return ids, _softmax_sorted(ll)


def draw(ids, probs, u):
    """inverse CDF: first candidate whose running probability exceeds u (u in [0, 1))."""
    run = F(0.0)
    for i in range(ids.size):
        run = F(run + probs[i])
        if F(u) < run:
            return int(ids[i])
    return int(ids[-1])


def uniform01(seed, idx):
    """splitmix64 of (seed, idx) -> [0, 1) with 24 bits, the device sampler's counter-based uniform (csrc/sampling.cu)."""
    m = (1 << 64) - 1
    z = (seed + 0x9E3779B97F4A7C15 * (idx + 1)) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    z = z ^ (z >> 31)
    return float(np.float32(z >> 40) * np.float32(1.0 / 16777216.0))
