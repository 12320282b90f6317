"""CPU restatement of the product's on-device negative sampler (`ntf_neg_sample`, opentf_b200/csrc/sampler.cu)
-- TEST INFRASTRUCTURE, NOT PRODUCT.

The reference samples negatives with torch's global generator (fnn.py:48-76: top-ns of iid U(0,1) keys, or
torch.multinomial without replacement = top-ns of p_j/Exp(1)); that stream cannot be reproduced on a GPU, so the
contract is: same DISTRIBUTION as the reference, and bit-exact index sets between the device
sampler and this restatement of its documented counter RNG.  Two checks live in tests/:
  * bit-exactness: device indices == `sample_negatives(...)` here, for every mode and the edge cases;
  * distribution: chi-square of this sampler against `oracle.fnn_oracle.ns_unigram / ns_uniform` (the reference's
    arithmetic) on a small expert set.
Everything is integer arithmetic: Philox4x32-10 (Salmon et al., SC'11) -> 64-bit draw -> multiply-high onto the
range -> (unigram) binary search in the inclusive CDF of the integer all-team expert counts / (unigram_b) that entry of
the batch's concatenated member lists (expert j owns count_j of them: the batch unigram of fnn.py:74-76 without a
histogram) / (uniform) that expert id -> rejection of members / duplicates.
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
PURPOSE_NEG = 0x6e656730
MASK32 = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK32, p1 & MASK32, ((p0 >> 32) ^ c3 ^ k1) & MASK32, p0 & MASK32
        k0, k1 = (k0 + W0) & MASK32, (k1 + W1) & MASK32
    return c0, c1, c2, c3


def draw64(seed, row, t, step):
    v = philox4x32_10(row & MASK32, t & MASK32, step & MASK32, (step >> 32) & MASK32, seed & MASK32, ((seed >> 32) & MASK32) ^ PURPOSE_NEG)
    return (v[1] << 32) | v[0]


def expert_cdf(member_rows, E):
    """member_rows: list of index arrays -> (counts uint32 [E], inclusive cdf)"""
    counts = np.zeros(E, dtype=np.int64)
    for r in member_rows: np.add.at(counts, np.asarray(r, dtype=np.int64), 1)
    return counts, np.cumsum(counts)


def sample_row(nsd, seed, step, row_id, pos, E, ns, cdf=None, pool=None):
    """one team: `pos` = its member ids; `cdf` (unigram) = inclusive CDF of the all-team counts; `pool` (unigram_b) = the
    member ids of every team of the (global) batch, concatenated in row order.  Mirrors neg_sample_kernel statement by statement."""
    pos = [int(p) for p in pos]
    out, t = [], 0
    max_tries = 32 * ns + 64
    weighted = nsd in ('unigram', 'unigram_b')
    all_experts = False
    if nsd == 'unigram_b':
        T = len(pool)
        if all(int(j) in pos for j in pool): weighted, all_experts = False, True
    elif weighted:
        T = int(cdf[E - 1])
        pos_mass = sum(int(cdf[j]) - (int(cdf[j - 1]) if j else 0) for j in pos)
        if T == pos_mass: weighted, all_experts = False, True
    if weighted:
        tries = 0
        while len(out) < ns and tries < max_tries:
            x = (draw64(seed, row_id, t, step) * T) >> 64
            j = int(pool[x]) if nsd == 'unigram_b' else int(np.searchsorted(cdf, x, side='right'))
            if j not in pos and j not in out: out.append(j)
            tries += 1; t += 1
    avail = E if all_experts else E - len(pos)
    want = min(ns, avail)
    tries = 0
    while len(out) < want and tries < max_tries:
        j = (draw64(seed, row_id, t, step) * E) >> 64
        if (all_experts or j not in pos) and j not in out: out.append(j)
        tries += 1; t += 1
    j = 0
    while len(out) < want and j < E:
        if (all_experts or j not in pos) and j not in out: out.append(j)
        j += 1
    return out + [-1] * (ns - len(out))


def sample_negatives(nsd, seed, step, row0, member_rows, E, ns, cdf=None, pool_rows=None):
    """[B, ns] int32; `cdf` = inclusive CDF of the all-team counts (unigram); `pool_rows` = member lists of the global batch the
    rows are a slice of (unigram_b; default: the rows themselves)."""
    pool = None
    if nsd == 'unigram_b':
        src = member_rows if pool_rows is None else pool_rows
        pool = np.concatenate([np.asarray(r, dtype=np.int64) for r in src]) if len(src) else np.zeros(0, dtype=np.int64)
    return np.array([sample_row(nsd, seed, step, row0 + n, r, E, ns, cdf, pool) for n, r in enumerate(member_rows)], dtype=np.int32)
