"""CPU model checks of the two selection algorithms of csrc/infer_topk.cu (no GPU, no product code): the adaptive radix select of
blockmax_threshold_kernel (256 bins over the current key range, stop at <= 128 survivors, restart on block numbers when more than 128 keys are
equal) and raise_bar of infer_topk_final_kernel (a bar with K <= count <= limit over unique 64-bit composites), restated step by step in
Python and compared with a plain sort on random keys, heavy ties and plateaus.  They pin the ALGORITHMS' corner cases (the device kernels are
checked against the oracle in tests/test_gpu_topk_fused.py); a logic change in the kernels should be mirrored here first."""
import numpy as np
def clz32(x): return 32 - int(x).bit_length()
def clz64(x): return 64 - int(x).bit_length()
def select_model(keys, K):
    """blockmax_threshold_kernel's algorithm: -> (Tkey, set of the K blocks at or above the K-th composite)"""
    nblk = len(keys)
    key = [int(k) for k in keys]
    v = list(key)
    lo, hi = min(key), max(key)
    need = K
    tie_key = 0
    while True:
        width = hi - lo
        shift = 0 if width < 256 else 24 - clz32(width)
        assert (width >> shift) < 256
        hist = [0] * 256
        for x in v:
            d = (x - lo) & 0xffffffff
            if d <= width: hist[d >> shift] += 1
        # scan from top
        cum = 0; sel = None
        for d in range(255, -1, -1):
            if cum < need <= cum + hist[d]: sel = (d, hist[d], cum); break
            cum += hist[d]
        assert sel is not None
        b, cbin, above = sel
        need -= above
        lo += b << shift
        hi = min(hi, lo + ((1 << shift) - 1))
        if cbin <= 128: break
        if shift == 0:
            tie_key = lo
            v = [(4096 - i) if key[i] == tie_key else 0 for i in range(nblk)]
            lo, hi = 1, 4096
    lst = [((key[i] << 12) | (4095 - i)) for i in range(nblk) if 0 <= ((v[i] - lo) & 0xffffffff) <= hi - lo]
    assert len(lst) <= 128, len(lst)
    kth = None
    for x in lst:
        if sum(1 for y in lst if y > x) == need - 1: kth = x
    assert kth is not None
    chosen = {i for i in range(nblk) if ((key[i] << 12) | (4095 - i)) >= kth}
    return kth >> 12, chosen
def raise_bar_model(comps, K, limit, kz):
    lo, hi, need = kz << 32, (1 << 64) - 1, K
    while True:
        width = hi - lo
        shift = 0 if width < 256 else 56 - clz64(width)
        assert (width >> shift) < 256
        hist = [0] * 256
        for x in comps:
            if x and lo <= x <= hi: hist[(x - lo) >> shift] += 1
        cum, d = 0, 255
        while d > 0:
            if cum + hist[d] >= need: break
            cum += hist[d]; d -= 1
        done = ((K - need) + cum + hist[d] <= limit) or shift == 0
        blo, nn = lo + (d << shift), need - cum
        if done: return blo
        hi = min(hi, blo + ((1 << shift) - 1)); lo = blo; need = nn


def okey(f):
    b = np.float32(f).view(np.uint32).item()
    return (~b & 0xffffffff) if b & 0x80000000 else (b | 0x80000000)


def test_adaptive_radix_select_model_equals_a_sort():
    rng = np.random.default_rng(1)
    for trial in range(240):
        nblk = int(rng.integers(1, 1500)); K = int(rng.integers(1, min(nblk, 1024) + 1))
        mode = trial % 4
        if mode == 0: vals = rng.normal(size=nblk)
        elif mode == 1: vals = np.round(rng.normal(size=nblk), 1)  # many ties
        elif mode == 2:
            vals = np.zeros(nblk); vals[rng.integers(0, nblk, size=3)] = 1.0  # a plateau: the tie restart
        else: vals = rng.normal(size=nblk) * 1e-3 - 5
        keys = [okey(x) for x in vals.astype(np.float32)]
        T, chosen = select_model(keys, K)
        order = sorted(range(nblk), key=lambda i: (-keys[i], i))
        assert chosen == set(order[:K]) and T == keys[order[K - 1]], (trial, nblk, K)


def test_raise_bar_model_keeps_the_top_k_within_the_limit():
    rng = np.random.default_rng(2)
    for trial in range(150):
        n = int(rng.integers(1100, 9000)); K = int(rng.integers(33, 1025)); limit = int(rng.choice([1024, 4096]))
        mode = trial % 3
        vals = rng.normal(size=n) if mode == 0 else (np.round(rng.normal(size=n), 1) if mode == 1 else np.zeros(n))
        ks = [okey(x) for x in vals.astype(np.float32)]
        comps = [(ks[i] << 32) | ((~i) & 0xffffffff) for i in range(n)]
        bar = raise_bar_model(comps, K, limit, min(ks))
        cnt = sum(1 for x in comps if x >= bar)
        assert K <= cnt <= limit, (trial, K, cnt, limit)
        assert all(x >= bar for x in sorted(comps, reverse=True)[:K])
