"""CPU restatements of two pieces of device-side selection logic, checked exhaustively/randomly without a GPU:

* the bitonic network of `cta_sort_trim` (cvt_b200/csrc/pq_kernels.cu): the index arithmetic
  i = 2t - (t & (stride - 1)), j = i + stride, ascending iff (i & size) == 0 must sort any power-of-two buffer;
* the shared bound of the one-pass u8 scan (cvt_b200/csrc/u8_scan_tc.cu, the bound warp): the k-th smallest of the
  minima of DISJOINT lists is an upper bound on the k-th smallest element overall, and rows AT the bound must be kept.
"""
import numpy as np
import pytest


def bitonic_like_kernel(a):
    a = a.copy()
    n2 = len(a)
    size = 2
    while size <= n2:
        stride = size >> 1
        while stride > 0:
            t = np.arange(n2 >> 1)
            i = 2 * t - (t & (stride - 1))
            j = i + stride
            assert np.all((i & stride) == 0) and np.all(j < n2) and len(np.unique(np.concatenate([i, j]))) == n2
            up = (i & size) == 0
            x, y = a[i], a[j]
            swap = (x > y) == up
            a[i] = np.where(swap, y, x)
            a[j] = np.where(swap, x, y)
            stride >>= 1
        size <<= 1
    return a


@pytest.mark.parametrize("n2", [256, 512, 1024, 2048, 4096])
def test_cta_sort_trim_network_sorts(n2):
    rng = np.random.Generator(np.random.PCG64(n2))
    for have in (0, 1, 5, n2 // 2 - 3, n2 - 1, n2):
        keys = rng.integers(0, 1 << 62, size=have, dtype=np.uint64)
        if have > 4:
            keys[: have // 4] = keys[have // 4: 2 * (have // 4)]  # duplicates must not break it either
        buf = np.full(n2, np.uint64(0xFFFFFFFFFFFFFFFF))
        buf[:have] = keys
        out = bitonic_like_kernel(buf)
        assert np.array_equal(out, np.sort(buf))


def test_shared_bound_is_valid_and_ties_are_kept():
    """lists = disjoint subsets of the rows (as the (slice, group) lists of a query are); bound = k-th smallest of the first
    min(32, L) lists' current minima at ANY moment of the scan (prefixes of the lists)."""
    rng = np.random.Generator(np.random.PCG64(7))
    for trial in range(300):
        L = int(rng.integers(1, 48))
        k = int(rng.integers(1, min(L, 32) + 1))
        n = int(rng.integers(k, 400))
        d = rng.integers(0, 50, size=n)  # small alphabet: many equal distances
        owner = rng.integers(0, L, size=n)
        true_kth = np.sort(d)[k - 1]
        for frac in (0.1, 0.5, 1.0):  # how far each list has got
            mins = []
            for l in range(min(L, 32)):
                rows = d[owner == l]
                seen = rows[: max(0, int(np.ceil(len(rows) * frac)))]
                mins.append(seen.min() if len(seen) else None)
            fin = sorted(m for m in mins if m is not None)
            if len(fin) < k:
                continue  # the bound warp posts "no bound yet"
            bound = fin[k - 1]
            assert bound >= true_kth, (trial, L, k, frac)
            # the filter keeps d <= bound: every row of the true top-k (incl. all rows tied at the k-th distance) survives
            assert np.all(d[d <= true_kth] <= bound)


@pytest.mark.parametrize("G", [1, 2, 4, 8])
@pytest.mark.parametrize("two_slot", [False, True])
def test_code_ring_schedule_never_overwrites_a_stage_in_use(G, two_slot):
    """adc_scan_topk_kernel's per-warp code ring (pq_kernels.cu, ScanCfg): lane group h reads block Bg - h while group 0 is at
    block Bg.  Three-slot ring: stage s+1 is requested at the START of stage s into the slot that held stage s-2.  Two-slot ring
    (VAR bit 0, M = 32): stage s+1 is requested in the MIDDLE of stage s into the slot that held stage s-1.  At the moment of
    every request no lane group may still have blocks of the overwritten stage ahead of it, and every block must be read
    from the slot that holds its stage."""
    if two_slot and G != 8:
        pytest.skip("the two-slot ring is only instantiated for G = 8")
    stage_blocks = 16 if (G <= 4 or two_slot) else 8
    slots = 2 if two_slot else 3
    n_st = 9
    holds = {0: 0}  # slot -> stage (stage 0 is requested before the loop)
    for st in range(n_st):
        for bb in range(stage_blocks):
            Bg = st * stage_blocks + bb
            request_now = (bb == stage_blocks // 2) if two_slot else (bb == 0)
            if request_now and st + 1 < n_st:
                slot = (st + 1) % slots
                victim = holds.get(slot)
                if victim is not None:
                    for h in range(G):  # blocks this group has still to read, from its current one on
                        cur = Bg - h
                        assert cur >= (victim + 1) * stage_blocks, (G, two_slot, st, bb, h)
                holds[slot] = st + 1
            for h in range(G):
                b = Bg - h
                if b >= 0:
                    assert holds[(b // stage_blocks) % slots] == b // stage_blocks, (G, two_slot, st, bb, h)
