"""CPU tier: the two pieces of index logic the long-row Jacobi kernel (csrc/jacobi.cu, jacobi_persistent_cols_kernel)
relies on, restated in Python and checked exhaustively:

* the compile-time pairings of a round - (i, i ^ s) inside each block for s = 1..BR-1 and (q, BR + (q ^ s)) across the
  two blocks for s = 0..BR-1 - meet every pair of the 2 BR rows exactly once, and the pairs of one step are disjoint;
* the halving butterfly that reduces the 4 BR partial sums of a step across the 32 lanes of a warp leaves sum number
  lane >> 1 (16 sums) or lane (32 sums) in v[0] of every lane, with 16 / 31 shuffles (5 per sum otherwise).
"""
import itertools

import pytest


def xor_pair_low(BR, s, t):
    for c in range(BR):
        if c < (c ^ s):
            if t == 0:
                return c
            t -= 1
    raise AssertionError


@pytest.mark.parametrize('BR', [4, 8])
def test_pairings_of_a_round_cover_every_pair_once(BR):
    met = []
    for s in range(1, BR):                       # inside the blocks (once per sweep, round 0)
        step = []
        for q in range(BR):
            blk = BR if q >= BR // 2 else 0
            lo = xor_pair_low(BR, s, q % (BR // 2))
            step.append((blk + lo, blk + (lo ^ s)))
        rows = [r for pr in step for r in pr]
        assert sorted(rows) == list(range(2 * BR))            # disjoint: every row in exactly one pair of the step
        met += step
    inside = {tuple(sorted(pr)) for pr in met}
    want = {pr for blk in (0, BR) for pr in itertools.combinations(range(blk, blk + BR), 2)}
    assert inside == want and len(met) == len(want)
    cross = []
    for s in range(BR):                          # block I against block J (every round)
        step = [(q, BR + (q ^ s)) for q in range(BR)]
        rows = [r for pr in step for r in pr]
        assert sorted(rows) == list(range(2 * BR))
        cross += step
    assert set(cross) == {(i, BR + j) for i in range(BR) for j in range(BR)} and len(cross) == BR * BR


@pytest.mark.parametrize('V', [16, 32])
def test_halving_butterfly_reduces_every_sum(V):
    # lane l starts with v[i] = a distinct integer weight, so that the sum over lanes identifies the slot
    lanes = [[(l + 1) * 1000 + i for i in range(V)] for l in range(32)]
    total = [sum(lanes[l][i] for l in range(32)) for i in range(V)]
    shuffles = 0
    off, c = 16, V // 2
    while c >= 1:
        new = [row[:] for row in lanes]
        for l in range(32):
            upper = (l & off) != 0
            for i in range(c):
                partner = lanes[l ^ off]
                p_upper = ((l ^ off) & off) != 0
                send_partner = partner[i] if p_upper else partner[i + c]
                keep = lanes[l][i + c] if upper else lanes[l][i]
                new[l][i] = keep + send_partner
        lanes = new
        shuffles += c
        off >>= 1
        c >>= 1
    if V == 16:
        lanes = [[lanes[l][0] + lanes[l ^ 1][0]] + lanes[l][1:] for l in range(32)]
        shuffles += 1
    for l in range(32):
        idx = l if V == 32 else (l >> 1)
        assert lanes[l][0] == total[idx]
    assert shuffles == (31 if V == 32 else 16)       # against 5 V for one butterfly per sum
