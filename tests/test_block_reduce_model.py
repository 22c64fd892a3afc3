"""Model check of the warp stage of `block_sum_into` (elimaloc_b200/csrc/icp_kernels.cu): the reduce-SCATTER over the per-lane
accumulators — at the step with lane distance o a lane keeps one half of its values (the lower half if its bit o is clear), hands the
other half to its partner `lane ^ o` and adds what the partner hands over — must leave in lane l the warp's total of accumulator l,
formed by the SAME pairing tree as the xor-butterfly it replaced (distance 16, 8, 4, 2, 1; IEEE addition commutes), i.e. bit-identical
sums.  That is what lets the kernels change their reduction (31 exchanges per lane instead of 5 per accumulator) without moving a
single bit of JtJ / Jtr, and what `profiles/ab_chain.py` confirms on the B200 through the pose checksums of the two builds.
The kernel's select / exchange sequence is restated lane by lane in numpy float64."""
import numpy as np
import pytest


def butterfly(x):
    v = x.copy()
    for o in (16, 8, 4, 2, 1):
        v = np.array([v[lane] + v[lane ^ o] for lane in range(32)])
    return v


def reduce_scatter(x, nacc):
    """x: (32 lanes, nacc) -> per lane the value the kernel leaves in acc[0]"""
    acc = np.zeros((32, 32))
    acc[:, :nacc] = x
    # first step: accumulator k pairs with k + 16; accumulators >= nacc do not exist (the kernel substitutes the constant 0.0)
    new = acc.copy()
    for lane in range(32):
        up = bool(lane & 16)
        partner = lane ^ 16
        for k in range(16):
            hi_mine = acc[lane, k + 16] if k + 16 < nacc else 0.0
            hi_theirs = acc[partner, k + 16] if k + 16 < nacc else 0.0
            keep = hi_mine if up else acc[lane, k]
            sent_by_partner = acc[partner, k] if bool(partner & 16) else hi_theirs
            new[lane, k] = keep + sent_by_partner
    acc = new
    for o in (8, 4, 2, 1):
        new = acc.copy()
        for lane in range(32):
            up = bool(lane & o)
            partner = lane ^ o
            for k in range(o):
                keep = acc[lane, k + o] if up else acc[lane, k]
                sent_by_partner = acc[partner, k] if bool(partner & o) else acc[partner, k + o]
                new[lane, k] = keep + sent_by_partner
        acc = new
    return acc[:, 0]


@pytest.mark.parametrize("nacc", [18, 29])  # AccSize<P2P> and AccSize<GICP / VGICP / AVGICP>
def test_reduce_scatter_equals_the_butterfly_bit_for_bit(nacc):
    rng = np.random.default_rng(nacc)
    for trial in range(50):
        x = rng.standard_normal((32, nacc)) * 10.0 ** rng.integers(-12, 12, (32, nacc))  # wild magnitudes: any other summation order shows
        if trial % 5 == 0:
            x[rng.random((32, nacc)) < 0.5] = 0.0                                           # lanes without a correspondence
        ref = butterfly(x)
        assert all((ref[lane] == ref[0]).all() for lane in range(32))                       # the butterfly leaves the same bits in every lane
        got = reduce_scatter(x, nacc)
        assert (got[:nacc] == ref[0]).all()
        assert np.array_equal(np.signbit(got[:nacc]), np.signbit(ref[0]))                   # (also the sign of a zero)


def test_the_model_detects_another_summation_order():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((32, 18)) * 10.0 ** rng.integers(-12, 12, (32, 18))
    assert not (x.sum(axis=0) == butterfly(x)[0]).all()


def _expand_p2p(a, k):
    """expand_p2p of icp_kernels.cu: canonical slot k (upper JtJ row-major, Jtr, residual, count) of the 18 structured P2P sums"""
    fixed = {0: a[0], 6: a[0], 11: a[0], 4: a[3], 5: -a[2], 8: -a[3], 10: a[1], 12: a[2], 13: -a[1], 27: a[16], 28: a[17]}
    if k in fixed:
        return fixed[k]
    if 15 <= k <= 20:
        return a[4 + k - 15]
    if 21 <= k <= 26:
        return a[10 + k - 21]
    return 0.0


def test_p2p_scatter_table_is_expand_p2p_read_backwards():
    """After the reduce-scatter lane l holds accumulator l; kP2pScatter tells every lane which canonical slots of the row it writes.
    The table is parsed out of the kernel source: every slot 0..28 written exactly once with the value (and sign) expand_p2p gives it."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "elimaloc_b200", "csrc", "icp_kernels.cu")).read()
    blk = src[src.index("__constant__ uint32_t kP2pScatter[32] = {"):src.index("#undef ELM_SC")]
    ents = re.findall(r"ELM_SC\((0x[0-9a-f]+|\d+), (0x[0-9a-f]+|\d+), (0x[0-9a-f]+|\d+), (\d+)\)", blk)
    assert len(ents) == 32
    a = np.random.default_rng(0).standard_normal(18)
    row = {}
    for lane, (s0, s1, s2, f) in enumerate((int(p, 0), int(q, 0), int(r, 0), int(g)) for p, q, r, g in ents):
        v = a[lane] if lane < 18 else 123.456          # (lanes >= 18 hold nothing of value)
        w = 0.0 if f & 1 else v
        w1 = -w if f & 2 else w
        for slot, val in ((s0, w), (s1, w1), (s2, w)):
            if slot != 0xFF:
                assert slot not in row, f"slot {slot} written twice"
                row[slot] = val
    assert sorted(row) == list(range(29))
    for k in range(29):
        e = _expand_p2p(a, k)
        assert row[k] == e and np.signbit(row[k]) == np.signbit(e), k
