"""Host-side model of the persistent tile schedule of csrc/gemm_tma.cu (struct Sched + launch_tma_g): data-parallel
waves + stream-K tail.  The model restates the device logic line by line and checks its invariants for many problem
shapes -- every (tile, k-iteration) is computed exactly once, a CTA publishes at most one partial tile, every tile
has exactly one finisher, and a finisher only ever depends on LOWER-indexed CTAs whose contribution is the FIRST
stream-K segment they process (so waiting cannot deadlock, whatever the residency of the CTAs)."""
import itertools

import pytest


def launch_params(tiles, ipt, sms=148, min_iters=16):
    """launch_tma_g: grid size and the split into stream-K tiles / data-parallel waves."""
    G = sms
    if tiles < G:
        by_work = max(tiles * ipt // min_iters, tiles)
        if by_work < G:
            G = by_work
    full, rem = divmod(tiles, G)
    if rem == 0:
        T_sk, dp_waves = 0, full
    else:
        dp_waves = full - 1 if full > 0 else 0
        T_sk = tiles - dp_waves * G
    return G, T_sk, dp_waves


def segments(cta, G, T_sk, dp_waves, ipt, dp_first):
    """struct Sched: the (tile, it0, it1) sequence of one CTA."""
    U = T_sk * ipt
    u0, seg_end = U * cta // G, U * (cta + 1) // G
    sk = []
    while seg_end > u0:
        tile = (seg_end - 1) // ipt
        b = max(u0, tile * ipt)
        sk.append((tile, b - tile * ipt, seg_end - tile * ipt))
        seg_end = b
    dp = [(T_sk + w * G + cta, 0, ipt) for w in range(dp_waves)]
    return (dp + sk) if dp_first else (sk + dp), sk


CASES = [(t, i) for t, i in itertools.product([1, 2, 5, 127, 128, 147, 148, 149, 200, 295, 296, 297, 640, 1024, 5120],
                                              [1, 2, 16, 33, 128, 640])]


@pytest.mark.parametrize("tiles,ipt", CASES)
@pytest.mark.parametrize("dp_first", [False, True])
def test_schedule_covers_every_iteration_once_and_fixup_is_deadlock_free(tiles, ipt, dp_first):
    G, T_sk, dp_waves = launch_params(tiles, ipt)
    assert 1 <= G <= 148 and T_sk + dp_waves * G == tiles
    if T_sk:
        assert T_sk * ipt >= G                      # every CTA gets at least one stream-K iteration
    covered = {}
    finishers, contributors = {}, {}
    for cta in range(G):
        segs, sk = segments(cta, G, T_sk, dp_waves, ipt, dp_first)
        n_partial = 0
        for idx, (tile, it0, it1) in enumerate(segs):
            assert 0 <= tile < tiles and 0 <= it0 < it1 <= ipt
            for it in range(it0, it1):
                assert (tile, it) not in covered, "k-iteration computed twice"
                covered[(tile, it)] = cta
            if it1 < ipt:                           # contributor: publishes its accumulators
                n_partial += 1
                contributors.setdefault(tile, []).append((cta, idx, (tile, it0, it1) == sk[0]))
            else:                                   # owns the last k-iteration: stores the tile
                assert tile not in finishers
                finishers[tile] = (cta, it0)
        assert n_partial <= 1, "one partial-tile slot per CTA"
    assert len(covered) == tiles * ipt
    assert set(finishers) == set(range(tiles))
    U = T_sk * ipt
    for tile, (fin, it0) in finishers.items():
        # the device loop: for cc = cta-1 down while U*(cc+1)/G > tile*ipt
        waits = []
        cc = fin - 1
        while it0 > 0 and cc >= 0 and U * (cc + 1) // G > tile * ipt:
            waits.append(cc)
            cc -= 1
        got = sorted(c for c, _, _ in contributors.get(tile, []))
        assert sorted(waits) == got, "finisher waits for exactly the contributors of its tile"
        for c, idx, is_first_sk in contributors.get(tile, []):
            assert c < fin                          # only lower-indexed CTAs
            assert is_first_sk                      # ... and their contribution is their FIRST stream-K segment
