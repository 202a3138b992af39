"""Host model of the 8-lane warp layout of the long-chain kernels (csrc/b2g_dynamics.cuh::Grp<8>, csrc/b200gym.cu::thread_ids):
chains 0-3 of the warp's four environments in lanes 0-15, chains 4-7 chain-major in lanes 16-31.  The index algebra must be a
bijection and the three-stage reduction must pair the same chains in the same order as the plain xor butterfly of the
consecutive layout, so that every group sum is bit-identical (the GPU parity suite then checks the kernels themselves)."""
import numpy as np


def thread_ids(wl):
    """(environment in warp, chain) of physical lane wl -- thread_ids<8>."""
    return (wl >> 2, wl & 3) if wl < 16 else (wl & 3, 4 + ((wl - 16) >> 2))


def phys_lane(e, c):
    """Grp<8>::phys_lane."""
    return 4 * e + c if c < 4 else 16 + 4 * (c - 4) + e


def partners(wl):
    """source lanes of the three shuffles of Grp<8>::sum."""
    up = wl >> 4
    first = 4 * (wl & 3) + ((wl - 16) >> 2) if up else 16 + 4 * (wl & 3) + (wl >> 2)
    return first, wl ^ (8 if up else 2), wl ^ (4 if up else 1)


def test_layout_is_a_bijection():
    seen = set()
    for wl in range(32):
        e, c = thread_ids(wl)
        assert 0 <= e < 4 and 0 <= c < 8
        assert phys_lane(e, c) == wl
        seen.add((e, c))
    assert len(seen) == 32


def test_reduction_pairs_chains_like_the_xor_butterfly():
    for wl in range(32):
        e, c = thread_ids(wl)
        for stage, mask in zip(partners(wl), (4, 2, 1)):
            assert thread_ids(stage) == (e, c ^ mask)


def test_group_sums_are_bit_identical_to_the_consecutive_layout():
    rng = np.random.default_rng(0)
    for _ in range(200):
        x = (rng.standard_normal((4, 8)) * 10.0 ** rng.integers(-3, 4)).astype(np.float32)      # [env][chain]
        # consecutive layout: lane = 8 e + c, butterfly xor 4, 2, 1
        ref = x.copy()
        for mask in (4, 2, 1):
            ref = (ref + ref[:, [c ^ mask for c in range(8)]]).astype(np.float32)
        # split layout: values live at phys_lane(e, c); three shuffles with partners()
        v = np.zeros(32, np.float32)
        for e in range(4):
            for c in range(8):
                v[phys_lane(e, c)] = x[e, c]
        for k in range(3):
            v = (v + v[[partners(wl)[k] for wl in range(32)]]).astype(np.float32)
        for wl in range(32):
            e, c = thread_ids(wl)
            assert v[wl].tobytes() == ref[e, c].tobytes()
