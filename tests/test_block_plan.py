"""CPU checks of the block path's planner (csrc/itn_block.cu): the exported geometry, operation lists and fibre tables are
replayed in NumPy (tests/block_emulator.py) and every outgoing message must equal the oracle's updated_message
(abstractbeliefpropagationcache.jl:225-239).  No device needed."""
import time

import numpy as np
import pytest

from oracle import itn_oracle as O

from block_emulator import Plan

SIGS = [
    # (dtype, d, chis, bucket size)
    (np.complex128, 2, [6] * 6, 2744),         # BASELINE config 5 bulk vertices
    (np.complex128, 2, [6] * 5, 1176),
    (np.complex128, 2, [6] * 3, 8),
    (np.complex128, 2, [8] * 4, 900),          # config 2
    (np.complex128, 2, [8] * 3, 120),
    (np.complex128, 2, [8] * 2, 4),
    (np.complex128, 2, [32] * 3, 36),          # config 3
    (np.complex128, 2, [32] * 2, 89),
    (np.complex128, 2, [16] * 4, 3844),        # config 4 (tile path in production; the block path must agree)
    (np.complex128, 2, [16] * 3, 248),
    (np.float64, 2, [2] * 4, 4),               # config 1
    (np.float64, 2, [2] * 2, 4),
    (np.float64, 3, [3, 5, 4], 7),             # odd extents, d = 3: no 16-byte rows
    (np.complex128, 2, [4, 16, 2, 16], 50),    # mixed extents during TEBD growth
    (np.complex128, 2, [1, 4, 1, 4], 50),
    (np.float64, 2, [7, 3, 12], 20),
    (np.complex128, 1, [5, 5, 5, 5], 10),
]


@pytest.mark.parametrize("dtype,d,chis,nverts", SIGS, ids=[f"{np.dtype(s[0]).name}_d{s[1]}_" + "x".join(map(str, s[2])) for s in SIGS])
def test_plan_replay_matches_oracle(dtype, d, chis, nverts):
    t0 = time.time()
    plan = Plan(dtype, d, chis, nverts)
    assert plan.supported
    assert time.time() - t0 < 20.0  # planning is host work done once per signature
    plan.check_tables()
    rng = np.random.default_rng(len(chis) * 100 + chis[0])
    cplx = np.dtype(dtype).kind == "c"

    def rnd(shape):
        x = rng.standard_normal(shape)
        return (x + 1j * rng.standard_normal(shape)) / np.sqrt(2) if cplx else x
    a = rnd((d,) + tuple(chis)).astype(dtype)
    msgs = [rnd((c, c)).astype(dtype) for c in chis]
    got = plan.sweep(a, msgs)
    for k in range(len(chis)):
        want = O.updated_message_local(a, msgs, k, normalize=False)
        assert np.linalg.norm(got[k] - want) < 1e-12 * np.linalg.norm(want), k


def test_unsupported_signatures_have_no_plan():
    assert not Plan(np.complex128, 2, [64, 64, 64]).supported     # extent above 32
    assert not Plan(np.complex128, 2, [6]).supported              # degree 1: nothing to share
    assert not Plan(np.complex128, 2, [16] * 6).supported         # a group of three 16s does not fit shared memory


def test_config5_plan_shape():
    # z = 6, chi = 6: two CTAs per SM on every pass, no avoidable bank conflicts, 22 units of work per vertex
    p = Plan(np.complex128, 2, [6] * 6, 2744)
    assert (p.KS, p.MT, p.h) == (3, 1, 3)
    for w in range(3):
        assert p.passes[w].smem <= 113 * 1024, (w, p.passes[w].smem)
    units = sum(1 for w in range(3) for o in p.passes[w].ops if o[0] in (0, 1))
    assert units == 22
    # in-place operation lists: one buffer for pass 1, three for passes 2 / 3 (X, the partially absorbed tensor, one
    # temporary); pass 2 takes the half-size cross-warp scratch and fits three CTAs per SM (228 KB / 3 - 1 KB - static)
    assert [q.nbuf for q in p.passes] == [1, 3, 3]
    assert p.passes[1].smem <= 228 * 1024 // 3 - 1024 - 2304
    assert not any(q.wl for q in p.passes)   # chi = 6: a warp's slice would hold one parity of the site index


def test_config2_plan_is_warp_local():
    # z = 4, chi = 8: every pass deals whole batch slices to the warps (exact tiles, no excess wavefronts planned), mode
    # products in place; the replay above runs such plans warp by warp
    p = Plan(np.complex128, 2, [8] * 4, 900)
    assert all(q.wl for q in p.passes)
    assert [q.nbuf for q in p.passes] == [1, 3, 3]
    assert all(q.excess == 0 for q in p.passes)
    for q in p.passes:
        for src_dst in [(o[1], o[2]) for o in q.ops if o[0] == 0]:
            assert src_dst[0] == src_dst[1] or src_dst[1] >= 2 or q is p.passes[0]
