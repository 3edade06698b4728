"""Block path (csrc/itn_block.cu) against the oracle: synchronous sweeps on graphs whose vertices exercise every kernel
instance (real / complex, packed and aligned complex stacking, 1 / 2 / 4 row tiles, odd extents, mixed extents)."""
import numpy as np
import pytest

import itn_b200 as E
from oracle import itn_oracle as O

from util import assert_messages_close, make_pair, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10

CASES = [
    # name, graph, chi (scalar or per edge), dtype
    ("grid5_chi8_c", lambda: O.grid_graph((5, 5)), 8, np.complex128),
    ("grid5_chi8_r", lambda: O.grid_graph((5, 5)), 8, np.float64),
    ("grid4_chi6_c", lambda: O.grid_graph((4, 4)), 6, np.complex128),
    ("grid4x3_chi12_c", lambda: O.grid_graph((4, 3)), 12, np.complex128),
    ("grid4x3_chi12_r", lambda: O.grid_graph((4, 3)), 12, np.float64),
    ("cubic3_chi4_c", lambda: O.grid_graph((3, 3, 3)), 4, np.complex128),
    ("cubic3_chi3_r", lambda: O.grid_graph((3, 3, 3)), 3, np.float64),
    ("cubic3x3x2_chi6_c", lambda: O.grid_graph((3, 3, 2)), 6, np.complex128),
    ("heavyhex_chi8_c", O.heavy_hex_eagle_graph, 8, np.complex128),
    ("heavyhex_chi20_c", O.heavy_hex_eagle_graph, 20, np.complex128),
    ("chain5_chi32_c", lambda: O.chain_graph(5), 32, np.complex128),
    ("grid3_chi16_r", lambda: O.grid_graph((3, 3)), 16, np.float64),
    ("grid3_chi32_r", lambda: O.grid_graph((3, 3)), 32, np.float64),
    ("grid3x2_ragged_c", lambda: O.grid_graph((3, 2)), [2, 3, 4, 2, 3, 2, 4], np.complex128),
    ("grid4_mixed_c", lambda: O.grid_graph((4, 4)), [1, 4, 16, 2, 8, 4, 1, 16, 4, 2, 8, 8, 4, 16, 2, 1, 4, 4, 8, 2, 16, 4, 1, 8][:24], np.complex128),
]


@pytest.mark.parametrize("name,mk,chi,dtype", CASES, ids=[c[0] for c in CASES])
def test_block_sweeps_match_oracle(name, mk, chi, dtype):
    g = mk()
    if isinstance(chi, list):
        chi = (chi * 4)[:len(g.edges)]
    net, psi = make_pair(g, chi, dtype)
    seq = O.parallel_edge_sequence(g)
    msgs, _, diff_o = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=3, tol=0.0)
    ctx = E.Context(0)
    c0 = ctx.path_counts()
    info = {}
    bpc = E.update(E.BeliefPropagationCache(psi, ctx=ctx), maxiter=3, tol=0.0, edge_sequence=[[e] for e in seq], info=info)
    c1 = ctx.path_counts()
    assert c1[1] - c0[1] > 0, "the block path did not run"
    assert_messages_close(bpc, msgs, TOL)
    assert abs(info["mean_diff"] - diff_o) < 1e-12
    # second opinion on the device: the shape-generic kernels
    c2 = E.Context(0)
    c2.set_path(1)
    b2 = E.update(E.BeliefPropagationCache(psi, ctx=c2), maxiter=3, edge_sequence=[[e] for e in seq])
    assert c2.path_counts()[1] == 0
    for k in msgs:
        assert rel_err(bpc.message(k), b2.message(k)) < 1e-12


def test_block_path_takes_boundary_vertices_next_to_the_tile_path():
    # 64 x 64 style lattice in small: interior degree-4 chi=16 vertices on the tile path, the rim (degree 2 and 3) on the
    # block path, nothing left for the shape-generic kernels
    g = O.grid_graph((5, 5))
    net, psi = make_pair(g, 16, np.complex128)
    seq = O.parallel_edge_sequence(g)
    msgs, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=2)
    ctx = E.Context(0)
    bpc = E.update(E.BeliefPropagationCache(psi, ctx=ctx), maxiter=2, edge_sequence=[[e] for e in seq])
    tile, block, generic = ctx.path_counts()
    assert tile == 2 * 9 * 4 and block == 2 * (len(seq) - 36) and generic == 0, (tile, block, generic)
    assert_messages_close(bpc, msgs, TOL)
