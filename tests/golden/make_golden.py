"""Generates tests/golden/*.npz from the oracle (run from the repo root: python tests/golden/make_golden.py).

These are REGRESSION pins of the oracle's own output, not pins against the Julia reference: the reference
cannot run in this image (no Julia runtime) and ships no golden vectors.  They guard the
oracle against accidental change and give the GPU tests a fixed set of bytes that does not depend on the
NumPy version's random stream."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import itn_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(name, dims, chi, dtype, iters, gate_edge):
    g = O.grid_graph(dims)
    net = O.random_network(g, chi, dtype=dtype, seed=1234)
    seq = O.parallel_edge_sequence(g)
    msgs, _, _ = O.bp_update(net, O.identity_messages(net), seq=seq, groups=O.synchronous_groups(seq), maxiter=iters)
    seqs = O.default_edge_sequence(g)
    msgs_seq, _, _ = O.bp_update(net, O.identity_messages(net), seq=seqs, maxiter=iters)
    zv, ze = O.region_scalars(net, msgs)
    ez = np.array([O.expect1(net, msgs, v, O.PAULI_Z) for v in range(g.nv)])
    zz = np.array([O.expect2(net, msgs, e, O.PAULI_Z, O.PAULI_Z) for e in range(g.ne)])
    gate = O.random_unitary(4, seed=11, dtype=dtype).reshape(2, 2, 2, 2)
    new, info = O.simple_update_bp(net, msgs, gate_edge, gate, maxdim=chi, cutoff=1e-12)
    u, v = g.edges[gate_edge]
    pair = np.tensordot(new.tensors[u], new.tensors[v], axes=([1 + g.slot(u, gate_edge)], [1 + g.slot(v, gate_edge)]))
    d = {"dims": np.array(dims), "chi": chi, "iters": iters, "gate_edge": gate_edge, "gate": gate,
         "zv": zv, "ze": ze, "logz": np.array(O.logscalar(net, msgs)), "expect_z": ez, "expect_zz": zz,
         "svals": info["svals"], "truncerr": info["truncerr"], "newdim": info["newdim"], "pair": pair,
         "seq_sync": np.array(seq), "seq_sequential": np.array(seqs)}
    for v_, t in enumerate(net.tensors):
        d[f"T{v_}"] = t
    for (a, b), m in msgs.items():
        d[f"M_{a}_{b}"] = m
    for (a, b), m in msgs_seq.items():
        d[f"S_{a}_{b}"] = m
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "written")


if __name__ == "__main__":
    case("grid4x4_chi2_f64", (4, 4), 2, np.float64, 6, 7)        # BASELINE config 1 shape
    case("grid3x3_chi3_c128", (3, 3), 3, np.complex128, 6, 5)
    case("cubic2x2x2_chi2_c128", (2, 2, 2), 2, np.complex128, 5, 3)
