import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/itensornetworks.jl_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import itn_b200 as E
from oracle import itn_oracle as O
from util import make_pair, rel_err
g = O.grid_graph((3,3,3))
for dtype in (np.complex128, np.float64):
    net, psi = make_pair(g, 2, dtype)
    seq = O.default_edge_sequence(g)
    for mode in (0,1,2):
        c = E.Context(0); c.set_path(mode)
        bpc = E.BeliefPropagationCache(psi, ctx=c)
        msgs = O.identity_messages(net)
        out=[]
        for it in range(4):
            msgs, _, _ = O.bp_update(net, msgs, seq=seq, maxiter=1)
            E.update(bpc, maxiter=1, edge_sequence=seq, inplace=True)
            worst = max(rel_err(bpc.message(k), m) for k, m in msgs.items())
            # conditioning proxy: smallest |message norm before normalisation|
            out.append(worst)
        print(np.dtype(dtype).name, 'mode', mode, ['%.2e'%w for w in out])
