"""NumPy replay of the block path's plan (csrc/itn_block.cu) -- test infrastructure.

`itn_block_plan_export` hands out exactly what the CUDA kernels consume: per pass the block geometry (how a block of the
site tensor maps into shared memory), the operation list (mode products, closes, store) and the shared-memory fibre
tables.  This module executes that plan literally on flat NumPy buffers, block by block, so that the planner (strides,
paddings, tables, buffer allocation, divide-and-conquer order) is checked against the oracle on the CPU, without a GPU.
The device-side fragment mapping of the DMMA instructions is what the `-m gpu` parity tests cover.
"""
import ctypes as C

import numpy as np

from itn_b200._lib import check, i32, lib

NO_FIBRE = 0xFFFF
OP_MP, OP_CLOSE, OP_STORE = 0, 1, 2


class Pass:
    pass


class Plan:
    def __init__(self, dtype, d, chis, nverts=4096):
        self.cplx = np.dtype(dtype).kind == "c"
        self.d, self.chis = int(d), [int(c) for c in chis]
        z = len(self.chis)
        _, pc = i32(self.chis)
        n = C.c_int32()
        check(lib().itn_block_plan_export(1 if self.cplx else 0, self.d, z, pc, int(nverts), None, 0, C.byref(n)))
        self.supported = n.value > 0
        if not self.supported:
            return
        buf = np.zeros(n.value, dtype=np.int32)
        check(lib().itn_block_plan_export(1 if self.cplx else 0, self.d, z, pc, int(nverts),
                                          buf.ctypes.data_as(C.POINTER(C.c_int32)), n.value, C.byref(n)))
        it = iter(buf.tolist())
        self.KS, self.MT, self.h = next(it), next(it), next(it)
        self.passes = []
        for _ in range(3):
            p = Pass()
            p.nblk, p.nrows, p.rowlen, p.grow, p.gblk, p.nlev = [next(it) for _ in range(6)]
            p.lev_n = [next(it) for _ in range(4)]
            p.lev_s = [next(it) for _ in range(4)]
            p.PL, p.bufsz, p.nbuf, p.bulk, p.load_p, p.smem, p.excess, p.wl, p.nmodes = [next(it) for _ in range(9)]
            p.modes = [dict(zip(("chi", "S", "ntile", "tab", "slot"), [next(it) for _ in range(5)])) for _ in range(p.nmodes)]
            nops = next(it)
            p.ops = [tuple(next(it) for _ in range(4)) for _ in range(nops)]
            tl = next(it)
            p.table = np.array([next(it) for _ in range(tl)], dtype=np.int64)
            self.passes.append(p)

    # -- replay ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _row_pos(p, row):
        off, q = 0, row
        for l in range(p.nlev):
            off += (q % p.lev_n[l]) * p.lev_s[l]
            q //= p.lev_n[l]
        return off

    def _stage(self, p, src_flat, blk):
        """global (one plane) -> shared-memory image of block `blk` (length PL, padding left as NaN)"""
        out = np.full(p.PL, np.nan, dtype=src_flat.dtype)
        for row in range(p.nrows):
            g0 = blk * p.gblk + row * p.grow
            s0 = self._row_pos(p, row)
            assert np.all(np.isnan(out[s0:s0 + p.rowlen])), "rows overlap in shared memory"
            out[s0:s0 + p.rowlen] = src_flat[g0:g0 + p.rowlen]
        return out

    def _unstage(self, p, buf, dst_flat, blk):
        for row in range(p.nrows):
            g0 = blk * p.gblk + row * p.grow
            s0 = self._row_pos(p, row)
            dst_flat[g0:g0 + p.rowlen] = buf[s0:s0 + p.rowlen]

    def _fibres(self, p, m, warp=None):
        """Fibre bases of a mode; warp-local passes: tile ft belongs to warp ft % 8."""
        t = p.table[m["tab"]:m["tab"] + 8 * m["ntile"]]
        if warp is not None:
            t = t.reshape(-1, 8)[warp::8].reshape(-1)
        return t[t != NO_FIBRE]

    def run_pass(self, w, x_flat, pin_flat, msgs):
        """Returns (Wout flat or None, {slot: sum of the per-block partial messages})."""
        p = self.passes[w]
        wout = np.zeros_like(x_flat) if any(o[0] == OP_STORE for o in p.ops) else None
        outs = {}
        for blk in range(p.nblk):
            bufs = [None] * p.nbuf
            bufs[0] = self._stage(p, x_flat, blk)
            if p.load_p:
                bufs[1] = self._stage(p, pin_flat, blk)
            # warp-local passes: every warp runs the WHOLE operation list on its own tiles before the next warp starts
            # (no CTA barrier between operations on the device); the store waits for all of them
            for warp in (range(8) if p.wl else [None]):
                for (typ, src, dst, mode) in p.ops:
                    if typ == OP_MP:
                        m = p.modes[mode]
                        f = self._fibres(p, m, warp)
                        if len(f) == 0:
                            continue
                        msg = msgs[m["slot"]]
                        idx = f[:, None] + m["S"] * np.arange(m["chi"])[None, :]      # [fibre, a]
                        res = bufs[src][idx] @ msg                                   # out[f, b] = sum_a in[f, a] M[a, b]
                        assert not np.any(np.isnan(res)), "a mode product read an element nobody wrote"
                        if bufs[dst] is None:
                            bufs[dst] = np.full(p.PL, np.nan, dtype=x_flat.dtype)
                        elif dst == src and not p.wl:
                            bufs[dst] = bufs[dst].copy()
                        bufs[dst][idx] = res
                    elif typ == OP_CLOSE:
                        m = p.modes[mode]
                        f = self._fibres(p, m, warp)
                        if len(f) == 0:
                            continue
                        idx = f[:, None] + m["S"] * np.arange(m["chi"])[None, :]
                        part = bufs[src][idx].T @ bufs[0][idx].conj()
                        assert not np.any(np.isnan(part)), "a close read an element nobody wrote"
                        outs[m["slot"]] = outs.get(m["slot"], 0) + part
            for (typ, src, dst, mode) in p.ops:
                if typ == OP_STORE:
                    self._unstage(p, bufs[src], wout, blk)
        return wout, outs

    def sweep(self, a, msgs):
        """All outgoing (un-normalised) messages of a vertex with site tensor a[s, a_1..a_z] and incoming messages
        msgs[k]; returns a list indexed by bond slot."""
        x = np.asarray(a).reshape(-1, order="F")
        pten, o1 = self.run_pass(0, x, None, msgs)
        sten, o2 = self.run_pass(1, x, pten, msgs)
        _, o3 = self.run_pass(2, x, sten, msgs)
        assert not o1
        out = dict(o2)
        out.update(o3)
        return [out[k] for k in range(len(self.chis))]

    def check_tables(self):
        """Every fibre of every mode appears exactly once and inside the block."""
        for p in self.passes:
            nb = p.nrows * p.rowlen
            for m in p.modes:
                f = self._fibres(p, m)
                assert len(f) == len(set(f.tolist())) == nb // m["chi"], (len(f), nb // m["chi"])
                assert f.min() >= 0 and f.max() + m["S"] * (m["chi"] - 1) < p.PL
            assert p.nbuf * p.bufsz * 8 + (0 if p.wl else 8192) <= p.smem <= 227 * 1024
            if p.wl:
                for m in p.modes:
                    assert m["ntile"] % 8 == 0
