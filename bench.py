#!/usr/bin/env python
"""bench.py -- BP edge-message updates/s (BASELINE.json metric) on synthetic lattices.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one synchronous BP sweep (every directed edge updated once from the pre-sweep messages,
abstractbeliefpropagationcache.jl:294-308) over the whole lattice.  Default workload: the configuration
north_star quotes the target on and that fits one GPU -- 64x64 square-lattice PEPS, chi=16, d=2,
ComplexF64 (BASELINE.json configs[3]; 16,128 directed messages and 1.034e12 algorithmic flops per sweep).

  value     device-resident: K sweeps in one `update(bpc; maxiter=K)` call, psi and messages already in HBM,
            timed with CUDA events on the library's stream, max over ranks.
  e2e       the same metric through the public API with HOST buffers: every step uploads psi from pinned
            host memory (BeliefPropagationCache(psi)), runs one sweep and downloads every message.
  roofline  FP64 tensor (DMMA) bound: algorithmic flops of SURVEY.md 8(d) (F_msg = 8*z*d*chi^(z+1)) over
            the device time of the contraction kernels, against the FP64 peak measured on this pool
            (profiles/r1_fp64_peak.json, tools/fp64_peak.cu; MEASURED_PEAKS.json has no FP64 entry).
  cpu_baseline / --impl reference
            the restated reference algorithm (oracle/itn_oracle.py: NumPy + OpenBLAS, one message at a time,
            updated_message :225-239) on the box's host cores over a bounded sample of the same workload.
            The Julia reference itself cannot run here (no Julia runtime).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "itensornetworks.jl_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (dims, chi, dtype, BASELINE.json config index); dims = "heavyhex" -> the 127-qubit Eagle graph
    "grid64x64_chi16_c128": ((64, 64), 16, np.complex128, 3),
    "grid32x32_chi8_c128": ((32, 32), 8, np.complex128, 1),
    "grid4x4_chi2_f64": ((4, 4), 2, np.float64, 0),
    "cubic16_chi6_c128": ((16, 16, 16), 6, np.complex128, 4),
    "heavyhex127_chi32_c128": ("heavyhex", 32, np.complex128, 2),
    "grid16x16_chi16_c128": ((16, 16), 16, np.complex128, None),
}
OTHER_CONFIGS = ("grid32x32_chi8_c128", "heavyhex127_chi32_c128", "cubic16_chi6_c128")  # BASELINE configs 2, 3, 5
METRIC = "bp_edge_message_updates_per_s"
UNIT = "updates/s"
SEED = 1234


def make_graph(E, dims):
    return E.heavy_hex_eagle() if dims == "heavyhex" else E.named_grid(dims)


def vertex_seed(v):
    """One generator per vertex (seed (+) v): every rank count N draws the same bytes for vertex v, so the N = 1, 2, 4, 8
    runs time identical inputs and their logZ / sampled messages can be compared."""
    return SEED * 1000003 + int(v)


def make_psi(torch, graph, chi, dtype, d, mine):
    """Synthetic psi in pinned host memory: iid N(0,1) / CN(0,1) entries, F-ordered [site, bonds...] per vertex, vertex v
    drawn from its own generator.  A rank only materialises the vertices it owns (the others alias one scratch block)."""
    cplx = np.dtype(dtype).kind == "c"
    comps = 2 if cplx else 1
    full = [d * chi ** graph.degree(v) for v in range(graph.nv)]
    sizes = [full[v] if mine[v] else 0 for v in range(graph.nv)]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    scratch = np.zeros(max(full) * comps)
    host = torch.empty(int(offs[-1]) * comps, dtype=torch.float64)
    try:
        host = host.pin_memory()
    except Exception:
        pass
    gen = torch.Generator()
    for v in range(graph.nv):
        if mine[v]:
            gen.manual_seed(vertex_seed(v))
            torch.randn(full[v] * comps, generator=gen, out=host[int(offs[v]) * comps:int(offs[v + 1]) * comps])
    if cplx:
        host.mul_(2.0 ** -0.5)
    hnp = host.numpy()
    tensors = []
    for v in range(graph.nv):
        flat = hnp[int(offs[v]) * comps:int(offs[v + 1]) * comps] if mine[v] else scratch[:full[v] * comps]
        if cplx:
            flat = flat.view(np.complex128)
        tensors.append(np.ndarray((d,) + (chi,) * graph.degree(v), dtype=dtype, buffer=flat, order="F"))
    return tensors, host, int(offs[-1]) * comps * 8


def sampled_parity(E, bpc, tensors, graph, seq, mine, nsample=16):
    """Full-size parity, outside every timed region: download the messages into `nsample` sampled vertices, run ONE more
    synchronous sweep through the production kernels, download the sampled outgoing messages and compare each with the
    oracle's updated_message (oracle/itn_oracle.py, abstractbeliefpropagationcache.jl:225-239) evaluated on the
    downloaded inputs.  Returns the largest relative Frobenius error (the oracle is the checker, never the thing timed)."""
    from oracle import itn_oracle as O
    cand = [v for v in range(graph.nv) if mine[v]]
    rng = np.random.default_rng(99)
    zmax = max(graph.degree(v) for v in cand)
    bulk = [v for v in cand if graph.degree(v) == zmax]
    rim = [v for v in cand if graph.degree(v) != zmax]
    picks = list(rng.choice(bulk, size=min(len(bulk), max(1, nsample - 4)), replace=False))
    if rim:
        picks += list(rng.choice(rim, size=min(len(rim), 4), replace=False))
    incoming = {}
    for v in picks:
        incoming[v] = [bpc.message((graph.other(e, v), v)) for e in graph.inc[v]]
    E.update(bpc, maxiter=1, edge_sequence=seq, inplace=True)
    worst = 0.0
    nmsg = 0
    for i, v in enumerate(picks):
        k = i % graph.degree(v)
        w = graph.other(graph.inc[v][k], v)
        want = O.updated_message_local(np.asarray(tensors[v]), incoming[v], k)
        got = bpc.message((v, w))
        worst = max(worst, float(np.linalg.norm(got - want) / np.linalg.norm(want)))
        nmsg += 1
    return worst, nmsg


def algorithmic_flops_per_sweep(graph, chi, d, cplx):
    """F_msg = c*z*d*chi^(z+1), c = 2 (Float64) / 8 (ComplexF64), summed over directed edges."""
    c = 8.0 if cplx else 2.0
    tot = 0.0
    for v in range(graph.nv):
        z = graph.degree(v)
        tot += z * (c * z * d * float(chi) ** (z + 1))
    return tot


def measured_traffic(workload, world):
    """DRAM bytes of the sweep's contraction kernels per sweep and GPU, from the newest committed ncu --set full capture
    that names this workload (profiles/*traffic.json: {"workload", "dram_bytes_per_sweep", "source", "commit"}).  Not
    measured live: ncu cannot run inside the timed region.  Returns (bytes or None, provenance string)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    try:
        names = sorted(f for f in os.listdir(pdir) if f.endswith("traffic.json"))
    except OSError:
        names = []
    for f in names:
        try:
            j = json.load(open(os.path.join(pdir, f)))
        except Exception:
            continue
        if j.get("workload") == workload:
            best = (float(j["dram_bytes_per_sweep"]) / world,
                    "profiles/%s: %s, captured at commit %s" % (f, j.get("source", "ncu --set full"), j.get("commit", "?")))
    return best if best else (None, None)


def fp64_peak():
    """FP64 roofline denominators measured on this pool (tools/fp64_peak.cu): cuBLAS ZGEMM (the roofline peak) and the
    raw DMMA issue rate (denominator of the executed-flops pipe fraction)."""
    path = os.path.join(ROOT, "profiles", "r1_fp64_peak.json")
    try:
        j = json.load(open(path))
        return {"burst": float(j["zgemm_tflops_burst"]), "sustained": float(j["zgemm_tflops_sustained"]),
                "dmma_pipe": float(j["dmma_pipe_tflops"]),
                "source": "profiles/r1_fp64_peak.json (cuBLAS ZGEMM 4096^3 measured on this pool's B200; "
                          "MEASURED_PEAKS.json has no FP64 entry)"}
    except Exception:
        return {"burst": 37.0, "sustained": 37.0, "dmma_pipe": 37.2, "source": "fallback: 148 SM x 64 FMA/clk x 1.965 GHz"}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(self.device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
            except ValueError:
                continue
            for nm, val in zip(names, r[5:9]):
                if val.strip().lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_sample(dims, chi, dtype, budget_s, seed=1234, threads=None):
    """Time the restated reference algorithm on a bounded sample: a (<=6)^n corner patch of the same
    lattice with the same chi/dtype, message updates one at a time in the reference's sequential
    order (only the interior degree mix differs slightly; flops per update are reported).
    threads: BLAS thread count for this sample (None = whatever the pool has)."""
    from oracle import itn_oracle as O
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    if dims == "heavyhex":
        g = O.heavy_hex_eagle_graph()
        pdims = ("heavyhex",)
    else:
        pdims = tuple(min(int(x), 6) for x in dims)
        g = O.grid_graph(pdims)
    net = O.random_network(g, chi, dtype=dtype, seed=seed)
    msgs = O.identity_messages(net)
    seq = O.default_edge_sequence(g)
    # prefer full-degree vertices first so the sample matches the bulk of the big lattice
    zmax = max(g.degree(v) for v in range(g.nv))
    seq = [e for e in seq if g.degree(e[0]) == zmax] + [e for e in seq if g.degree(e[0]) != zmax]

    def body():
        O.updated_message(net, msgs, *seq[0])  # warm-up (BLAS thread pool)
        n = 0
        t0 = time.perf_counter()
        while True:
            v, w = seq[n % len(seq)]
            msgs[(v, w)] = O.updated_message(net, msgs, v, w)
            n += 1
            if time.perf_counter() - t0 > budget_s or n >= 4 * len(seq):
                break
        return n, time.perf_counter() - t0

    if threads is not None and threadpool_limits is not None:
        with threadpool_limits(limits=int(threads), user_api="blas"):
            n, dt = body()
    else:
        n, dt = body()
    return n, dt, f"{n} sequential updated_message calls (vertices of degree {zmax} first) on a {'x'.join(map(str, pdims))} patch, chi={chi}, {np.dtype(dtype).name}"


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] + [1])
    except Exception:
        return os.cpu_count() or 1


def max_blas_threads():
    """Host cores the BLAS pool may use: under torchrun OMP_NUM_THREADS is forced to 1, so ask the OS."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline_both(dims, chi, dtype, budget_s):
    """the CPU port at 1 BLAS thread and at every host core, in every line; `value` is the better
    of the two (the most favourable statement for the reference)."""
    nmax = max_blas_threads()
    n1, t1, sample = cpu_reference_sample(dims, chi, dtype, budget_s / 2, threads=1)
    na, ta, _ = cpu_reference_sample(dims, chi, dtype, budget_s / 2, threads=nmax)
    v1, va = n1 / t1, na / ta
    best_threads = 1 if v1 >= va else nmax
    return {"value": max(v1, va), "unit": UNIT, "cores": best_threads, "kind": "port", "sample": sample,
            "by_threads": {"1": v1, str(nmax): va}, "host_cores": nmax,
            "blas": "NumPy + OpenBLAS (threadpoolctl limits); per-update rate on a corner patch, optimistic for the "
                    "reference: no per-update cache copy (a5), no per-call contraction-order search (a4)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dims, chi, dtype, cfg = WORKLOADS[args.workload]
    nmax = max_blas_threads()
    for _ in range(args.warmup):
        cpu_reference_sample(dims, chi, dtype, 0.5, threads=nmax)
    per_step_budget = max(1.0, min(20.0, 90.0 / max(args.steps, 1)))
    # every host thread the arm can use -- and, because one BLAS thread is often faster at these matrix sizes, the
    # single-thread rate too; the line's value is the better of the two
    rates = {}
    sample = ""
    for th in sorted({1, nmax}):
        tot_n, tot_t = 0, 0.0
        for _ in range(args.steps):
            n, dt, sample = cpu_reference_sample(dims, chi, dtype, per_step_budget / len({1, nmax}), threads=th)
            tot_n += n
            tot_t += dt
        rates[th] = (tot_n / tot_t, tot_t)
    cores = max(rates, key=lambda th: rates[th][0])
    val, tot_t = rates[cores]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "c128" if np.dtype(dtype).kind == "c" else "f64",
        "data": "synthetic", "config": {"workload": args.workload, "baseline_config_index": cfg,
                                        "note": "restated reference algorithm (NumPy+OpenBLAS oracle port), not ITensorNetworks.jl: no Julia runtime in this image"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + f", per step, x{args.steps} steps",
                         "by_threads": {str(th): r[0] for th, r in rates.items()}, "host_cores": nmax},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def ms_estimate_short(args, world):
    """True when the timed region is likely shorter than the clock sampler's period."""
    return args.steps * 30.0 / max(world, 1) < 400.0


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPU cores NVML lists as local to its GPU, so that the pinned host buffers are allocated on
    the GPU's NUMA node and the host->device copies of 8 ranks do not cross the socket interconnect (torchrun does not
    bind ranks).  Best effort: any failure leaves the affinity untouched."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cores = [c for c in cores if c in allowed]
        if cores:
            os.sched_setaffinity(0, cores)
        return len(cores)
    except Exception:
        return 0


def time_sweeps(E, torch, bpc, seq, stream, steps, warmup):
    """Device time (CUDA events on the library's stream) of `steps` synchronous sweeps in one update() call."""
    E.update(bpc, maxiter=max(warmup, 3), edge_sequence=seq, inplace=True)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        E.update(bpc, maxiter=steps, edge_sequence=seq, inplace=True)
        ev1.record(stream)
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / steps


def bench_other_config(E, torch, ctx, stream, name, peak_tf):
    """BASELINE.json configs 2, 3, 5 at full size on one GPU: ms per synchronous sweep, updates/s, algorithmic FP64 rate
    and the sampled oracle parity of the same sweep kernels."""
    dims, chi, dtype, cfg = WORKLOADS[name]
    d = 2
    cplx = np.dtype(dtype).kind == "c"
    graph = make_graph(E, dims)
    mine = [True] * graph.nv
    tensors, host, _ = make_psi(torch, graph, chi, dtype, d, mine)
    psi = E.ITensorNetwork(graph, tensors, dtype)
    # the schedule in wire format, made once (a driver calls update() with the same schedule every step)
    seq = E.prepare_sequence(E.parallel_edge_sequence(graph))
    bpc = E.BeliefPropagationCache(psi, ctx=ctx)
    steps = 20 if graph.nv < 2000 else 5
    ms = time_sweeps(E, torch, bpc, seq, stream, steps, 3)
    tm = bpc.last_timing()
    err, nmsg = sampled_parity(E, bpc, tensors, graph, seq, mine, nsample=8)
    logz = E.logscalar(bpc)
    bpc.close()
    fl = algorithmic_flops_per_sweep(graph, chi, d, cplx)
    tf = fl / (ms * 1e-3) / 1e12
    return {"baseline_config_index": cfg, "ms_per_sweep": ms, "updates_per_s": 2 * graph.ne / (ms * 1e-3),
            "messages_per_sweep": 2 * graph.ne, "algorithmic_tflops": tf, "frac_of_fp64_peak": tf / peak_tf,
            "contract_ms_per_sweep": tm["contract_ms"] / steps,
            "parity": {"max_rel_err": err, "messages_checked": nmsg, "logz_re": float(np.real(logz))}}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import itn_b200 as E

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libitn_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dims, chi, dtype, cfg = WORKLOADS[args.workload]
    cplx = np.dtype(dtype).kind == "c"
    d = 2
    graph = make_graph(E, dims)
    stream = torch.cuda.Stream()
    ctx = E.Context(local_rank, stream=stream.cuda_stream)
    if args.path is not None:
        ctx.set_path(args.path)
    if world > 1:
        E.init_distributed(ctx, rank, world)  # the library's own NCCL communicator (boundary messages)
    owner = E.partition_vertices(graph, world, kind=args.partition) if world > 1 else None

    # synthetic psi in pinned host memory, one generator per vertex: the same bytes at every N
    mine = [owner is None or owner[v] == rank for v in range(graph.nv)]
    tensors, host, h2d = make_psi(torch, graph, chi, dtype, d, mine)
    comps = 2 if cplx else 1
    psi = E.ITensorNetwork(graph, tensors, dtype)
    # the schedule in wire format, made once (a driver calls update() with the same schedule every step)
    seq = E.prepare_sequence(E.parallel_edge_sequence(graph))
    n_updates = 2 * graph.ne
    flops_sweep = algorithmic_flops_per_sweep(graph, chi, d, cplx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ---------------------------------------------------------------------
    bpc = E.BeliefPropagationCache(psi, ctx=ctx, owner=owner, dist=(rank, world) if world > 1 else None)
    # clocks are sampled from the warm-up sweeps to the end of the timed region (the GPU is under the same load in both;
    # nvidia-smi needs ~100 ms to deliver its first sample, longer than a short timed region)
    sampler = ClockSampler(local_rank)
    sampler.start()
    E.update(bpc, maxiter=max(args.warmup, 3), edge_sequence=seq, inplace=True)
    barrier()
    if ms_estimate_short(args, world):
        E.update(bpc, maxiter=20, edge_sequence=seq, inplace=True)  # extra untimed sweeps so that the sampler sees load
        barrier()
    l0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        E.update(bpc, maxiter=args.steps, edge_sequence=seq, inplace=True)
        ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - l0
    tm = bpc.last_timing()
    if world > 1:
        t = torch.tensor([ms, tm["contract_ms"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, tm["contract_ms"] = float(t[0]), float(t[1])
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt[0])
    value = n_updates * args.steps / (ms * 1e-3)

    # ---- parity at full size (outside the timed region): sampled updated_message vs the oracle, and logZ ----------
    # The state after warm-up + timed sweeps is the same at every N (same bytes, synchronous schedule), so logZ must
    # agree across the N = 1, 2, 4, 8 lines and the sampled messages must match the oracle on every rank.
    parity = None
    if args.parity:
        err, nmsg = sampled_parity(E, bpc, tensors, graph, seq, mine, nsample=16)
        if world > 1:
            t = torch.tensor([err], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            err = float(t[0])
            nt = torch.tensor([nmsg], device="cuda", dtype=torch.int64)
            dist.all_reduce(nt)
            nmsg = int(nt[0])
        logz = E.logscalar(bpc)
        parity = {"max_rel_err": err, "messages_checked": nmsg, "tolerance": 1e-10,
                  "logz_re": float(np.real(logz)), "logz_im": float(np.imag(logz)),
                  "sweeps_before_logz": max(args.warmup, 3) + (20 if ms_estimate_short(args, world) else 0) + args.steps + 1,
                  "what": "sampled messages of one more production sweep vs oracle.updated_message_local on the downloaded "
                          "inputs (max relative Frobenius error over all ranks); logZ = logscalar(bpc) of the common state"}

    # ---- end-to-end arm: host buffers in, host buffers out, every step ------------------------------
    e2e_steps = max(1, min(args.steps, 3))
    n_stored = n_updates if owner is None else sum(
        2 for (u, v) in graph.edges if owner[u] == rank or owner[v] == rank)
    d2h = n_stored * chi * chi * comps * 8
    out_host = torch.empty(n_stored * chi * chi * comps, dtype=torch.float64).pin_memory()
    phases = {"construct": 0.0, "update": 0.0, "download": 0.0, "close": 0.0}

    def e2e_step():
        # defer_upload: the constructor registers the pinned host tensors, update() streams them in chunks and runs the
        # sweep of chunk c while chunk c + 1 crosses PCIe (itn_net_set_tensors / ITN_HOST_DEFERRED)
        t0 = time.perf_counter()
        c2 = E.BeliefPropagationCache(psi, ctx=ctx, owner=owner, dist=(rank, world) if world > 1 else None, defer_upload=True)
        t1 = time.perf_counter()
        E.update(c2, maxiter=1, edge_sequence=seq, inplace=True)
        t2 = time.perf_counter()
        c2.messages_into(out_host.numpy())
        t3 = time.perf_counter()
        c2.close()
        t4 = time.perf_counter()
        for k, dt in zip(("construct", "update", "download", "close"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            phases[k] += dt

    e2e_step()  # untimed warm-up: the second network's device allocations grow the stream-ordered pool once
    barrier()
    for k in phases:
        phases[k] = 0.0
    per_step = []
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ta = time.perf_counter()
        e2e_step()
        per_step.append(time.perf_counter() - ta)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
        hb = torch.tensor([h2d, d2h], device="cuda", dtype=torch.int64)
        dist.all_reduce(hb)
        h2d, d2h = int(hb[0]), int(hb[1])
    e2e_val = n_updates / e2e_s

    # ---- simple-update gates/s (second half of BASELINE.json's metric) ------------------------------------
    su = None
    if args.gates and dims != "heavyhex":
        su = bench_simple_update(E, bpc, graph, chi, d, dtype, stream, torch, dist if world > 1 else None, seq)

    peak = fp64_peak()
    # ---- BASELINE configs 2, 3, 5 on the same GPU (single-GPU runs of the default workload only) -------------------
    other = None
    if args.other_configs and world == 1 and rank == 0 and args.workload == "grid64x64_chi16_c128":
        bpc.close()
        other = {}
        for name in OTHER_CONFIGS:
            if name == args.workload:
                continue
            try:
                other[name] = bench_other_config(E, torch, ctx, stream, name, peak["sustained"])
            except Exception as ex:  # a failing side config must not hide the headline line
                other[name] = {"error": str(ex)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # per-GPU roofline of the contraction kernels (message-update DMMA / FMA kernels), live CUDA events inside the library
    contract_ms_per_sweep = tm["contract_ms"] / max(args.steps, 1)
    flops_per_gpu = flops_sweep / world
    achieved = flops_per_gpu / (contract_ms_per_sweep * 1e-3) / 1e12 if contract_ms_per_sweep > 0 else None
    # flops the kernels actually issue: degree-4 chi=16 vertices run 12 of the 16 algorithmic units (itn_fast.cu header)
    shared = sum(0.25 * 4 * (8.0 if cplx else 2.0) * 4 * d * float(chi) ** 5 for v in range(graph.nv)
                 if graph.degree(v) == 4 and chi == 16)
    executed_tflops = ((flops_sweep - shared) / world / (contract_ms_per_sweep * 1e-3) / 1e12
                       if contract_ms_per_sweep > 0 and (tm.get("path", "auto") != 1 and args.path in (None, 0)) else None)
    traffic, traffic_src = measured_traffic(args.workload, world)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c128" if cplx else "f64", "data": "synthetic",
        "config": {"workload": args.workload, "baseline_config_index": cfg, "lattice": list(dims) if dims != "heavyhex" else "heavy-hex Eagle, 127 sites",
                   "chi": chi, "d": d,
                   "schedule": "synchronous sweep (one group per directed edge)", "messages_per_sweep": n_updates,
                   "algorithmic_flops_per_sweep": flops_sweep, "path": tm.get("path", "auto"),
                   "psi": "iid CN(0,1), one torch generator per vertex (seed %d*1000003 + v): identical bytes at every N" % SEED,
                   "l2": "inputs larger than L2 (psi %.2f GB per sweep)" % (h2d / 1e9) if h2d > 2e8 else "working set fits L2; no flush (latency-bound config)",
                   "parallelism": ("graph partition (%s) x%d, NCCL boundary messages" % (args.partition, world)) if world > 1 else "single GPU"},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak["sustained"], "unit": "TFLOP/s",
                     "frac": (achieved / peak["sustained"]) if achieved else None,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "message-update contraction kernels (per sweep, per GPU)", "peak_source": peak["source"],
                     "contract_ms_per_sweep": contract_ms_per_sweep,
                     "executed_tflops": executed_tflops,
                     "executed_frac_of_dmma_pipe": executed_tflops / peak["dmma_pipe"] if executed_tflops else None,
                     "dmma_pipe_peak": peak["dmma_pipe"],
                     "note": "achieved = ALGORITHMIC flops of SURVEY.md 8(d) (8*z*d*chi^(z+1) per message, i.e. four independent "
                             "message updates per degree-4 vertex = 16 units of d*chi^5 MACs) / device time of the contraction kernels; the "
                             "DMMA path shares partial absorptions between the four outputs of a vertex and EXECUTES 12 units, which is why "
                             "the algorithmic rate can exceed the measured ZGEMM peak; executed_tflops is the rate of the flops actually "
                             "issued (3/4 of the algorithmic count on degree-4 chi=16 vertices), dmma_pipe_peak the measured DMMA issue "
                             "rate of tools/fp64_peak.cu. FP64 peak, not bf16"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "s_per_step": per_step,
                "host_phases_s_per_step_rank0": {k: v / e2e_steps for k, v in phases.items()}, "what": "BeliefPropagationCache(psi in pinned host memory, defer_upload=True) + update(maxiter=1) [host->device copy pipelined with the sweep] + download of all messages"},
        "gpu_launches": launches, "clocks": clocks,
    }
    if parity is not None:
        line["parity"] = parity
    if su is not None:
        su["frac_of_fp64_peak_per_gpu_gemm_only"] = su["algorithmic_tflops_gemm_only"] / peak["sustained"] / world
        line["simple_update"] = su
    if other is not None:
        line["other_configs"] = other
    if args.cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_both(dims, chi, dtype, args.cpu_budget)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def gate_flops(graph, chi, d, cplx):
    """SURVEY.md 8(d) per site: F_gemm = c*(2(z-1)+d)*d*chi^(z+1) (absorb, un-absorb, rebuild) and
    F_qr = (c/2)*4*d^2*chi^(z+1); one gate touches two sites.  Summed over every edge of the lattice.
    Returns (F_gemm, F_qr): the engine replaces the QR by a Cholesky factorisation of the bond environment
    (csrc/itn_linalg.cu), so F_qr is algorithmic-only -- it is reported separately and never credited as executed."""
    c = 8.0 if cplx else 2.0
    fg = fq = 0.0
    for (u, v) in graph.edges:
        for w in (u, v):
            z = graph.degree(w)
            fg += c * (2 * (z - 1) + d) * d * float(chi) ** (z + 1)
            fq += (c / 2) * 4 * d * d * float(chi) ** (z + 1)
    return fg, fq


def bench_simple_update(E, bpc, graph, chi, d, dtype, stream, torch, dist=None, seq=None):
    """One Trotter step = one two-site gate on every edge, applied as vertex-disjoint colour layers
    (apply(o, psi; envs = BP messages, maxdim = chi, cutoff = nothing), src/apply.jl:97-146), in place on the device.
    Every bond is truncated from d^2 chi = 64 back to chi = 16 singular values, so the lattice stays on the chi = 16 kernels.
    Two measurements:
      bp_gauged        north_star's apply(...; cache_update_kwargs): ONE itn_apply_layers call per Trotter step; after every
                       colour layer the gated messages are set to diag(singular values) (msg_mode = 1, the Vidal-gauge
                       fixed point of the pair) and `bp_sweeps_per_layer` synchronous BP sweeps refresh every environment
                       before the next layer -- what a TEBD driver does.
      trivial_gauges   the layers alone, gated messages reset to the identity and no BP in between (round-1 number)."""
    rng = np.random.default_rng(7)
    m = rng.standard_normal((d * d, d * d)) + (1j * rng.standard_normal((d * d, d * d)) if np.dtype(dtype).kind == "c" else 0)
    h = (m + m.conj().T) / 2
    w, v = np.linalg.eigh(h)
    # exp(-i tau H): a real-time (unitary) Trotter gate; repeated steps keep the state well scaled.  (The slow-down after a
    # few BP-gauged steps that round 2 first blamed on the gate was the phase drift of the BP messages, DESIGN.md
    # section 3b / tests/test_phase_drift.py: fixed in the engine, k_commit keeps the Hermitian part.)
    if np.dtype(dtype).kind == "c":
        gate = ((v * np.exp(-0.05j * w)) @ v.conj().T).astype(dtype).reshape(d, d, d, d)
    else:
        gate = ((v * np.exp(-0.05 * w)) @ v.conj().T).astype(dtype).reshape(d, d, d, d)
    layers = E.edge_coloring(graph)
    fg, fq = gate_flops(graph, chi, d, np.dtype(dtype).kind == "c")

    def sync_max(dt):
        if dist is not None:  # every rank applies the same layers to its part of the lattice: the slowest rank counts
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        return dt

    # ---- trivial gauges: layers only ------------------------------------------------------------------------------
    work = bpc.copy()
    # the gate layers of a Trotter step in wire format, packed once (a TEBD driver applies the same layers every step)
    prepared = [E.prepare_layer(work, [gate] * len(layer), [graph.edges[e] for e in layer]) for layer in layers]
    # warm-up: full untimed Trotter steps (grow the stream-ordered memory pool to its steady state: every layer
    # allocates the new site tensors of its 2 x |layer| vertices in one slab and releases the old ones)
    for _ in range(3):  # three untimed steps: the slab ping-pong of the layers needs three slabs before it stops allocating
        for pl in prepared:
            E.apply_layer(pl, work, maxdim=chi, cutoff=None)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    ngates = 0
    terr = 0.0
    per_layer = []
    for layer, pl in zip(layers, prepared):
        ta = time.perf_counter()
        info = E.apply_layer(pl, work, maxdim=chi, cutoff=None)
        per_layer.append(round(1e3 * (time.perf_counter() - ta), 2))  # host-side time of the call: the rebuild of a layer overlaps the next call, the step total is synchronised
        ngates += len(layer)
        terr = max(terr, float(np.max(info["truncation_error"])))
    torch.cuda.synchronize()
    dt = sync_max(time.perf_counter() - t0)
    work.close()
    out = {"gates_per_s": ngates / dt, "ms_per_trotter_step": 1e3 * dt, "gates": ngates, "colour_layers": len(layers), "ms_per_layer": per_layer,
           "max_truncation_error": terr, "algorithmic_tflops_gemm_only": fg / dt / 1e12,
           "algorithmic_tflops_gemm_plus_qr": (fg + fq) / dt / 1e12,
           "mode": "trivial_gauges (gated messages reset to identity, no BP between layers)",
           "note": "wall clock (max over ranks) incl. the host round trip per layer (new bond dimensions are read back); "
                   "F_gemm = absorb + un-absorb + rebuild gemms of SURVEY.md 8(d); F_qr is algorithmic only (the engine runs a "
                   "Cholesky of the bond environment instead) and is not part of the fraction-of-peak figure"}
    # ---- BP-gauged Trotter step in one library call -----------------------------------------------------------------
    k_bp = 2
    work = bpc.copy()
    lay = [([gate] * len(layer), [graph.edges[e] for e in layer]) for layer in layers]
    try:
        for _ in range(2):
            E.tebd_step(work, lay, maxdim=chi, cutoff=None, msg_mode=1, bp_maxiter=k_bp, edge_sequence=seq)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        info = E.tebd_step(work, lay, maxdim=chi, cutoff=None, msg_mode=1, bp_maxiter=k_bp, edge_sequence=seq)
        torch.cuda.synchronize()
        dtg = sync_max(time.perf_counter() - t0)
        out["bp_gauged"] = {"gates_per_s": ngates / dtg, "ms_per_trotter_step": 1e3 * dtg, "bp_sweeps_per_layer": k_bp,
                            "msg_mode": "diag(singular values) on the gated edge, then BP sweeps",
                            "max_truncation_error": float(np.max(info["truncation_error"])),
                            "call": "itn_apply_layers (one call per Trotter step: 4 colour layers x (gates + %d synchronous BP sweeps))" % k_bp}
    except Exception as ex:
        out["bp_gauged"] = {"error": str(ex)[:200]}
    work.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="grid64x64_chi16_c128", choices=sorted(WORKLOADS))
    ap.add_argument("--path", type=int, default=None, help="0 = auto (DMMA fast path), 1 = generic kernels only")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-gates", dest="gates", action="store_false", help="skip the simple-update gates/s section")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-parity", dest="parity", action="store_false", help="skip the sampled oracle parity check")
    ap.add_argument("--no-other-configs", dest="other_configs", action="store_false",
                    help="skip the BASELINE configs 2, 3, 5 section (single-GPU runs only)")
    ap.add_argument("--partition", default="auto", choices=["auto", "strips", "bricks"],
                    help="graph partition for N > 1: strips along the first axis or near-cubic bricks (auto = fewest cut edges)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
