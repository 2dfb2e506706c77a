"""bench.py -- gene alignments / second of the Genes2Genes hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4] [--impl reference]

Workload (every N, strong scaling): BASELINE.json configs[3] "C4" -- 20,000 genes x (100,000 reference +
100,000 query cells), 50 interpolation points, the shape the north-star target is quoted on; it fits one
B200 (16 GB of f32 expression values).  Genes are sharded over the N ranks in contiguous blocks
(genes2genes_b200/distributed.py); every rank generates ITS block of the synthetic matrices on its GPU
(genes2genes_b200/synthetic.py: a gene's values do not depend on the sharding).

One "step" = one pass of the hot path over the whole workload: K0 weight tables, K1a pilot means, K1m
tensor-core interpolation, fused K3 + K4 (fix-up, cost, DP, traceback) on every rank's block, then the NCCL
gather of the packed per-gene result records to rank 0 (overlapped with the next step's kernels).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs already in HBM); `e2e` is the
same metric through the public API -- `Main.RefQueryAligner(adata_ref, adata_query, gene_list, 50,
distributed=True)` + `align_all_pairs()` -- from pinned HOST matrices to host result objects on rank 0,
with the host->device copies, the time sort, the gather and the device->host result read inside the timed
region.  `--impl reference` times the reference package's own CPU path (baseline/_ref through
baseline/ref_runner.py; the oracle port if the package is absent) on the box's host cores.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gene_alignments_per_sec"
UNIT = "genes/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # defaults long enough (0.6 s of back-to-back steps at N = 1) for the GPU to settle under its power cap: at C4 the
    # first ~10 steps run at the 1965 MHz boost clock (5.49 ms per step), the steady state is 1800 MHz (5.96 ms)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--dropout", type=float, default=0.9)
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C2 / C3 sub-records (N = 1)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--moments", default="auto", choices=["auto", "mma", "ffma"])
    ap.add_argument("--cpu-genes-per-core", type=int, default=2, help="genes per host core in the bounded CPU sample")
    ap.add_argument("--gather", choices=("peer", "nccl"), default="peer",
                    help="N > 1: how the result records reach rank 0 (peer: copies into rank 0's window; nccl: dist.gather)")
    ap.add_argument("--nccl-max-ctas", type=int, default=0, help="N > 1: CTA cap of the NCCL communicator (0 = NCCL's default)")
    ap.add_argument("--sm-reserve", type=int, default=0,
                    help="SMs the persistent interpolation kernel leaves free (experiments with overlapped collectives)")
    return ap.parse_args()


def shapes(name):
    from genes2genes_b200 import synthetic
    return synthetic.CONFIGS[name]


def config_dict(name, dropout, n_gpus, extra=None):
    G, nr, nq, T = shapes(name)
    cfg = {"workload": "%s: %d genes x (%d ref + %d query cells), %d interpolation points, dense f32 [cells x genes], "
                       "synthetic log1p counts, dropout parameter %.2f" % (name, G, nr, nq, T, dropout),
           "n_genes": G, "genes_per_gpu": -(-G // n_gpus), "n_ref_cells": nr, "n_query_cells": nq,
           "n_interpolation_points": T,
           "parallelism": ("genes sharded over %d GPUs in contiguous blocks, one gather of the result records "
                           "to rank 0 per step" % n_gpus) if n_gpus > 1 else "single GPU",
           "l2": "inputs larger than L2 (%.1f GB of expression values per rank per step); no flush"
                 % (4.0 * (nr + nq) * (-(-G // n_gpus)) / 1e9)}
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(Xr, Xq, tr, tq, T):
    from oracle import g2g_oracle as o
    orr, oq = o.sort_cells(tr), o.sort_cells(tq)
    _CPU.update(o=o, Xr=Xr[orr], Xq=Xq[oq], taur=tr[orr], tauq=tq[oq], eps=o.epsilon_table(T),
                Wr=o.weight_matrix(tr[orr], T), Wq=o.weight_matrix(tq[oq], T))


def _cpu_gene(g):
    c = _CPU
    r = c["o"].align_gene_faithful(c["Xr"][:, g].astype(np.float64), c["Xq"][:, g].astype(np.float64), c["taur"],
                                   c["tauq"], c["Wr"], c["Wq"], c["eps"])
    return r["alignment_str"]


def _one_blas_thread():
    try:
        from threadpoolctl import threadpool_limits
        _CPU["_limit"] = threadpool_limits(1)
    except Exception:
        pass


def cpu_port_rate(Xr, tr, Xq, tq, T, n_procs):
    """genes/s of the oracle port (oracle/g2g_oracle.py, per-gene loop like Main.py:449-455)."""
    n = Xr.shape[1]
    import multiprocessing as mp
    _cpu_init(Xr, Xq, tr, tq, T)
    if n_procs <= 1:
        _one_blas_thread()
        t0 = time.perf_counter()
        for g in range(n):
            _cpu_gene(g)
        return n / (time.perf_counter() - t0)
    ctx = mp.get_context("fork")
    with ctx.Pool(n_procs, initializer=_one_blas_thread) as pool:
        pool.map(_cpu_gene, range(min(n_procs, n)))  # spin the workers up
        t0 = time.perf_counter()
        pool.map(_cpu_gene, range(n), chunksize=max(1, n // (4 * n_procs)))
        dt = time.perf_counter() - t0
    return n / dt


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_sample(name, dropout, n_genes, device=None):
    """`n_genes` genes of the workload at FULL cell counts as host arrays (evenly spaced gene indices)."""
    from genes2genes_b200 import synthetic
    G, nr, nq, T = shapes(name)
    sel = np.unique(np.linspace(0, G - 1, min(G, n_genes)).astype(int))
    if device is None:
        Xr, tr, Xq, tq, _ = synthetic.make_pair(G, nr, nq, dropout=dropout, seed=100)
        return Xr[:, sel].copy(), tr, Xq[:, sel].copy(), tq, T, sel
    import torch
    params = synthetic.gene_params(G, 100)
    cols_r, cols_q = [], []
    tr = tq = None
    for g in sel:   # one 256-gene generator block per sampled gene: cheap on the GPU
        xr, tr = synthetic.make_side_device(nr, params, 1, 0, device, int(g), int(g) + 1, dropout=dropout)
        xq, tq = synthetic.make_side_device(nq, params, 2, 1, device, int(g), int(g) + 1, dropout=dropout, shift=0.08)
        cols_r.append(xr[:, 0].cpu().numpy())
        cols_q.append(xq[:, 0].cpu().numpy())
    torch.cuda.synchronize()
    return np.stack(cols_r, axis=1), tr, np.stack(cols_q, axis=1), tq, T, sel


def cpu_reference_measure(name, dropout, genes_per_core, device=None):
    """The CPU baseline of a record: the reference package's own RefQueryAligner + align_all_pairs on a bounded
    gene sample of the workload at full cell counts -- sequential (1 core) and concurrent=True with one
    process per core (Main.py:438-457) -- or, when the package is not present, the oracle port run the same
    two ways.  Returns the cpu_baseline object (value = the all-cores rate)."""
    from baseline import ref_runner
    cores = os.cpu_count() or 1
    n_conc = max(8, genes_per_core * cores)
    n_seq = max(4, min(8, n_conc))
    t0 = time.perf_counter()
    Xr, tr, Xq, tq, T, sel = cpu_sample(name, dropout, n_conc, device)
    out = {"unit": UNIT, "cores": cores, "cpu_model": cpu_model()}
    if ref_runner.reference_available():
        dt_seq, _ = ref_runner.time_reference(Xr[:, :n_seq], tr, Xq[:, :n_seq], tq, T)
        dt_conc, _ = ref_runner.time_reference(Xr, tr, Xq, tq, T, concurrent=True, n_processes=cores)
        out.update(kind="reference", value=len(sel) / dt_conc, sequential_value=n_seq / dt_seq,
                   what="unmodified genes2genes v0.2.0 (baseline/_ref) RefQueryAligner + align_all_pairs, "
                        "constructor included; value = concurrent=True with %d processes, sequential_value = 1 core"
                        % cores)
    else:
        seq = cpu_port_rate(Xr[:, :n_seq], tr, Xq[:, :n_seq], tq, T, 1)
        conc = cpu_port_rate(Xr, tr, Xq, tq, T, cores)
        out.update(kind="port", value=conc, sequential_value=seq,
                   what="oracle port (oracle/g2g_oracle.py align_gene_faithful): the reference package is not present")
    out["sample"] = ("%d genes (concurrent) / %d genes (sequential) of the workload at full cell counts, %.0f s"
                     % (len(sel), n_seq, time.perf_counter() - t0))
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    device = None
    try:
        import torch
        if torch.cuda.is_available():
            device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    except Exception:
        device = None
    G, nr, nq, T = shapes(args.workload)
    if device is None and G * (nr + nq) > 2e9:
        emit({"impl": "reference", "unavailable": "no GPU to generate the %s sample and the host generator needs "
              "%.0f GB" % (args.workload, 4e-9 * G * (nr + nq))})
        return
    rates, base = [], None
    t0 = time.perf_counter()
    n_steps = max(1, min(args.steps, 2))   # every step is a bounded sample; the run stays within minutes
    for _ in range(n_steps):
        base = cpu_reference_measure(args.workload, args.dropout, args.cpu_genes_per_core, device)
        rates.append(base["value"])
    wall = time.perf_counter() - t0
    rate = float(np.mean(rates))
    base["value"] = rate
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": n_steps, "warmup": 0, "ms_per_step": 1e3 * wall / n_steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, args.dropout, args.gpus),
            "cpu_baseline": base,
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this configuration."""
    path = os.path.join(ROOT, "profiles", "k1_dram_traffic.json")
    try:
        return json.load(open(path)).get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def fused_ncu(name):
    """issue / occupancy counters of the fused fix-up + DP launch from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "fused_launch_ncu.json")
    try:
        return json.load(open(path)).get(name, {})
    except Exception:
        return {}


def result_digest(host):
    """sha256 over every gene's integer results and optimal cost (bitwise shard invariant)."""
    h = hashlib.sha256()
    for k in ("align_len", "align_str", "flags", "nnz", "opt_cost", "l_states"):
        h.update(np.ascontiguousarray(host[k]).tobytes())
    return h.hexdigest()[:24]


class DeviceWorkload:
    """One rank's gene block of a workload on its GPU: packed (time-sorted) matrices + the engine."""

    def __init__(self, name, dropout, lo, hi, dev, moments="auto", capacity=None, keep_raw=False):
        import torch
        from genes2genes_b200 import engine, synthetic
        from genes2genes_b200.TimeSeriesPreprocessor import TrajectoryInterpolator
        G, nr, nq, T = shapes(name)
        self.T, self.lo, self.hi, self.n_ref, self.n_query = T, lo, hi, nr, nq
        Xr, tr, Xq, tq, genes = synthetic.make_pair_device(G, nr, nq, dev, lo, hi, dropout=dropout, seed=100)
        self.tr, self.tq, self.genes = tr, tq, genes
        self.raw = (Xr, Xq) if keep_raw else None
        tir = TrajectoryInterpolator(tr, T, adaptive_kernel=False).run()
        tiq = TrajectoryInterpolator(tq, T, adaptive_kernel=False).run()
        Xrd = engine.pack_dense(Xr, tir.order, None, dev)
        Xqd = engine.pack_dense(Xq, tiq.order, None, dev)
        del Xr, Xq
        self.engine = engine.AlignmentEngine(Xrd, tir.cell_pseudotimes, Xqd, tiq.cell_pseudotimes, hi - lo, T,
                                             tir.kernel_denominators(), tiq.kernel_denominators(), device=dev,
                                             moments=moments, capacity=capacity)
        torch.cuda.synchronize()


def capture_graph(run):
    """The kernel launches of ``run()`` as a CUDA graph (None when capture is not possible)."""
    import torch
    try:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        return g
    except Exception as exc:  # capture is an optimisation; fall back to eager launches
        print("CUDA graph capture failed (%s); running eagerly" % exc, file=sys.stderr)
        torch.cuda.synchronize()
        return None


def time_small_workload(name, dropout, dev, steps, warmup):
    """Secondary record (N = 1): device-resident step time and the dominant kernel of a small configuration,
    L2 flushed between steps."""
    import torch
    G, nr, nq, T = shapes(name)
    w = DeviceWorkload(name, dropout, 0, G, dev, keep_raw=True)
    eng = w.engine
    eng.run()
    torch.cuda.synchronize()
    graph = capture_graph(eng.run)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    run = graph.replay if graph is not None else eng.run
    for _ in range(warmup):
        flush.fill_(1)
        run()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(2)
        a.record(); run(); b.record()
    k1 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in k1:
        flush.fill_(3)
        eng.time_moments(a, b)
    fu = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in fu:
        flush.fill_(4)
        eng.time_fused(a, b)
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in k1]))
    fu_ms = float(np.mean([a.elapsed_time(b) for a, b in fu]))
    peak, _ = measured_peak()
    alg = 4.0 * (nr + nq) * G + 8.0 * (nr + nq)
    out = {"workload": config_dict(name, dropout, 1)["workload"], "ms_per_step": ms, "value": G / (ms * 1e-3),
           "unit": UNIT, "steps": steps, "l2": "flushed (256 MiB write) between steps", "cuda_graph": graph is not None,
           "kernel": "moments_mma_kernel" if eng.variant == "mma" else "moments_kernel", "kernel_ms": k1_ms,
           "roofline_frac": alg / (k1_ms * 1e-3) / 1e9 / peak, "fused_fixup_dp_ms": fu_ms}
    # end to end through the public API from pinned host matrices (original cell order) to host result objects
    from genes2genes_b200 import Main
    from genes2genes_b200.anndata_lite import AnnDataLite
    Xr, Xq = w.raw
    Xr_pin = torch.empty(Xr.shape, dtype=torch.float32, pin_memory=True); Xr_pin.copy_(Xr)
    Xq_pin = torch.empty(Xq.shape, dtype=torch.float32, pin_memory=True); Xq_pin.copy_(Xq)
    torch.cuda.synchronize()
    w.raw = None
    ar = AnnDataLite(Xr_pin[:, :G], w.tr, w.genes)
    aq = AnnDataLite(Xq_pin[:, :G], w.tq, w.genes)
    e2e = []
    for k in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        al = Main.RefQueryAligner(ar, aq, w.genes, T, verbose=False)
        al.align_all_pairs()
        n = sum(a.alignment_str.count("M") for a in al.results)
        torch.cuda.synchronize()
        if k:
            e2e.append((time.perf_counter() - t0) * 1e3)
        del al
    out["e2e"] = {"value": G / (float(np.mean(e2e)) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(e2e)),
                  "ms_each": [round(v, 3) for v in e2e], "h2d_bytes_per_step": int(4 * (nr + nq) * G + 8 * (nr + nq))}
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    from genes2genes_b200 import Main, distributed, engine
    from genes2genes_b200.anndata_lite import AnnDataLite

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the gather's NCCL kernels share the GPU with the next step's kernels: cap their CTAs at the number of SMs
        # the persistent interpolation kernel leaves free, so that the two really run side by side
        opts = dist.ProcessGroupNCCL.Options()
        if args.nccl_max_ctas > 0:
            opts.config.max_ctas = args.nccl_max_ctas
            opts.config.min_ctas = 1
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    G, nr, nq, T = shapes(args.workload)
    lo, hi = distributed.shard_range(G, rank, world)
    cap = distributed.max_shard(G, world)
    all_genes = ["g%d" % i for i in range(G)]

    # ---- device-resident arm ------------------------------------------------------------------------
    w = DeviceWorkload(args.workload, args.dropout, lo, hi, dev, moments=args.moments, capacity=cap, keep_raw=True)
    eng = w.engine
    # pinned host copies of this rank's blocks in the ORIGINAL cell order: the e2e arm's inputs
    Xr_pin = torch.empty(w.raw[0].shape, dtype=torch.float32, pin_memory=True)
    Xq_pin = torch.empty(w.raw[1].shape, dtype=torch.float32, pin_memory=True)
    Xr_pin.copy_(w.raw[0]); Xq_pin.copy_(w.raw[1])
    torch.cuda.synchronize()
    w.raw = None
    torch.cuda.empty_cache()

    eng.sm_reserve = args.sm_reserve
    eng.run()   # first run sets kernel attributes / loads modules
    torch.cuda.synchronize()
    graph = graph_moments = graph_fused = None
    if not args.no_graph:
        if world == 1:
            graph = capture_graph(eng.run)
        else:   # two graphs: the gather of the previous step is launched between them (see one_step)
            graph_moments, graph_fused = capture_graph(eng.run_moments), capture_graph(eng.run_fused)
            graph = graph_moments if (graph_moments is not None and graph_fused is not None) else None
    # N > 1, the gather.  Default: every rank copies its packed records into rank 0's peer-mapped window with one
    # device-to-device copy over NVLink on a side stream (a copy engine, no kernel), under the next step's kernels.
    # --gather nccl (or a box that cannot map the window): the NCCL gather of step k-1 runs on a side stream next to
    # step k's fused fix-up / DP launch -- not next to the interpolation kernel, whose persistent CTAs fill every SM
    # (an NCCL kernel resident first delays the CTAs of the SMs it holds, and with them the whole statically
    # scheduled pass); the records of a step are first copied to one of two send buffers.
    window = None
    if world > 1 and args.gather == "peer":
        window = distributed.peer_window(eng.packed.numel(), dev)
    use_nccl = world > 1 and window is None
    send = [torch.empty_like(eng.packed) for _ in range(2)] if use_nccl else None
    recv = torch.empty((world, eng.packed.numel()), dtype=torch.uint8, device=dev) if (use_nccl and rank == 0) else None
    if window is not None and rank == 0:
        recv = window.rows()
    recv_list = list(recv.unbind(0)) if (use_nccl and recv is not None) else None
    comm_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()
    moments_done = torch.cuda.Event()
    fused_done = torch.cuda.Event()
    push_done = torch.cuda.Event()
    gather_done = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"k": 0, "pending": None}

    def launch_gather(buf_index):
        with torch.cuda.stream(comm_stream):
            dist.gather(send[buf_index], recv_list, dst=0)
            gather_done[buf_index].record(comm_stream)

    def one_step():
        if world == 1:
            graph.replay() if graph is not None else eng.run()
            return
        k = state["k"]
        graph_moments.replay() if graph is not None else eng.run_moments()
        if window is not None:
            main_stream.wait_event(push_done)          # the previous step's copy has read eng.packed
            graph_fused.replay() if graph is not None else eng.run_fused()
            fused_done.record(main_stream)
            comm_stream.wait_event(fused_done)
            window.push(eng.packed, comm_stream)       # lands in rank 0's window under the next step's kernels
            push_done.record(comm_stream)
            state["k"] = k + 1
            return
        if state["pending"] is not None:              # records of step k-1: gather them under this step's DP
            moments_done.record(main_stream)
            comm_stream.wait_event(moments_done)
            launch_gather(state["pending"])
        graph_fused.replay() if graph is not None else eng.run_fused()
        main_stream.wait_event(gather_done[k & 1])     # the gather that last read this send buffer (step k-2)
        send[k & 1].copy_(eng.packed)
        state["pending"] = k & 1
        state["k"] = k + 1

    def drain():
        """The records of the last step (nothing is left to overlap them with)."""
        if window is not None:
            main_stream.wait_event(push_done)
        elif world > 1 and state["pending"] is not None:
            moments_done.record(main_stream)
            comm_stream.wait_event(moments_done)
            launch_gather(state["pending"])
            main_stream.wait_event(gather_done[state["pending"]])
            state["pending"] = None

    push_done.record(main_stream)
    for ev in gather_done:
        ev.record(main_stream)
    for _ in range(args.warmup):
        one_step()
    drain()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    if world > 1:
        dist.barrier()   # after rank 0 started its clock sampler, so all ranks enter the timed loop together
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    t_start.record()
    for _ in range(args.steps):
        one_step()
    drain()                               # the last gather belongs to the timed region
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    total_ms = t_start.elapsed_time(t_end)
    # the dominant kernel alone (CUDA events on its launch stream; inputs >> L2)
    k1_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in k1_events:
        eng.time_moments(a, b)
    torch.cuda.synchronize()
    k1_ms = [a.elapsed_time(b) for a, b in k1_events]
    # the same two kernels inside whole steps (eager launches of the same sequence, events around them)
    step_events = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for evs in step_events:
        eng.time_step(*evs)
    torch.cuda.synchronize()
    k1_in_step = float(np.mean([e[0].elapsed_time(e[1]) for e in step_events]))
    fu_in_step = float(np.mean([e[1].elapsed_time(e[2]) for e in step_events]))
    # the duration of those eager steps themselves (graph replay of the same launches is a few per cent faster)
    eager_step = step_events[0][0].elapsed_time(step_events[-1][0]) / max(1, args.steps - 1) if args.steps > 1 else float("nan")
    # the fused fix-up + DP launch alone
    fu_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in fu_events:
        eng.time_fused(a, b)
    torch.cuda.synchronize()
    fu_ms = float(np.mean([a.elapsed_time(b) for a, b in fu_events]))
    clocks = sampler.finish() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = G / (ms_per_step * 1e-3)

    # sharded == single-GPU: rank 0 recomputes the LAST rank's block on its own GPU and compares the records
    shard_check = None
    if world > 1 and rank == 0:
        lo2, hi2 = distributed.shard_range(G, world - 1, world)
        w2 = DeviceWorkload(args.workload, args.dropout, lo2, hi2, dev, moments=args.moments, capacity=cap)
        w2.engine.run()
        torch.cuda.synchronize()
        shard_check = {"block": "rank %d's genes [%d, %d) recomputed on rank 0" % (world - 1, lo2, hi2),
                       "records_identical": bool(torch.equal(w2.engine.packed, recv[world - 1][:w2.engine.packed.numel()]))}
        del w2
        torch.cuda.empty_cache()

    # ---- end-to-end arm: public API, pinned host blocks -> host result objects on rank 0 ----------------
    e2e_ms, ctor_ms = [], []
    d2h, digest = 0, None
    adata_ref = AnnDataLite(Xr_pin[:, :hi - lo] if Xr_pin.shape[1] != hi - lo else Xr_pin, w.tr, w.genes)
    adata_query = AnnDataLite(Xq_pin[:, :hi - lo] if Xq_pin.shape[1] != hi - lo else Xq_pin, w.tq, w.genes)
    for k in range(args.e2e_steps + 1 if args.e2e_steps > 0 else 0):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        aligner = Main.RefQueryAligner(adata_ref, adata_query, all_genes, T, verbose=False, distributed=world > 1)
        t_ctor = time.perf_counter() - t0     # returns with the uploads still in flight
        aligner.align_all_pairs()
        n_match = sum(a.alignment_str.count("M") for a in aligner.results)  # touch every result
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if k > 0:  # first call is warm-up
            e2e_ms.append(dt * 1e3)
            ctor_ms.append(t_ctor * 1e3)
        d2h = int(sum(v.nbytes for v in aligner._host.values() if hasattr(v, "nbytes")))
        if rank == 0:
            digest = result_digest(aligner._host)
            n_results = len(aligner.results)
        del aligner
    e2e_t = float(np.mean(e2e_ms)) * 1e-3 if e2e_ms else float("nan")
    if world > 1:
        t = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t.item())
    h2d = int(4 * (nr + nq) * G + 8 * (nr + nq) * world)   # all ranks: column blocks + pseudotimes

    if rank == 0:
        peak, peak_src = measured_peak()
        n_loc = hi - lo
        alg_bytes = 4.0 * (nr + nq) * n_loc + 8.0 * (nr + nq)
        k1 = k1_in_step * 1e-3                 # the kernel's duration inside the step
        achieved = alg_bytes / k1 / 1e9
        mma = eng.variant == "mma"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": ("f16 hi/lo split operands, f32 tensor-core chains of 128 cells, f32/f64 accumulation (interpolation); "
                      "f64 (cost + DP)") if mma else "f32 products / f64 accumulation (interpolation), f64 (cost + DP)",
            "data": "synthetic",
            "config": config_dict(args.workload, args.dropout, world),     # identical in the reference arm
            "run": {"cuda_graph": graph is not None, "wall_s_timed_region": wall,
                    "gather": "none" if world == 1 else (
                        "peer copy: each rank writes its records into rank 0's peer-mapped window with one device-to-"
                        "device copy over NVLink (copy engine, side stream), under the next step's kernels"
                        if window is not None else
                        "NCCL gather of step k-1 on a side stream next to step k's fused fix-up / DP launch "
                        "(two send buffers)")},
            "e2e": {"value": G / e2e_t, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_t * 1e3, "steps": len(e2e_ms),
                    "ms_each": [round(v, 3) for v in e2e_ms],
                    "constructor_ms": round(float(np.mean(ctor_ms)), 3) if ctor_ms else None,
                    "results_on_rank0": n_results if e2e_ms else None,
                    "what": "Main.RefQueryAligner(..., distributed=%s) + align_all_pairs() from pinned host column "
                            "blocks (original cell order) to host result objects on rank 0" % (world > 1)},
            "gpu_launches": eng.launches_per_run * args.steps,
            "roofline": {"kernel": "moments_mma_kernel (K1m g2g_moments_mma)" if mma else "moments_kernel (K1 g2g_moments)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("%s_n%d" % (args.workload, world)),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": k1 * 1e3, "kernel_share_of_step": k1 * 1e3 / eager_step,
                         "kernel_ms_launched_alone": float(np.mean(k1_ms)), "eager_step_ms": eager_step,
                         "timing": "CUDA events around the launch inside %d whole steps launched eagerly (events cannot be "
                                   "timed inside the replayed graph; eager_step_ms is the duration of those steps, the "
                                   "share is taken of it); launched_alone = the kernel back to back without the rest"
                                   % args.steps,
                         "genes_in_launch": n_loc,
                         "tensor": {"flops_per_launch": 2.0 * 3 * 2 * 64 * (nr + nq) * n_loc,
                                    "note": "six f16 MMAs (M 128, N 64, K 16) per 16 cells x 128 genes; the tensor "
                                            "pipe is not the bound"} if mma else None},
            "fused_fixup_dp": dict({"kernel": "fixup_dp_kernel (K3 + K4 g2g_fixup_cost_dp)", "kernel_ms": fu_in_step,
                                    "kernel_share_of_step": fu_in_step / eager_step, "kernel_ms_launched_alone": fu_ms, "bound": "latency (one warp per gene, "
                                    "2T+1 dependent wavefront steps)"}, **fused_ncu(args.workload)),
            "result_sha256_24": digest,
            "shard_check": shard_check,
            "clocks": clocks,
        }
        if world == 1 and not args.no_secondary:
            del eng, w
            torch.cuda.empty_cache()
            line["secondary"] = [time_small_workload(n, args.dropout, dev, 20, 3) for n in ("C2", "C3")]
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference_measure(args.workload, args.dropout, args.cpu_genes_per_core, dev)
        emit(line)
    if world > 1:
        del graph
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line):
    """The contract's ONE JSON line goes to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    args = parse_args()
    # anything a library prints on stdout (NCCL's version banner, warnings) must not mix with the result
    # line: file descriptor 1 is pointed at stderr for the run and the line is written to the saved one
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
