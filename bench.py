#!/usr/bin/env python
"""bench.py -- node-states/sec per EM iteration (emission + cost arrays + posteriors/costs +
statistics), the metric of BASELINE.json.  GCO and the M-step are excluded by the metric's
definition; phase B runs on arg-min labels (the GCO stand-in of 8(d)).

    python bench.py --gpus N --steps K --warmup W           # this repo, one process per GPU
    python bench.py --impl reference --steps K --warmup W   # the reference's CPU path

A *step* is one pass of the hot path over one row band of the synthetic contact map:
phase A1 emission (+max|logp|), A2 integer unary, B fused posteriors/costs/statistics, and
for N>1 one NCCL all-reduce of the K(1+d+d^2)+3 statistics.  Weak scaling: every rank owns
one band of `nodes_per_gpu` nodes of the chr1@10kb triangle (B=24895 bins, d=9, K=30); at
N=8 the bands tile the whole map (config 5 of BASELINE.json).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B bins, d leaves, K states, number of bands the map is cut into)
    "cfg5_chr1_10kb_band": (24895, 9, 30, 8),
    "cfg3_chr1_50kb": (4979, 5, 20, 1),
    "cfg4_genome_50kb_band": (4979, 5, 20, 1),
    "tiny": (400, 9, 30, 1),
    "mid_d9_k30": (5000, 9, 30, 1),
}
METRIC = "node-states/sec per EM iteration (emission+costs+stats)"
UNIT = "node-states/s"
BETA, BETA1, ESTIMATE_TYPE = 1.0, 0.1, 3


def algorithmic_bytes_per_node(d, K):
    return 16 * d + 4 * K + 68


def algorithmic_flops_per_node_state(d):
    return 2 * d * d + 6 * d + 12


def estep_flops_per_node_state(d):
    return 6 + 2 + 2 * d + d * (d + 1)  # posterior arithmetic + packed statistics share of F(d)


def emit_flops_per_node_state(d):
    return d * d + 3 * d + 2


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed `ncu --set full` capture of this workload (profiles/r1_traffic.json); None if the
    workload was not captured."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        return table[workload][kernel]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import phylo_hmrf_b200 as ph
    from phylo_hmrf_b200 import synth, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    seed = 20261017 + 5
    B, d, K, n_bands = WORKLOADS[args.workload]
    # model: identical on every rank (a pure function of the seed and the first rows of the map)
    g0 = synth.make_band(seed, B, d, 0, min(B, 24), beta1=BETA1)
    means, covars = synth.model(seed, g0["X_own"], K, d)
    del g0
    V = synth.potts(K, BETA)
    stream = torch.cuda.Stream()
    m = ph.Model(K, d, device=local_rank)
    m.set_model(means, covars, V)

    # The band is built on the device from the grid geometry (phmrf_region_create_grid): the host
    # only prepares the window's feature rows, no edge array exists on the host or crosses PCIe.
    r0, r1 = synth.band_rows(B, n_bands)[rank % n_bands]
    g = synth.window_xy(B, r0, r1)
    n, n_window, own_offset = g["n_own"], g["n_window"], g["own_offset"]
    X_window = synth.features(seed, g["x"], g["y"], d)
    g.clear()
    reg = m.region_grid(X_window, 1, B, B, r0, r1, 8, BETA1, stream=stream.cuda_stream)
    E = reg.n_edges
    X_pin = torch.empty((n, d), dtype=torch.float64, pin_memory=True)  # pinned: end-to-end leg
    X_pin.numpy()[:] = X_window[own_offset:own_offset + n]
    # labels for phase B: arg-min of the integer unary over the *window* (owned + halo rows); the
    # halo labels come from a throw-away region over the window (what the neighbouring bands' graph
    # cuts would have produced)
    if n_window != n:
        win = m.region(X_window, np.zeros((0, 2), np.int64), np.zeros(0))
        win.emit_loglik()
        win.quantise(want_unary=False, want_edges=False)
        labels_window = win.labels_argmin_unary()
        win.close()
    else:
        reg.emit_loglik()
        reg.quantise(want_unary=False, want_edges=False)
        labels_window = reg.labels_argmin_unary()
    reg.set_labels(labels_window)
    del X_window

    stats_len = m.stats_len

    from phylo_hmrf_b200 import dist as pdist
    stats_dev = pdist.stats_tensor(reg)      # the library's device statistics buffer (for NCCL)
    absmax_dev = pdist.absmax_tensor(reg)    # max|logp| of the band as ordered int64 bits
    if world > 1:
        pdist.share_weight_max(reg, dist)    # the bands are ONE region: shared down-weight factor

    def step():
        reg.emit_loglik_async()
        if world > 1:                        # region-wide max|logp| before the integer conversion
            with torch.cuda.stream(stream):
                dist.all_reduce(absmax_dev, op=dist.ReduceOp.MAX)
        ev[1].record(stream)
        reg.quantise_async()
        ev[2].record(stream)
        reg.estep_stats_async(ESTIMATE_TYPE)
        ev[3].record(stream)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(stats_dev)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            ev[0].record(stream)
            step()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = engine.launch_count()
        phase_ms = np.zeros(3)
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        per_step_events = []
        t_start.record(stream)
        for _ in range(args.steps):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[:] = evs
            ev[0].record(stream)
            step()
            per_step_events.append(evs)
        t_end.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = engine.launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    for evs in per_step_events:
        for i in range(3):
            phase_ms[i] += evs[i].elapsed_time(evs[i + 1])
    phase_ms /= args.steps
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    nodes_total = torch.tensor([float(n)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(nodes_total)
    nodes_total = float(nodes_total.item())
    value = nodes_total * K / (ms_per_step * 1e-3)

    # ---- end-to-end leg: host buffers in, host results out, every step
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    unary_pin = torch.empty((n, K), dtype=torch.int32, pin_memory=True)
    wi_pin = torch.empty(E, dtype=torch.int32, pin_memory=True)
    lab_pin = torch.empty(n_window, dtype=torch.int32, pin_memory=True)
    lab_pin.numpy()[:] = labels_window
    lib = ph._lib.lib()
    import ctypes as C
    K_i32 = np.empty((K, K), np.int32)
    stats_host = np.empty(stats_len - 3)
    sums_host = np.empty(3)

    def e2e_step():
        reg.update_X(X_pin.numpy())                       # H2D: X
        ph._lib.check(lib.phmrf_emit_loglik(reg._h, None))
        dwf, nb = C.c_double(), C.c_int64()
        ph._lib.check(lib.phmrf_quantise(reg._h, 0.0, 1e-9, ph._lib.i32ptr(unary_pin.numpy()),
                                         ph._lib.i32ptr(wi_pin.numpy()), ph._lib.i32ptr(K_i32), C.byref(dwf), None, 0,
                                         C.byref(nb)))     # D2H: integer unary + edge weights (GCO inputs)
        ph._lib.check(lib.phmrf_set_labels(reg._h, ph._lib.i32ptr(lab_pin.numpy())))   # H2D: labels (GCO output)
        ph._lib.check(lib.phmrf_estep_stats(reg._h, ESTIMATE_TYPE, None, ph._lib.dptr(stats_host),
                                            ph._lib.dptr(sums_host)))                  # D2H: statistics
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(stats_dev)
            stream.synchronize()

    e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d = n * d * 8 + n_window * 4
    d2h = n * K * 4 + E * 4 + stats_len * 8

    if rank != 0:
        reg.close()
        m.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (phase B) and of the whole step
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    fp64_peak = engine.probe(0, device=local_rank)  # TFLOP/s, DFMA chains, measured in this run
    F, Bb = algorithmic_flops_per_node_state(d), algorithmic_bytes_per_node(d, K)
    step_s = phase_ms.sum() * 1e-3
    fp64_roof = fp64_peak * 1e12 / F
    hbm_roof = hbm_peak * 1e9 / Bb * K
    bound = "fp64" if fp64_roof < hbm_roof else "hbm"
    b_s = phase_ms[2] * 1e-3
    b_tflops = n * K * estep_flops_per_node_state(d) / b_s / 1e12
    roofline = {
        "kernel": "estep_mma_kernel (phase B: posteriors+costs+statistics, csrc/kernels_b2.cu)",
        "bound": bound, "achieved": b_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": b_tflops / fp64_peak,
        "traffic": ncu_traffic(args.workload, "estep"),
        "peak_source": "DFMA-chain probe in this run (csrc/probe.cu); HBM %s" % hbm_src,
        "kernel_share_of_step": phase_ms[2] / phase_ms.sum(),
        "phase_ms": {"A1_emit": phase_ms[0], "A2_quantise": phase_ms[1], "B_estep": phase_ms[2]},
        "emit_kernel": {"achieved": n * K * emit_flops_per_node_state(d) / (phase_ms[0] * 1e-3) / 1e12,
                        "unit": "TFLOP/s", "frac": n * K * emit_flops_per_node_state(d) / (phase_ms[0] * 1e-3) / 1e12 / fp64_peak},
        "quantise_kernel": {"achieved": n * K * 12 / (phase_ms[1] * 1e-3) / 1e9, "unit": "GB/s",
                            "frac": n * K * 12 / (phase_ms[1] * 1e-3) / 1e9 / hbm_peak},
        "step": {"node_states_per_s_per_gpu": n * K / step_s, "fp64_roof": fp64_roof, "hbm_roof": hbm_roof,
                 "frac_of_slower_roof": n * K / step_s / min(fp64_roof, hbm_roof),
                 "algorithmic_flops_per_node_state": F, "algorithmic_bytes_per_node": Bb},
    }
    cpu = cpu_baseline(d, K, seconds=args.cpu_seconds) if world == 1 and not args.no_cpu else None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: B=%d bins, d=%d leaves, K=%d states, rows [%d,%d) of rank 0; %d nodes per GPU, "
                               "%d nodes total" % (args.workload, B, d, K, r0, r1, n, int(nodes_total)),
                   "nodes_per_gpu": n, "edges_per_gpu": E, "estimate_type": ESTIMATE_TYPE, "beta": BETA, "beta1": BETA1,
                   "labels": "arg-min of the integer unary (GCO stand-in, SURVEY 8(d)); GCO and M-step excluded",
                   "l2": "inputs per step (X + log-likelihood + graph, %.1f GB) exceed the 126 MB L2; no flush needed"
                         % ((n * d * 8 + n * K * 8 + n * 96) / 1e9),
                   "parallelism": "row bands, one per GPU; all-reduce of %d doubles" % stats_len},
        "e2e": {"value": nodes_total * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(out))
    reg.close()
    m.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------
# CPU legs (the only place outside tests/ and smoke() that executes oracle/)
# ---------------------------------------------------------------------------------------
def _cpu_region(args):
    seed, B, d, K = args
    from phylo_hmrf_b200 import synth
    from oracle import phmrf_oracle as orc
    try:  # one BLAS/OpenMP thread per region process: the processes already fill the cores
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    g = synth.make_band(seed, B, d, beta1=BETA1)
    g0 = synth.make_band(20261017 + 5, 24895 if d == 9 else 4979, d, 0, 24, beta1=BETA1)
    means, covars = synth.model(20261017 + 5, g0["X_own"], K, d)
    V = synth.potts(K, BETA)
    el = np.column_stack([g["edge_ids"].astype(np.float64), g["edge_dist"]])
    t0 = time.perf_counter()
    orc.estep_region(g["X_own"], means, covars, V, el, BETA1, ESTIMATE_TYPE, faithful=True)
    return g["n_own"], time.perf_counter() - t0


def cpu_sample(d, K, B_crop, n_regions, pool=None):
    """One 'iteration' of the reference's CPU path: one process per region (base.py:357-362),
    each running the loop-faithful restatement of _predict_posteriors minus the graph cut."""
    import multiprocessing as mp
    jobs = [(900 + r, B_crop, d, K) for r in range(n_regions)]
    t0 = time.perf_counter()
    if n_regions == 1:
        res = [_cpu_region(jobs[0])]
    else:
        with mp.get_context("fork").Pool(n_regions) as p:
            res = p.map(_cpu_region, jobs)
    wall = time.perf_counter() - t0
    nodes = sum(r[0] for r in res)
    return nodes, wall, max(r[1] for r in res)


def cpu_baseline(d, K, seconds=15.0):
    cores = min(os.cpu_count() or 1, 32)
    B_crop = 60
    nodes, wall, inner = cpu_sample(d, K, B_crop, cores)
    reps = max(1, int(seconds / max(wall, 1e-3)) - 1)
    best = inner
    for _ in range(min(reps, 3)):
        nodes, wall, inner = cpu_sample(d, K, B_crop, cores)
        best = min(best, inner)
    return {"value": nodes * K / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d regions of a %d-bin triangle (%d nodes total), d=%d K=%d, one forked process per region "
                      "like base.py:357-362; loop-faithful Python-3 restatement of phylo_hmrf.py:297-468 "
                      "(the py2 reference cannot run here); GCO excluded" % (cores, B_crop, nodes, d, K)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, d, K, _ = WORKLOADS[args.workload]
    cores = min(os.cpu_count() or 1, 32)
    B_crop = 60
    for _ in range(min(args.warmup, 1)):
        cpu_sample(d, K, B_crop, cores)
    tot_nodes, tot_t = 0, 0.0
    for _ in range(args.steps):
        nodes, wall, inner = cpu_sample(d, K, B_crop, cores)
        tot_nodes += nodes
        tot_t += inner
    value = tot_nodes * K / tot_t
    sample = ("each step = %d regions of a %d-bin triangle (%d nodes), d=%d K=%d, one forked process per region "
              "(the reference's own parallel width, base.py:357-362); loop-faithful Python-3 port of "
              "phylo_hmrf.py:297-468 because the py2 reference cannot be imported; GCO excluded like the GPU arm"
              % (cores, B_crop, tot_nodes // args.steps, d, K))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "%s (bounded sample: %s)" % (args.workload, sample)},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5_chr1_10kb_band", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
