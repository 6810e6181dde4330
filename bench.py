#!/usr/bin/env python
"""bench.py -- grid nodes kriged per second (fp64) on N B200s, for BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl ours|reference]

A "step" is one pass of the krige3d hot path (search kernel + assembly/solve kernel)
over the whole synthetic grid of the chosen configuration (default c3 = BASELINE.json
configs[2], the 256^3 / ndmax=32 ordinary-kriging case the metric is quoted on).  With N>1
(torchrun, one rank per GPU) the grid is split into z-slabs, samples replicated, and each
step ends with the all-gather of the slabs (total work fixed: strong scaling).

Prints ONE JSON line on rank 0 (see README/DESIGN.md for every key).
"""
import argparse
import json
from ctypes import c_double as C_double, byref
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import k3d_configs  # noqa: E402

METRIC = "grid nodes kriged/sec (fp64, ndmax=32)"
UNIT = "nodes/s"
WORKLOADS = {
    "c2": "synthetic 10k samples, 128^3 grid, simple kriging, spherical+nugget, ndmax=16, isotropic",
    "c3": "synthetic 100k samples, 256^3 grid, ordinary kriging, nested anisotropic exp+sph, "
          "ndmax=32, octant search (noct=4)",
    "c4": "synthetic 100k samples, 256^3 grid, universal kriging with 9 drift terms, ndmax=32",
    "c5": "synthetic 1M samples, 512x512x256 grid, block kriging 4x4x2 with external drift, ndmax=32",
}


_RESULT_OUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's version
    banner under torchrun goes through printf), so the process's fd 1 is pointed at stderr
    and the result line is written to a private duplicate of the original stdout."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def algorithmic_flops(par, nbar, cbar, mdt):
    """flops per node with n = mean neighbours used, C = mean
    candidates distance-tested (add/mul = 1, FMA = 2, div/sqrt/exp/pow/cos = 1 each)."""
    S = par['nst']
    D = max(1, par['nxdis']) * max(1, par['nydis']) * max(1, par['nzdis'])
    n, C, m = nbar, cbar, mdt
    N = n + m
    f = 24.0 * C + 33.0 * S * n * (n + 1) / 2.0
    f += (33.0 * S + (8 if D > 1 else 0)) * n * D
    f += 2.0 * n * max(m - 1, 0)
    f += N ** 3 / 3.0 + 2.0 * N ** 2 + 2.0 * N + 2.0 * n
    return f


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, nm in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"]):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(nm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_baseline(par, data, sec, seconds_target=12.0):
    """The oracle (reference-derived CPU port, oracle/) on a bounded sample of the same
    workload: whole central z-planes, all host cores (OpenMP)."""
    from oracle import kt3d_oracle as orc
    cores = os.cpu_count() or 1
    cols = dict(data)
    st = orc.prepare(par, cols, [n for n in data if n not in ('x', 'y', 'z')])
    plane = par['nx'] * par['ny']
    zc = par['nz'] // 2
    # calibrate on a slice of the central plane, then size the sample for ~seconds_target
    probe = np.arange(zc * plane, zc * plane + min(plane, 2048), dtype=np.int64)
    t0 = time.time()
    orc.run(st, extgrid=sec, node_list=probe, threads=cores, neighbours=False)
    rate = len(probe) / max(time.time() - t0, 1e-6)
    nplanes = int(max(1, min(par['nz'], round(rate * seconds_target / plane))))
    z0 = max(0, zc - nplanes // 2)
    nodes = np.arange(z0 * plane, (z0 + nplanes) * plane, dtype=np.int64)
    t0 = time.time()
    orc.run(st, extgrid=sec, node_list=nodes, threads=cores, neighbours=False)
    dt = time.time() - t0
    return st, {"value": len(nodes) / dt, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d central z-planes (%d nodes) of the same grid, oracle/kt3d_oracle.c "
                          "with OpenMP over nodes, %.1f s" % (nplanes, len(nodes), dt)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is
    pure Python and cannot travel to the GPU box, so this times the oracle port of it
    (same algorithm, compiled C, all host threads) on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import kt3d_oracle as orc
    par, data, sec = k3d_configs.config(args.config, args.scale)
    cores = os.cpu_count() or 1
    st = orc.prepare(par, dict(data), [n for n in data if n not in ('x', 'y', 'z')])
    plane = par['nx'] * par['ny']
    zc = par['nz'] // 2
    probe = np.arange(zc * plane, zc * plane + min(plane, 2048), dtype=np.int64)
    t0 = time.time()
    orc.run(st, extgrid=sec, node_list=probe, threads=cores, neighbours=False)
    rate = len(probe) / max(time.time() - t0, 1e-6)
    per_step = 60.0 / max(1, args.steps + args.warmup)          # whole run ~1 minute
    nn = int(max(plane // 8, min(plane * par['nz'], rate * per_step)))
    nodes = np.arange(zc * plane, zc * plane + nn, dtype=np.int64) % (plane * par['nz'])
    for _ in range(args.warmup):
        orc.run(st, extgrid=sec, node_list=nodes, threads=cores, neighbours=False)
    t0 = time.time()
    for _ in range(args.steps):
        orc.run(st, extgrid=sec, node_list=nodes, threads=cores, neighbours=False)
    dt = time.time() - t0
    value = len(nodes) * args.steps / dt
    sample = "%d consecutive nodes from the central z-plane per step" % len(nodes)
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "name": args.config, "scale": args.scale},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure Python+Numba (808 nodes/s on config 1 in the build container, "
                "SURVEY.md section 6); this arm times its compiled C restatement (oracle/) "
                "with OpenMP on all host cores"})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from pygeostatistics_b200 import Krige3d, _native
    from pygeostatistics_b200.krige3d import slab_ranges, gather_slabs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    par, data, sec = k3d_configs.config(args.config, args.scale)
    k = Krige3d(par, data=data, secondary=sec)
    k.verbose = False
    k.prepare()
    P = k._params_struct()
    plane = par['nx'] * par['ny']
    total_nodes = plane * par['nz']
    ranges = slab_ranges(par['nz'], world)
    iz0, iz1 = ranges[rank]
    nodes = (iz1 - iz0) * plane
    host = k._host_inputs()
    ext = k._external_grid()
    ext_slab = None if ext is None else ext[iz0 * plane: iz1 * plane].pin_memory()
    d = k._upload(host, ext_slab)
    out = k._alloc_outputs(nodes, neighbours=False)
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        k._launch(P, d, iz0, iz1, out)
        if world > 1:
            return (gather_slabs(out['est'], ranges, plane), gather_slabs(out['var'], ranges, plane))
        return out['est'], out['var']

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    # statistics for the roofline (n-bar, C-bar), outside the timed region
    sout = k._alloc_outputs(nodes, neighbours=True)
    k._launch(P, d, iz0, iz1, sout)
    torch.cuda.synchronize(dev)
    st_cnt = sout['nbr_cnt'].double()
    nbar = float(st_cnt.mean()); cbar = float(sout['ncand'].double().mean())
    unest = int((sout['status'] != 0).sum())
    del sout

    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = _native.last_launch()['launches']
    t_wall0 = time.time()
    for s in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (untimed)
        ev[s][0].record(stream)
        k._launch(P, d, iz0, iz1, out)
        ev[s][1].record(stream)
        if world > 1:
            gather_slabs(out['est'], ranges, plane); gather_slabs(out['var'], ranges, plane)
        ev[s][2].record(stream)
    sync_all()
    t_wall = time.time() - t_wall0
    sampler.stop_flag = True
    linfo = _native.last_launch()       # also: device time of the last search / solve kernels
    launches = linfo['launches'] - launches0
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in ev)
    step_ms = sum(a.elapsed_time(c) for a, _, c in ev)
    t = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms = float(t[0]), float(t[1])
    value = total_nodes * args.steps / (step_ms * 1e-3)

    # ---- end to end through the public API: Krige3d.kt3d() with host arrays ------------
    e2e = None
    if not args.no_e2e:
        h2d = sum(v.numel() * v.element_size() for v in host.values() if v is not None)
        if ext_slab is not None:
            h2d += ext_slab.numel() * 8
        d2h = nodes * (8 + 8 + 1)
        ke = Krige3d(par, data=data, secondary=sec)
        ke.verbose = False
        ke.kt3d()                           # warm
        sync_all()
        t0 = time.time()
        reps = max(1, min(args.steps, 3))
        for _ in range(reps):
            ke = Krige3d(par, data=data, secondary=sec)
            ke.verbose = False
            ke.kt3d()
        sync_all()
        dt = (time.time() - t0) / reps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": total_nodes / float(tt[0]), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "breakdown_s": {k: round(v, 4) for k, v in ke.timings.items()},
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(tt[0]) * 1e3,
               "includes": "Krige3d(par, data).kt3d(): super-block set-up (binning, template "
                           "kernel), pinned H2D of samples/index, search + solve kernels, D2H of "
                           "est/var/status" + ("; all-gather of slabs" if world > 1 else
                                               " (overlapped with the kernels over runs of z-planes)")}

    if rank == 0:
        tfl = C_double()
        peak_ms = C_double()
        _native.check(_native.lib().k3d_measure_fp64_peak(byref(tfl), byref(peak_ms),
                                                          stream.cuda_stream))
        flops_node = algorithmic_flops(par, nbar, cbar, k.mdt)
        # dominant kernel = the solve kernel: everything but the 24*C search term
        flops_solve = flops_node - 24.0 * cbar
        solve_ms, search_ms = linfo['solve_ms'], linfo['search_ms']
        share = solve_ms / max(solve_ms + search_ms, 1e-9)
        solve_ms_avg = kern_ms / args.steps * share     # events of the timed region x share
        achieved = flops_solve * nodes / (solve_ms_avg * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.config)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config], "name": args.config,
                       "scale": args.scale, "nodes": total_nodes, "samples": int(len(data['x'])),
                       "mean_neighbours": nbar, "mean_candidates_tested": cbar,
                       "unestimated_nodes_rank0": unest, "l2": "flushed between timed steps "
                       "(256 MiB memset, untimed); outputs (16 B/node) also exceed L2",
                       "parallelism": "z-slabs x%d, samples replicated" % world},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": tfl.value,
                         "unit": "TFLOP/s", "frac": achieved / tfl.value, "traffic": traffic,
                         "peak_source": "DFMA microbenchmark in this run (k3d_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no fp64 entry; nominal 37.2",
                         "kernel": "k3d_solve_kernel (assembly + LDLt solve)",
                         "flops_per_node": flops_solve,
                         "kernel_ms": solve_ms_avg,
                         "search_kernel": {"ms": kern_ms / args.steps * (1.0 - share),
                                           "flops_per_node": 24.0 * cbar,
                                           "achieved": 24.0 * cbar * nodes /
                                           (kern_ms / args.steps * (1.0 - share) * 1e-3) / 1e12},
                         "path_flops_per_node": flops_node,
                         "path_achieved": flops_node * nodes * args.steps / (kern_ms * 1e-3) / 1e12,
                         "convention": "SURVEY.md 8(d) algorithmic flops at measured n-bar, C-bar",
                         "kernel_ms_per_step": kern_ms / args.steps,  # search + solve
                         "hbm_gbs_algorithmic": nodes * 16.0 * args.steps / (kern_ms * 1e-3) / 1e9},
            "gpu_launches": int(launches), "clocks": sampler.summary(),
            "launch": linfo,
            "wall_s_timed_region": t_wall,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            _, line["cpu_baseline"] = cpu_baseline(par, data, sec)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
