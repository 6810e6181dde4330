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


def config_of(name, scale, par, nsamples, world):
    """The `config` object of the JSON line: the same keys and values in both arms."""
    return {"workload": WORKLOADS[name], "name": name, "scale": scale,
            "nodes": int(par['nx'] * par['ny'] * par['nz']), "samples": int(nsamples),
            "l2": "flushed between timed steps (256 MiB memset, untimed); outputs (16 B/node) "
                  "also exceed L2",
            "parallelism": "z-slabs x%d, samples replicated" % world}


def strided_nodes(total, count, offset=0):
    """`count` node indices spread evenly over the whole grid."""
    count = int(max(1, min(total, count)))
    return (offset + (np.arange(count, dtype=np.int64) * total) // count) % total


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
    workload: nodes spread evenly over the whole grid, all host cores (OpenMP)."""
    from oracle import kt3d_oracle as orc
    cores = os.cpu_count() or 1
    cols = dict(data)
    st = orc.prepare(par, cols, [n for n in data if n not in ('x', 'y', 'z')])
    total = par['nx'] * par['ny'] * par['nz']
    probe = strided_nodes(total, 4096)
    t0 = time.time()
    orc.run(st, extgrid=sec, node_list=probe, threads=cores, neighbours=False)
    rate = len(probe) / max(time.time() - t0, 1e-6)
    nodes = strided_nodes(total, rate * seconds_target, 3)
    t0 = time.time()
    orc.run(st, extgrid=sec, node_list=nodes, threads=cores, neighbours=False)
    dt = time.time() - t0
    return st, {"value": len(nodes) / dt, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d nodes spread evenly over the whole grid, oracle/kt3d_oracle.c "
                          "with OpenMP over nodes, %.1f s" % (len(nodes), dt)}


def parity_check(par, data, sec, est, var, count=2000):
    """Outside the timed region: the arrays Krige3d.kt3d() returned (for N > 1: through the
    multi-rank branch) against the oracle on `count` nodes spread over the whole grid."""
    from oracle import kt3d_oracle as orc
    total = par['nx'] * par['ny'] * par['nz']
    nodes = np.unique(np.concatenate([strided_nodes(total, count, 11), [0, total - 1]]))
    st = orc.prepare(par, dict(data), [n for n in data if n not in ('x', 'y', 'z')])
    oe, ov, _, _ = orc.run(st, extgrid=sec, node_list=nodes, threads=os.cpu_count() or 1,
                           neighbours=False)
    ge, gv = est[nodes], var[nodes]
    nan_equal = bool(np.array_equal(np.isnan(ge), np.isnan(oe)) and
                     np.array_equal(np.isnan(gv), np.isnan(ov)))
    ok = ~np.isnan(oe) & ~np.isnan(ge)
    vmax = float(np.max(np.abs(data['v'])))
    rel = np.abs(ge[ok] - oe[ok]) / np.maximum(np.abs(oe[ok]), 1e-3 * vmax) if ok.any() else np.zeros(1)
    dv = np.abs(gv[ok] - ov[ok]) if ok.any() else np.zeros(1)
    res = {"nodes": int(len(nodes)), "estimated": int(ok.sum()), "max_rel_est": float(rel.max()),
           "max_abs_var": float(dv.max()), "nan_equal": nan_equal,
           "tolerance": "est 1e-9 relative (floor 1e-3 max|v|), var 1e-9 absolute"}
    res["pass"] = bool(nan_equal and res["max_rel_est"] <= 1e-9 and res["max_abs_var"] <= 1e-9)
    return res


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is
    pure Python and cannot travel to the GPU box, so this times the oracle port of it
    (same algorithm, compiled C, all host threads) on a bounded sample per step: nodes
    spread evenly over the WHOLE grid of the same configuration."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import kt3d_oracle as orc
    par, data, sec = k3d_configs.config(args.config, args.scale)
    cores = os.cpu_count() or 1
    st = orc.prepare(par, dict(data), [n for n in data if n not in ('x', 'y', 'z')])
    total = par['nx'] * par['ny'] * par['nz']
    probe = strided_nodes(total, 4096)
    t0 = time.time()
    orc.run(st, extgrid=sec, node_list=probe, threads=cores, neighbours=False)
    rate = len(probe) / max(time.time() - t0, 1e-6)
    per_step = 60.0 / max(1, args.steps + args.warmup)          # whole run ~1 minute
    nn = int(max(8192, min(total, rate * per_step)))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    for s_ in range(args.warmup):
        orc.run(st, extgrid=sec, node_list=strided_nodes(total, nn, s_), threads=cores, neighbours=False)
    t0 = time.time()
    for s_ in range(args.steps):
        orc.run(st, extgrid=sec, node_list=strided_nodes(total, nn, 7 + s_), threads=cores,
                neighbours=False)
    dt = time.time() - t0
    value = nn * args.steps / dt
    sample = "%d nodes per step spread evenly over the whole grid (every %.1f-th node)" % (nn, total / nn)
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args.config, args.scale, par, len(data['x']), world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure Python+Numba (808 nodes/s on config 1 in the build container, "
                "SURVEY.md section 6); this arm times its compiled C restatement (oracle/) "
                "with OpenMP on all host cores"})


def device_timed(name, scale, steps, warmup, world, rank, dev, with_stats=True):
    """K timed passes of the two kernels over this rank's z-slab of configuration `name`
    (inputs resident in HBM; for N > 1 each step ends with the all-gather of the slabs).
    Returns the numbers of the JSON line that come from the device."""
    import torch
    import torch.distributed as dist
    from pygeostatistics_b200 import Krige3d, _native
    from pygeostatistics_b200.krige3d import slab_ranges, gather_slabs
    par, data, sec = k3d_configs.config(name, scale)
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
            gather_slabs(out['est'], ranges, plane)
            gather_slabs(out['var'], ranges, plane)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(warmup, 3)):
        step()
    nbar = cbar = None
    unest = 0
    if with_stats:  # statistics for the roofline (n-bar, C-bar), outside the timed region
        sout = k._alloc_outputs(nodes, neighbours=True)
        k._launch(P, d, iz0, iz1, sout)
        torch.cuda.synchronize(dev)
        nbar = float(sout['nbr_cnt'].double().mean())
        cbar = float(sout['ncand'].double().mean())
        unest = int((sout['status'] != 0).sum())
        del sout
    local = dev.index
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches0 = _native.last_launch()['launches']
    t_wall0 = time.time()
    for s_ in range(steps):
        flush.zero_()                       # L2 flush between timed iterations (untimed)
        ev[s_][0].record(stream)
        k._launch(P, d, iz0, iz1, out)
        ev[s_][1].record(stream)
        if world > 1:
            gather_slabs(out['est'], ranges, plane)
            gather_slabs(out['var'], ranges, plane)
        ev[s_][2].record(stream)
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
    del flush
    return dict(par=par, data=data, sec=sec, k=k, host=host, ext_slab=ext_slab, nodes=nodes,
                total_nodes=total_nodes, step_ms=step_ms, kern_ms=kern_ms, nbar=nbar, cbar=cbar,
                unest=unest, linfo=linfo, launches=launches, clocks=sampler.summary(),
                t_wall=t_wall, value=total_nodes * steps / (step_ms * 1e-3))


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
    ap.add_argument("--gather", default="host", choices=["host", "device"],
                    help="how Krige3d.kt3d() assembles the grid of a multi-rank run (e2e)")
    ap.add_argument("--also-c5", action="store_true",
                    help="add a device-timed line for configuration 5 under the key \"c5\" "
                         "(default at 8 GPUs, the configuration BASELINE.json states for them)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from pygeostatistics_b200 import Krige3d, _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    m = device_timed(args.config, args.scale, args.steps, args.warmup, world, rank, dev)
    par, data, sec = m['par'], m['data'], m['sec']
    total_nodes, nodes = m['total_nodes'], m['nodes']
    step_ms, kern_ms, linfo = m['step_ms'], m['kern_ms'], m['linfo']
    nbar, cbar = m['nbar'], m['cbar']
    stream = torch.cuda.current_stream(dev)

    # ---- end to end through the public API: Krige3d.kt3d() with host arrays ------------
    e2e, ke = None, None
    if not args.no_e2e:
        h2d = sum(v.numel() * v.element_size() for v in m['host'].values() if v is not None)
        if m['ext_slab'] is not None:
            h2d += m['ext_slab'].numel() * 8
        t1 = torch.tensor([float(h2d)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t1)             # every rank uploads the sample set and its index
        ke = Krige3d(par, data=data, secondary=sec)
        ke.verbose = False
        ke.kt3d(gather=args.gather)         # warm
        sync_all()
        t0 = time.time()
        reps = max(1, min(args.steps, 3))
        for _ in range(reps):
            ke = Krige3d(par, data=data, secondary=sec)
            ke.verbose = False
            ke.kt3d(gather=args.gather)
        sync_all()
        dt = (time.time() - t0) / reps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        # est + var + status; gather="device": every rank copies the whole grid
        d2h = total_nodes * 17 * (world if (world > 1 and args.gather == "device") else 1)
        e2e = {"value": total_nodes / float(tt[0]), "unit": UNIT,
               "h2d_bytes_per_step": int(t1[0]),
               "breakdown_s": {k_: round(v, 4) for k_, v in ke.timings.items()},
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(tt[0]) * 1e3,
               "includes": "Krige3d(par, data).kt3d(): data checksum + cached super-block set-up "
                           "(binning, sort, template: computed by the warm-up call), pinned H2D of "
                           "samples/index on every rank, search + solve kernels, D2H of est/var/"
                           "status overlapped with the kernels over runs of z-planes" +
                           ("; each rank copies its own slab into one host buffer the ranks share "
                            "(whole-job bytes)" if (world > 1 and args.gather == "host") else
                            "; all-gather of the slabs, every rank copies the grid" if world > 1 else "")}
    elif rank == 0 or world > 1:
        ke = Krige3d(par, data=data, secondary=sec)
        ke.verbose = False
        ke.kt3d(gather=args.gather)
        sync_all()

    extra_c5 = None
    if (args.also_c5 or world == 8) and args.config != "c5" and not os.environ.get("K3D_BENCH_NO_C5"):
        c5 = device_timed("c5", 1.0, max(2, min(args.steps, 3)), 3, world, rank, dev, with_stats=False)
        extra_c5 = {"metric": METRIC, "value": c5['value'], "unit": UNIT, "n_gpus": world,
                    "ms_per_step": c5['step_ms'] / max(2, min(args.steps, 3)),
                    "kernel_ms_per_step": c5['kern_ms'] / max(2, min(args.steps, 3)),
                    "config": config_of("c5", 1.0, c5['par'], len(c5['data']['x']), world),
                    "clocks": c5['clocks'], "gpu_launches": int(c5['launches'])}
        del c5

    fail = False
    if rank == 0:
        tfl = C_double()
        peak_ms = C_double()
        _native.check(_native.lib().k3d_measure_fp64_peak(byref(tfl), byref(peak_ms),
                                                          stream.cuda_stream))
        flops_node = algorithmic_flops(par, nbar, cbar, m['k'].mdt)
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
            "metric": METRIC, "value": m['value'], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_of(args.config, args.scale, par, len(data['x']), world),
            "stats": {"mean_neighbours": nbar, "mean_candidates_tested": cbar,
                      "unestimated_nodes_rank0": m['unest']},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": tfl.value,
                         "unit": "TFLOP/s", "frac": achieved / tfl.value, "traffic": traffic,
                         "peak_source": "DFMA microbenchmark in this run (k3d_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no fp64 entry; nominal 37.2",
                         "frac_of_nominal_37.2": achieved / 37.2,
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
            "gpu_launches": int(m['launches']), "clocks": m['clocks'],
            "launch": linfo,
            "wall_s_timed_region": m['t_wall'],
        }
        if e2e is not None:
            line["e2e"] = e2e
        if ke is not None:
            line["parity_check"] = parity_check(par, data, sec, ke.estimation, ke.estimation_variance)
            line["parity_check"]["through"] = "Krige3d.kt3d(gather=%r), %d rank(s)" % (args.gather, world)
            fail = not line["parity_check"]["pass"]
        if extra_c5 is not None:
            line["c5"] = extra_c5
        if not args.no_cpu_baseline and world == 1:
            _, line["cpu_baseline"] = cpu_baseline(par, data, sec)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if fail:
        sys.exit(3)


if __name__ == "__main__":
    main()
