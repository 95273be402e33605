#!/usr/bin/env python
"""Benchmark of the harmonic power-flow solve path (driver contract, see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm (oracle port)
    torchrun ... bench.py --gpus N ...                             # N > 1, one rank per GPU

Workload (BASELINE.json configs[2], the configuration the metric and target are quoted on):
net3, coupled SMPS Norton equivalent, fundamental + odd harmonics to the 25th (N = 101),
65,536 seeded randomised load / spectrum scenarios PER GPU (weak scaling: scenarios are
independent, each rank solves its own block, no data-path collective).  One "step" = one
complete solve of the batch: fundamental NR + harmonic NR + post-processing.

Keys: `value` = converged solves/s with the default (structured Newton step) strategy, inputs
resident in HBM, CUDA-event timed on the launch stream, max over ranks; `dense_lu_path` = the
same workload forced through the dense shared-memory LU kernel (HPF_SOLVE_DENSE).  `e2e` = the same through the host-buffer C-ABI entry point
(hpf_solve_host): pinned host inputs -> H2D -> solve -> D2H of every result, plus (N > 1) the
final NCCL gather of flags and results.  `roofline` = the dominant kernel (fused Newton
kernel) against the FP64 pipe; `roofline_kernels` = the standalone mismatch / Jacobian kernels
against measured HBM bandwidth.  `cpu_baseline` = the oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NET, H_MAX, COUPLED, SPREAD = "net3", 25, True, "tight"
B_PER_GPU = 65536
METRIC = "converged harmonic-PF solves/sec (FP64)"
UNIT = "solves/s"


def load_net():
    """net3 + smps Norton equivalent from the committed fixtures (tests/golden/, SI units),
    through the product's own per-unit conversion (netio)."""
    import numpy as np
    from harmonic_power_flow_b200 import netio
    import pandas as pd
    g = os.path.join(ROOT, "tests", "golden")
    tab = json.load(open(os.path.join(g, "networks.json")))[NET]
    st = netio.Settings(H_MAX=H_MAX)
    buses, lines = pd.DataFrame(tab["buses"]), pd.DataFrame(tab["lines"])
    for col in ("S", "P", "Q"):
        buses[col] = buses[col].astype(float) / st.BASE_POWER
    buses["X_sh"] = buses["X_sh"].astype(float) / st.base_impedance
    for col in ("R", "X"):
        lines[col] = lines[col].astype(float) / st.base_impedance
    for col in ("G", "B"):
        lines[col] = lines[col].astype(float) / st.base_admittance
    nl = buses.index[buses["type"] == "nonlinear"]
    m, n = int(min(nl)), len(buses)
    c = int((buses.type == "PV").sum()) + 1
    dev = np.load(os.path.join(g, "ne_devices.npz"))
    freqs = np.array(st.HARMONICS_FREQ)
    col = np.searchsorted(dev["smps__freqs"], freqs)
    row = np.searchsorted(dev["smps__Y_N_c_rowfreq"], freqs)
    net = netio.pack_network(buses, lines, m, n, c, st.HARMONICS)
    net.devices = ["smps"]
    net.dev_of_nl_bus = np.zeros(n - m, np.int32)
    net.coupled = True
    net.Y_N = np.ascontiguousarray((dev["smps__Y_N_c"][np.ix_(row, col)] / st.base_admittance)[None])
    net.I_N = np.ascontiguousarray(np.repeat((dev["smps__I_N_c"][col] / st.base_current)[None], n - m, 0))
    return net


def load_other(kind):
    """Networks of the other BASELINE configurations through the product's own CSV loaders:
    'net2ev' (configs[1]: net2 with an SMPS and an EV Norton-equivalent load, odd harmonics to the
    19th), 'radial200' (configs[3]) and 'meshed1000' (configs[4]) from the seeded generators."""
    import tempfile
    import numpy as np
    import pandas as pd
    from harmonic_power_flow_b200 import netio, synthetic
    g = os.path.join(ROOT, "tests", "golden")
    tmp = tempfile.mkdtemp(prefix="hpf_bench_")
    dev = np.load(os.path.join(g, "ne_devices.npz"))
    for d in ("smps", "ev"):
        netio.write_ne_csv(os.path.join(tmp, d + "_NE.csv"), dev[d + "__freqs"], dev[d + "__Y_N_c"],
                           dev[d + "__I_N_c"], dev[d + "__Y_N_uc"], dev[d + "__I_N_uc"])
    if kind in ("net2ev", "net1"):
        tab = json.load(open(os.path.join(g, "networks.json")))[kind]
        pb, pl = os.path.join(tmp, kind + "_buses.csv"), os.path.join(tmp, kind + "_lines.csv")
        pd.DataFrame(tab["buses"]).to_csv(pb, sep=";", index=False)
        pd.DataFrame(tab["lines"]).to_csv(pl, sep=";", index=False)
        h_max = 19 if kind == "net2ev" else 51
    elif kind == "radial200":
        pb, pl = synthetic.radial_feeder(tmp, tmp, n=200, load_scale=0.005)
        h_max = 25
    else:
        pb, pl = synthetic.meshed(tmp, tmp, n=1000, load_scale=0.002)
        h_max = 25
    st = netio.Settings(H_MAX=h_max, ne_dir=tmp)
    buses, lines, m, n, c = netio.init_network(pb, pl, st)
    NE = netio.import_Norton_Equivalents(buses, True, st)
    return netio.pack_network(buses, lines, m, n, c, st.HARMONICS, NE, True)


OTHER_CONFIGS = [
    ("configs[0]: net1 (20 buses, 7 SMPS loads) coupled, odd harmonics <= 51 (N=1038), ONE scenario (latency)",
     "net1", 1, 5),
    ("configs[1]: net2 + SMPS/EV Norton loads, odd harmonics <= 19 (N=78)", "net2ev", 1024, 3),
    ("configs[3]: synthetic 200-bus radial feeder, 40% nonlinear buses, odd harmonics <= 25 (N=5198)",
     "radial200", 8192, 2),
    ("configs[4]: synthetic 1000-bus meshed network, 40% nonlinear buses, full coupled Norton, odd "
     "harmonics <= 25 (N=25998)", "meshed1000", 1024, 1),
]


def run_other_configs(torch, dist, BatchSolver, scenarios, local, rank, world):
    """Device-resident solves/s of the other BASELINE configurations (weak scaling like the
    headline: every rank solves its own block of scenarios of the same network).  The operator
    set-up (once per network) is timed separately."""
    out = []
    for label, kind, B, steps in OTHER_CONFIGS:
        net = load_other(kind)
        sol = BatchSolver(net, local)
        t0 = time.perf_counter()
        info = sol.struct_info()
        torch.cuda.synchronize()
        t_setup = time.perf_counter() - t0
        P, Q, I_N = scenarios.make_batch(net, B, SPREAD, seed0=rank * B)
        dP, dQ, dI = sol.prepare(P, Q, I_N)
        r = sol.solve(dP, dQ, dI)                                   # warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = sol.solve(dP, dQ, dI, out=r)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        vals = torch.tensor([ms, float((r.status == 0).sum().item()), r.n_iter_h.double().sum().item()],
                            dtype=torch.float64, device=r.status.device)
        if world > 1:
            mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        else:
            mx, sm = vals, vals
        out.append({"config": label, "batch_per_gpu": B, "N": int(sol.N), "strategy_variant": int(info["available"]),
                    "value": sm[1].item() / (mx[0].item() / 1e3), "unit": UNIT, "ms_per_step": mx[0].item(),
                    "converged_fraction": sm[1].item() / (B * world),
                    "mean_harmonic_iterations": sm[2].item() / (B * world),
                    "operator_setup_s": t_setup})
        sol.close()
        del sol, r, dP, dQ, dI
        torch.cuda.empty_cache()
    return out


def ncu_traffic(kernel_substr):
    """DRAM bytes (read + write) of one launch of a kernel from the committed ncu --set full summary
    (profiles/r1_ncu_kernels.csv), or None.  ncu numbers are per launch at the BASELINE batch."""
    import csv
    try:
        rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r1_ncu_kernels.csv"))))
        hdr, units = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        for r in rows[2:]:
            if kernel_substr in r[0]:
                return float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    except Exception:
        pass
    return None


def config_dict(n_gpus):
    return {"workload": "BASELINE configs[2]: net3 coupled (smps_NE), fundamental + odd harmonics <= 25 "
                        "(N=101), randomised load/spectrum scenarios (P,Q x U(0.9,1.1), I_N x U(0.95,1.05) "
                        "e^{jU(-0.02,0.02)}), thresh 1e-6/1e-4, caps 30/50",
            "batch_per_gpu": B_PER_GPU, "global_batch": B_PER_GPU * n_gpus,
            "parallelism": "scenario-sharded x%d, no data-path collective" % n_gpus,
            "cache": "L2 flushed (256 MiB write) between timed steps"}


# ------------------------------------------------------------------------------- CPU arm
_W = {}


def _cpu_init():
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hpf_oracle as O
    from harmonic_power_flow_b200 import scenarios
    on = O.net_from_golden(os.path.join(ROOT, "tests", "golden"), NET, H_MAX, COUPLED)
    _W.update(O=O, on=on, Y=O.build_admittance_matrices(on), draw=scenarios.draw_scenario)


def _cpu_worker(args):
    lo, hi, seed0 = args
    O, on, Y = _W["O"], _W["on"], _W["Y"]
    conv = iters = 0
    for s in range(lo, hi):
        P, Q, I_N = _W["draw"](seed0 + s, on.P, on.Q, on.I_N, SPREAD)
        o = O.hpf(on, P=P, Q=Q, I_N=I_N, Y=Y)
        conv += int(o["status"] == 0)
        iters += o["n_iter_h"]
    return conv, iters


class CpuPool:
    """Oracle port (numpy + SuperLU step, like the reference) fanned out over the host cores:
    one worker process per core, network/Y(h) set up once per worker (outside the timing),
    scenarios split evenly."""

    def __init__(self, procs):
        from multiprocessing import get_context
        self.procs = procs
        self.pool = get_context("fork").Pool(procs, initializer=_cpu_init)
        self.pool.map(_cpu_worker, [(0, 1, 0)] * procs)          # imports + first-call warm-up

    def solve(self, n_scen, seed0=0):
        chunk = max(1, n_scen // (self.procs * 4))
        jobs = [(lo, min(n_scen, lo + chunk), seed0) for lo in range(0, n_scen, chunk)]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
        return sum(r[0] for r in res), sum(r[1] for r in res), dt

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = 256 * cores           # ~12 ms per solve per core -> ~3 s per step
    pool = CpuPool(cores)
    for _ in range(a.warmup):
        pool.solve(min(per_step, 16 * cores))
    conv = iters = 0
    dt = 0.0
    for k in range(a.steps):
        c_, i_, d_ = pool.solve(per_step, seed0=k * per_step)
        conv += c_; iters += i_; dt += d_
    pool.close()
    v = conv / dt
    sample = "%d scenarios per step x %d steps of the same seeded workload, %d worker processes" % (
        per_step, a.steps, cores)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(a.gpus),
            "us_per_nr_iteration": dt * 1e6 * cores / max(iters, 1),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "oracle/hpf_oracle.py: numpy restatement of the reference solve path "
                                     "(SuperLU step like the reference); the reference's own pandas code "
                                     "measured 0.78 solves/s/core in the build container (BASELINE.md)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names)
                   if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------- GPU arm
def measure_fp64_peak(torch, dev):
    """FP64 denominator: cuBLAS DGEMM 8192^3 (uses the DMMA path) - MEASURED_PEAKS.json has no
    FP64 figure.  Burst (best of 5) in TFLOP/s."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2 * n ** 3 / best / 1e9


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from harmonic_power_flow_b200 import BatchSolver, scenarios
    from harmonic_power_flow_b200 import dist as hdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_pool = None
    if world == 1 and not a.no_cpu:        # fork the CPU workers BEFORE the CUDA context exists
        cpu_pool = CpuPool(os.cpu_count() or 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if a.gpus != world:
        a.gpus = world

    net = load_net()
    sol = BatchSolver(net, local)
    B = a.batch
    P, Q, I_N = scenarios.make_batch(net, B, SPREAD, seed0=rank * B)
    hP = torch.as_tensor(P).pin_memory(); hQ = torch.as_tensor(Q).pin_memory()
    hI = torch.as_tensor(I_N).pin_memory()
    dP, dQ, dI = sol.prepare(hP, hQ, hI)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    N = sol.N
    n, H, q, m, c = net.n, net.H, net.q, net.m, net.c
    Nf = 2 * n - 1 - c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ----
    out = None
    for _ in range(a.warmup):
        out = sol.solve(dP, dQ, dI, out=out)
    sol.set_profiling(True)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    l0 = sol.launch_count
    ev, kms = [], []
    t_wall0 = time.perf_counter()
    for _ in range(a.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = sol.solve(dP, dQ, dI, out=out); e1.record()
        ev.append((e0, e1))
        kms.append(sol.last_kernel_ms())          # waits for this step's kernels
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sol.launch_count - l0
    ms = sum(x.elapsed_time(y) for x, y in ev)
    ms_fund = sum(k[0] for k in kms) / a.steps
    ms_harm = sum(k[1] for k in kms) / a.steps
    conv = int((out.status == 0).sum().item())
    it_h = out.n_iter_h.double().sum().item()
    it_f = out.n_iter_f.double().sum().item()
    strategy = "structured" if sol.struct_info()["available"] else "dense"

    # ---- the same workload forced through the dense-LU kernel ----
    dense_info = None
    if rank == 0 and not a.no_dense:
        od = sol.solve(dP, dQ, dI, dense=True)
        torch.cuda.synchronize()
        tot, kd = 0.0, 0.0
        nd = 2
        for _ in range(nd):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); od = sol.solve(dP, dQ, dI, dense=True, out=od); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1); kd += sol.last_kernel_ms()[1]
        dense_info = dict(ms=tot / nd, kernel_ms=kd / nd, conv=int((od.status == 0).sum().item()),
                          it_h=od.n_iter_h.double().sum().item(), it_f=od.n_iter_f.double().sum().item(),
                          iter_mismatch_vs_structured=int((od.n_iter_h != out.n_iter_h).sum().item()))
        del od
    sol.set_profiling(False)

    # ---- end to end through the host-buffer C-ABI entry point (e2e) ----
    npP, npQ, npI = hP.numpy(), hQ.numpy(), hI.numpy()
    # N > 1: the results of a rank live in ONE contiguous device slab (fields = views), so that the
    # final gather is one NCCL collective and the copy back to the host one D2H - with several
    # processes per host the number of driver calls per step matters
    res_e2e = slab = host_slab = None
    d2h_stream = torch.cuda.Stream()
    if world > 1:
        from harmonic_power_flow_b200 import solver as hsolver
        res_e2e, slab = sol.alloc_result_slab(B)
        host_slab = torch.empty(slab.numel(), dtype=torch.uint8).pin_memory()
    for _ in range(max(1, a.warmup - 1)):
        r = sol.solve_host(npP, npQ, npI)
        if world > 1:
            sol.solve(dP, dQ, dI, out=res_e2e)
            hdist.gather_slab(slab)
            host_slab.copy_(slab)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        if world == 1:
            r = sol.solve_host(npP, npQ, npI)
        else:
            # shard in (pinned H2D), solve into the slab, then the single collective of the path (final
            # NCCL gather of flags + results, rank-major, on the compute stream) OVERLAPPED with the
            # copy of this rank's own slab back to pinned host memory on a second stream (NVLink and
            # PCIe are independent links)
            sol.solve(hP.to(dev, non_blocking=True), hQ.to(dev, non_blocking=True),
                      hI.to(dev, non_blocking=True), out=res_e2e)
            done = torch.cuda.Event(); done.record()
            gathered = hdist.gather_slab(slab)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                host_slab.copy_(slab, non_blocking=True)
                d2h_stream.synchronize()
            torch.cuda.current_stream().synchronize()
            rr = hsolver.result_from_slab(host_slab, net.H, net.n, net.q, B)
            r = {k: getattr(rr, k).numpy() for k in ("V_m", "V_a", "I_inj", "n_iter_f", "n_iter_h", "err_h", "status")}
    barrier()
    t_e2e = time.perf_counter() - t0
    clk = clocks.stop()       # sampled over the device-timed AND the end-to-end region
    conv_e2e = int((r["status"] == 0).sum())
    h2d = npP.nbytes + npQ.nbytes + npI.nbytes
    d2h = sum(r[k].nbytes for k in ("V_m", "V_a", "I_inj", "n_iter_f", "n_iter_h", "err_h", "status"))

    # ---- standalone kernels against the HBM roofline (rank 0) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks \
        else (6650.0, "fallback (B200_PROFILING.md)")
    kernels = []
    fp64_peak = None
    if rank == 0:
        def timed(fns, reps=12):
            """Average device time of one launch: `reps` launches captured in ONE CUDA graph and
            replayed between two events (no host launch overhead between the kernels: the Python
            wrapper costs ~30 us per call, as much as the mismatch kernel itself), rotating over
            input/output sets whose total footprint exceeds the 126 MB L2; best of 5 replays."""
            side = torch.cuda.Stream()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                for fn in fns:
                    fn()
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=side):
                    for r in range(reps):
                        fns[r % len(fns)]()
            torch.cuda.synchronize()
            best = float("inf")
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                flush.fill_(1)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / reps)
            del g
            return best
        raw = sol.solve(dP, dQ, dI, raw=True, max_iter_h=3, want_I_inj=False)   # a mid-iteration state
        Vm, Va = raw.V_m, raw.V_a
        nset = 3                                             # 3 x 124 MB of traffic per rotation
        sets = [(Vm.clone(), Va.clone(), dP.clone(), dQ.clone(), dI.clone(),
                 (torch.empty((N, B), dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.float64, device=dev)))
                for _ in range(nset)]
        t_mis = timed([(lambda s_=s_: sol.mismatch(s_[0], s_[1], s_[2], s_[3], s_[4], out=s_[5])) for s_ in sets])
        del sets
        by_mis = 16 * n * H + 8 * N + 16 * q * H + 16 * (m - 1) + 8
        kernels.append({"kernel": "mismatch_lane_kernel (one thread per scenario)", "bound": "hbm",
                        "ms": t_mis,
                        "achieved": by_mis * B / t_mis / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": by_mis * B / t_mis / 1e6 / hbm_peak, "bytes_per_scenario": by_mis,
                        "traffic": ncu_traffic("mismatch_lane_kernel"),
                        "timing": "CUDA graph of 12 launches over 3 rotating buffer sets (372 MB > L2), best of 5 replays"})
        Bj = min(B, 16384)
        Vmj, Vaj = Vm[:, :, :Bj].contiguous(), Va[:, :, :Bj].contiguous()
        Js = [sol.jacobian(Vmj, Vaj) for _ in range(2)]      # 2 x 1.35 GB outputs
        t_jac = timed([(lambda J_=J_: sol.jacobian(Vmj, Vaj, out=J_)) for J_ in Js], reps=6)
        J = Js[0]
        del Js
        by_jac = 8 * N * N + 16 * n * H
        kernels.append({"kernel": "jacobian_kernel", "bound": "hbm", "ms": t_jac, "batch": Bj,
                        "achieved": by_jac * Bj / t_jac / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": by_jac * Bj / t_jac / 1e6 / hbm_peak, "bytes_per_scenario": by_jac,
                        "traffic": ncu_traffic("jacobian_kernel")})
        f, _ = sol.mismatch(Vmj, Vaj, dP[:, :Bj].contiguous(), dQ[:, :Bj].contiguous(), dI[:, :, :Bj].contiguous())
        t_lu = timed([lambda: sol.lu_solve(J, f)], reps=3)
        fl_lu = 2.0 / 3.0 * N ** 3 + 2.0 * N * N
        del J
        fp64_peak = measure_fp64_peak(torch, dev)
        kernels.append({"kernel": "lu_solve_kernel", "bound": "fp64", "ms": t_lu, "batch": Bj,
                        "achieved": fl_lu * Bj / t_lu / 1e9, "peak": fp64_peak, "unit": "TFLOP/s",
                        "frac": fl_lu * Bj / t_lu / 1e9 / fp64_peak, "flops_per_scenario": fl_lu})

    # ---- the other BASELINE configurations (all ranks, weak scaling) ----
    other = None
    if not a.no_other:
        other = run_other_configs(torch, dist, BatchSolver, scenarios, local, rank, world)

    # ---- reduce over ranks ----
    vals = torch.tensor([ms, t_e2e, float(conv), float(conv_e2e), it_h, it_f, t_wall], dtype=torch.float64, device=dev)
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = vals, vals
    ms_max, e2e_max = mx[0].item(), mx[1].item()
    conv_tot, conv_e2e_tot, it_h_tot, it_f_tot = sm[2].item(), sm[3].item(), sm[4].item(), sm[5].item()

    if rank == 0:
        value = conv_tot * a.steps / (ms_max / 1e3)
        e2e_v = conv_e2e_tot * a.steps / e2e_max
        # algorithmic FP64 work per Newton iteration of one scenario (DESIGN.md "Roofline accounting")
        nZ = n * H - m
        fl_mis = 8 * (H * n * n + q * H * H + (m - 1) * n)
        # structured step: u_Z = -V_Z - G (V_F + u_F) - w_N, polar conversion, small border system
        fl_struct = fl_mis + 8 * nZ * m + 10 * n * H + 8 * (2 * m) ** 2
        fl_dense = fl_mis + 16 * (H * n * n + q * H * H) + 2.0 / 3.0 * N ** 3 + 2.0 * N * N
        fl_f = 24 * n * n + 2.0 / 3.0 * Nf ** 3 + 2.0 * Nf * Nf
        if strategy == "structured":
            kname = "harm_tile_kernel (harmonic Newton, structured step, 32 scenarios per CTA)"
            # n_iter+1 mismatches, n_iter steps (sincos: 52 per round, NOT counted as flops)
            flops = (it_h + B) * fl_mis + it_h * (fl_struct - fl_mis)
            t_kernel = ms_harm / 1e3
            fl_iter = fl_struct
        else:
            kname = "solve_kernel<0> (fused fundamental + harmonic Newton, dense LU)"
            flops = it_h * fl_dense + it_f * fl_f
            t_kernel = ms_harm / 1e3
            fl_iter = fl_dense
        ach = flops / t_kernel / 1e12
        by_solve = 16 * n + 16 * q * H + 16 * n * H + 16 * q * H + 24 + 2 * 16 * n   # in + out + fundamental hand-over
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(world), "strategy": strategy,
                "converged_fraction": conv_tot / (B * world),
                "mean_nr_iterations": {"fundamental": it_f_tot / (B * world), "harmonic": it_h_tot / (B * world)},
                "us_per_nr_iteration": (ms_max / a.steps) * 1e3 / max(it_h / 1.0, 1.0),
                "gpu_launches": int(launches),
                "kernel_ms_per_step": {"fundamental_stage": ms_fund, "harmonic_stage": ms_harm},
                "wall_s_timed_region": t_wall,
                "clocks": clk,
                "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_max / a.steps * 1e3,
                        "path": ("BatchSolver.solve_host -> hpf_solve_host (C ABI, host buffers)" if world == 1 else
                                 "per rank: pinned H2D of its shard -> hpf_solve into one contiguous result slab -> "
                                 "ONE NCCL all_gather of the slabs (rank-major) overlapped with ONE D2H of the "
                                 "rank's own slab")},
                "roofline": {"kernel": kname, "bound": "fp64", "achieved": ach, "peak": fp64_peak,
                             "unit": "TFLOP/s", "frac": (ach / fp64_peak) if fp64_peak else None,
                             "traffic": ncu_traffic("harm_tile_kernel") if strategy == "structured" else None,
                             "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu "
                                               "--set full (profiles/r1_ncu_kernels.csv)",
                             "peak_source": "cuBLAS DGEMM 8192^3 burst measured in this run (no FP64 figure "
                                            "in MEASURED_PEAKS.json)",
                             "flops_per_nr_iteration": fl_iter,
                             "sincos_per_nr_iteration": n * H,
                             "note": "latency/issue-bound kernel: flops exclude the FP64 sincos per phasor; see "
                                     "profiles/ for issue-slot and FP64-pipe utilisation from ncu",
                             "hbm_bytes_per_scenario": by_solve,
                             "hbm_gbs_of_whole_solve": by_solve * B / (ms / a.steps) / 1e6},
                "roofline_kernels": kernels, "hbm_peak_source": hbm_src}
        if other is not None:
            line["other_configs"] = other
        if dense_info is not None:
            fld = dense_info["it_h"] * fl_dense + dense_info["it_f"] * fl_f
            achd = fld / (dense_info["kernel_ms"] / 1e3) / 1e12
            line["dense_lu_path"] = {
                "value": dense_info["conv"] / (dense_info["ms"] / 1e3), "unit": UNIT,
                "ms_per_step": dense_info["ms"],
                "iteration_count_mismatches_vs_structured": dense_info["iter_mismatch_vs_structured"],
                "roofline": {"kernel": "solve_kernel<0> (fused Newton, dense smem LU, warp-shuffle pivoting)",
                             "bound": "fp64", "achieved": achd, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": achd / fp64_peak if fp64_peak else None,
                             "flops_per_nr_iteration": fl_dense}}
        if cpu_pool is not None:
            cores = cpu_pool.procs
            nscen = 1024 * cores
            c_, i_, d_ = cpu_pool.solve(nscen)
            cpu_pool.close()
            line["cpu_baseline"] = {
                "value": c_ / d_, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "first %d scenarios of the same seeded workload, %d worker processes, %.1f s" % (
                    nscen, cores, d_),
                "us_per_nr_iteration_per_core": d_ * 1e6 * cores / max(i_, 1)}
        print(json.dumps(line))
    sol.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="scenarios per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-LU comparison leg")
    ap.add_argument("--no-other", action="store_true", help="skip the other BASELINE configurations")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
