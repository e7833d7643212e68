#!/usr/bin/env python
"""bench.py -- atom-environments/sec for ACE energy + forces (FP64), BASELINE.json's metric.

Workload (config.workload): BASELINE config 2 -- LinearACEModel, Invariant, ord = 3, maxdeg = 12, wL = 1.5
(SparseBasis), 40 random neighbours per environment (rho ~ U[rin, rcut], direction uniform), 10^6
environments per GPU, synthetic (Philox-seeded) positions and random coefficients.

One "step" = one evaluate + grad_config pass over the whole batch (every environment's energy and all
40 x 3 force components).  `value` times K steps with the batch resident in HBM; `e2e` times the same
call through the C ABI with PINNED HOST buffers (host->device copy of positions and offsets and
device->host copy of energies and forces inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); environments are independent, so each rank evaluates its
own shard (weak scaling: E environments per GPU) and the only collective is the all-reduce of the total
energy (NCCL), issued inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "atom-environments/sec (energy+forces, FP64)"
UNIT = "env/s"
J = 40
ORD, MAXDEG, WL = 3, 12, 1.5
SEED = 20242


def build_model(nprop: int = 1):
    import ace_jl_b200 as ace
    from ace_jl_b200.utils import RnYlm_1pbasis, philox
    Bsel = ace.SparseBasis(maxorder=ORD, p=1, default_maxdeg=MAXDEG, weight={"n": 1.0, "l": WL})
    B1p = RnYlm_1pbasis(maxdeg=MAXDEG, maxL=math.ceil(MAXDEG / WL), Bsel=Bsel)
    basis = ace.SymmetricBasis(ace.Invariant(), B1p, Bsel)
    c = philox(SEED + 1000).random(len(basis)) - 0.5
    return basis, c


def algorithmic_flops(basis, nJ: int, P: int = 1):
    """SURVEY.md section 8(d), split by kernel."""
    b1p = basis.pibasis.basis1p
    Nn = len(b1p.component(0).R)
    L = max(b[b1p.sym_index("l")] for b in b1p.spec)
    sizeP, sizeY, nA = (L + 1) * (L + 2) // 2, (L + 1) ** 2, len(b1p)
    orders = basis.pibasis.spec.orders
    f = lambda nu: 6 * (nu - 1) + 18 * max(nu - 2, 0) + (4 * nu + 2) * P  # noqa: E731
    pool = nJ * (28 + 5 * Nn + 15 * sizeP + 4 * nA)
    adj = float(sum(int((orders == nu).sum()) * f(nu) for nu in range(1, int(orders.max()) + 1)))
    forces = nJ * (6 * Nn + 22 * sizeP + (8 * nA + 16 * sizeY) * P)
    return {"pool": float(pool), "adjoint": adj, "forces": float(forces)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 10 and r[2].replace(".", "").isdigit()]
        # samples taken inside the timed region (a sample reports the ~100 ms before it was printed)
        inreg = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or 1e30) + 0.12]
        scope = "timed region"
        if not inreg:   # region shorter than the sampling period: fall back to everything under load since warm-up
            inreg, scope = rows, "warm-up + timed region"
        sm = [float(r[2]) for r in inreg]
        smax = [float(r[3]) for r in inreg]
        reasons = set()
        for r in inreg:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


def cpu_reference_rate(basis, c, nenv_sample: int, repeats: int = 1):
    """The restated reference (oracle, C + OpenMP over environments) on this box's host cores."""
    import oracle as orc
    from ace_jl_b200.descriptor import basis_descriptor
    from ace_jl_b200.utils import philox, rand_envs
    o = orc.Oracle(basis_descriptor(basis, c.reshape(-1, 1)))
    o.set_threads(len(os.sched_getaffinity(0)))   # all host cores (torchrun exports OMP_NUM_THREADS=1)
    R, off, _ = rand_envs(philox(SEED + 7), basis.pibasis.basis1p.component(0), nenv_sample, J)
    o.energy_forces(R[: J * 64], off[:65])  # warm-up (thread pool, page faults)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        o.energy_forces(R, off)
        best = min(best, time.perf_counter() - t0)
    return nenv_sample / best, o.num_threads(), best


def run_reference_arm(args):
    """--impl reference: the reference's own CPU path.  ACE.jl is Julia and cannot be installed here
    (no Julia, no registry), so this arm times the line-by-line C restatement in oracle/ -- same
    algorithm incl. the materialised dA matrix and the full (maxL+1)^2 harmonics -- with OpenMP over
    environments standing in for `Threads.@threads` over configurations, on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    basis, c = build_model()
    sample = args.cpu_envs
    times = []
    cores = 1
    for s in range(args.warmup + args.steps):
        rate, cores, dt = cpu_reference_rate(basis, c, sample)
        if s >= args.warmup:
            times.append(dt)
    value = sample * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.envs, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} environments x {J} neighbours per step (bounded sample of the 1e6-environment workload)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(nenv, ngpu):
    return {"workload": f"LinearACEModel energy+forces, Invariant, ord={ORD}, maxdeg={MAXDEG}, wL={WL} SparseBasis, "
                        f"{J} neighbours, {nenv} environments per GPU (BASELINE config 2)",
            "envs_per_gpu": nenv, "neighbours": J, "parallelism": f"env-shard x{ngpu}",
            "l2_policy": "inputs (24*J B/env) and outputs exceed the 126 MB L2: no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs", type=int, default=1_000_000, help="environments per GPU")
    ap.add_argument("--cpu-envs", type=int, default=100_000, help="environments in the CPU baseline sample")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import ace_jl_b200 as ace
    from ace_jl_b200.api import measure_fp64_tflops
    from ace_jl_b200.utils import philox, rand_envs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ace_jl_b200 has no CPU path")
    torch.cuda.set_device(local)
    from ace_jl_b200 import _lib
    _lib.check(_lib.load().aceb200_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    basis, c = build_model()
    model = ace.LinearACEModel(basis, c)
    h = model.evaluator.handle
    nenv = args.envs
    rng = philox(SEED + 17 * rank)
    R, off, _ = rand_envs(rng, basis.pibasis.basis1p.component(0), nenv, J)
    Rd, offd = torch.from_numpy(R).to(dev), torch.from_numpy(off).to(dev)
    batch = ace.B200Batch(Rd, offd)
    E = torch.empty((nenv, 1, 1), dtype=torch.float64, device=dev)
    G = torch.empty((nenv * J, 1, 3, 1), dtype=torch.float64, device=dev)
    etot = torch.zeros(1, dtype=torch.float64, device=dev)

    def step():
        h.energy_forces(batch, E, G)
        etot.copy_(E.sum().reshape(1))
        if world > 1:
            dist.all_reduce(etot)          # the one collective of the path: total energy over all shards

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = measure_fp64_tflops()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    stage = {"pool": 0.0, "adjoint": 0.0, "forces": 0.0}
    l0 = h.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step()
        for k, v in h.last_stage_ms().items():
            stage[k] += v
    e1.record()
    barrier()
    sampler.mark_end()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    launches = h.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = world * nenv * args.steps / (total_ms * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers
    Rh = torch.from_numpy(R).pin_memory()
    offh = torch.from_numpy(off).pin_memory()
    Eh = torch.empty((nenv, 1, 1), dtype=torch.float64).pin_memory()
    Gh = torch.empty((nenv * J, 1, 3, 1), dtype=torch.float64).pin_memory()
    hb = ace.B200Batch(Rh.numpy(), offh.numpy())
    e2e_steps = max(1, min(args.steps, 10))
    h.energy_forces(hb, Eh.numpy(), Gh.numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h.energy_forces(hb, Eh.numpy(), Gh.numpy())   # returns after the D2H copies have completed
        _ = float(Eh[0, 0, 0])
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * nenv * e2e_steps / float(dt.item())
    assert np.array_equal(Eh.numpy(), E.cpu().numpy()), "host-buffer path and device-resident path disagree"

    # ---- end to end through the caller-side entry: a whole periodic structure with its neighbour
    # list in pinned host memory -> site energies, atomic forces and the virial back in pinned host memory.  The
    # environments are built and the forces assembled on the device, so the pair gradients never cross PCIe.
    from ace_jl_b200.structure import B200Structure
    from ace_jl_b200.utils import fcc_structure
    ncell = max(2, round((nenv / 4.0) ** (1.0 / 3.0)))
    sX, scell, sfirst, snbr, simg = fcc_structure(philox(SEED + 99 + rank), ncell)
    pin = lambda a: torch.from_numpy(a).pin_memory()   # noqa: E731
    pX, pfirst, pnbr, pimg = pin(sX), pin(sfirst), pin(snbr), pin(simg)
    st = B200Structure(pX.numpy(), pfirst.numpy(), pnbr.numpy(), pimg.numpy(), scell)
    sE = torch.empty((st.natoms, 1, 1), dtype=torch.float64).pin_memory()
    sF = torch.empty((st.natoms, 1, 3, 1), dtype=torch.float64).pin_memory()
    sW = torch.empty((1, 3, 3), dtype=torch.float64).pin_memory()
    h.structure_energy_forces(st, True, sE.numpy(), sF.numpy(), sW.numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h.structure_energy_forces(st, True, sE.numpy(), sF.numpy(), sW.numpy())   # returns after the D2H copies
        _ = float(sE[0, 0, 0])
    torch.cuda.synchronize()
    dts = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dts, op=dist.ReduceOp.MAX)
    struct_value = world * st.natoms * e2e_steps / float(dts.item())
    struct_h2d = sX.nbytes + sfirst.nbytes + snbr.nbytes + simg.nbytes
    struct_d2h = sE.numel() * 8 + sF.numel() * 8 + sW.numel() * 8
    fsum = float(np.abs(sF.numpy().sum(axis=0)).max() / np.abs(sF.numpy()).max())
    assert fsum < 1e-9, "forces of a periodic structure must sum to zero"

    if rank == 0:
        flops = algorithmic_flops(basis, J)
        per_launch_ms = {k: v / args.steps for k, v in stage.items()}
        dom = max(per_launch_ms, key=per_launch_ms.get)
        ach = flops[dom] * nenv / (per_launch_ms[dom] * 1e-3) / 1e12
        tot_flops = sum(flops.values())
        hbm_bytes = 24 * J + 24 * J + 8 + 8
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/), per launch
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            kname = {"pool": "k_pool", "adjoint": "k_adjoint_stream", "forces": "k_forces"}[dom]
            t = tr[kname]
            traffic = (t["dram_read_bytes"] + t["dram_write_bytes"]) * nenv / t["envs_per_launch"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(nenv, world),
            "roofline": {
                "bound": "fp64", "kernel": {"pool": "k_pool", "adjoint": "k_adjoint", "forces": "k_forces"}[dom],
                "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None,
                "traffic": traffic,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, profiles/r1_traffic.json (includes the "
                                  "[slot][env] workspace the three kernels exchange: 1.2 KB/env each way)",
                "peak_source": "FP64 FMA probe run in this process (MEASURED_PEAKS.json holds no FP64 figure)",
                "algorithmic_flops_per_env": flops, "ms_per_launch": per_launch_ms,
                "whole_step_tflops": tot_flops * nenv / (total_ms / args.steps * 1e-3) / 1e12,
                "hbm": {"achieved": value / world * hbm_bytes / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": value / world * hbm_bytes / 1e9 / hbm_peak, "algorithmic_bytes_per_env": hbm_bytes,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"},
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(R.nbytes + off.nbytes),
                    "d2h_bytes_per_step": int(Eh.numel() * 8 + Gh.numel() * 8), "steps": e2e_steps,
                    "timer": "host wall clock around the C-ABI calls (they return after the D2H copy)"},
            "e2e_structure": {"value": struct_value, "unit": UNIT, "h2d_bytes_per_step": int(struct_h2d), "d2h_bytes_per_step": int(struct_d2h),
                              "steps": e2e_steps, "atoms_per_gpu": st.natoms, "pairs_per_gpu": st.npairs,
                              "workload": "aceb200_structure_energy_forces: jittered periodic FCC crystal, 42 neighbours per atom inside "
                                          "rcut, positions + neighbour list (i, j, S) in pinned host memory -> site energies, atomic "
                                          "forces, virial in pinned host memory (same model; environments built and forces assembled on "
                                          "the device)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if not args.no_cpu:
            rate, cores, dtc = cpu_reference_rate(basis, c, args.cpu_envs)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_envs} environments x {J} neighbours, {dtc:.1f} s, restated reference "
                                              "(C + OpenMP over environments, materialised dA like src/evaluator.jl:169)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
