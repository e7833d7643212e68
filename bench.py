#!/usr/bin/env python
"""bench.py -- atom-environments/sec of the ACE evaluation hot path (FP64), BASELINE.json's metric.

  python bench.py [--config 2] [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl reference]

--config selects one of BASELINE.json's configurations (ace_jl_b200/workloads.py); the default, 2, is the one the
metric is quoted on: LinearACEModel energy+forces, Invariant, ord = 3, maxdeg = 12, wL = 1.5 SparseBasis, 40 random
neighbours per environment, 10^6 environments per GPU, synthetic (Philox-seeded) positions, random coefficients.
  1   evaluate(basis::SymmetricBasis, cfg) (B values), ord 3 / deg 10, 30 neighbours
  3   energy+forces, ord 4 / deg 14, 60 neighbours (north_star's 8-GPU configuration: --config 3 --gpus 8)
  4a / 4   B values of the EuclideanVector / EuclideanMatrix bases, ord 3 / deg 10
  5 / 5f   16-property, 4-species model: energies / energies + 16 force fields

One "step" = one pass of the call over the whole batch.  `value` times K steps with the batch resident in HBM;
`e2e` times the call a user makes with PINNED HOST buffers (host->device copy of the inputs and device->host copy
of the results inside the timed region).  For the energy+forces configurations the primary `e2e` is the
caller-shaped entry point aceb200_structure_energy_forces (positions + neighbour list in, site energies + atomic
forces + virial out); the per-environment call (48 B per pair over PCIe, the reference's own call shape) is reported
beside it as `e2e_per_environment`.

N > 1 is launched by torchrun (one rank per GPU); environments are independent, so each rank evaluates its own
shard (weak scaling: E environments per GPU) and the only collective is the all-reduce of the total energy (NCCL),
issued inside the timed region.
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

import numpy as np  # noqa: E402

UNIT = "env/s"
METRICS = {"dB": "atom-environments/sec (basis Jacobian dB, FP64)",
           "EF": "atom-environments/sec (energy+forces, FP64)",
           "E": "atom-environments/sec (energies, FP64)",
           "B": "atom-environments/sec (basis values, FP64)"}


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


def make_inputs(w, basis, nenv, seed):
    from ace_jl_b200.utils import philox, rand_envs
    return rand_envs(philox(seed), basis.pibasis.basis1p.component(0), nenv, w.J, w.nspecies)


def cpu_reference_rate(w, basis, c, nenv_sample: int, min_seconds: float = 10.0):
    """The restated reference (oracle, C + OpenMP over environments) on this box's host cores, same call."""
    import oracle as orc
    from ace_jl_b200.descriptor import basis_descriptor
    o = orc.Oracle(basis_descriptor(basis, c))
    o.set_threads(len(os.sched_getaffinity(0)))   # all host cores (torchrun exports OMP_NUM_THREADS=1)
    R, off, sp = make_inputs(w, basis, nenv_sample, w.seed + 7)
    fn = {"EF": o.energy_forces, "E": o.energy, "B": o.eval_B, "dB": o.eval_dB}[w.call]
    nw = min(64, nenv_sample)
    fn(R[: w.J * nw], off[: nw + 1], None if sp is None else sp[: w.J * nw])  # warm-up (thread pool, page faults)
    # repeat the pass over the sample until min_seconds of CPU work have been timed (memory bounds the sample, not time)
    passes, dt = 0, 0.0
    while (dt < min_seconds and passes < 64) or passes == 0:
        t0 = time.perf_counter()
        fn(R, off, sp)
        dt += time.perf_counter() - t0
        passes += 1
    return nenv_sample * passes / dt, o.num_threads(), dt, passes


def default_cpu_sample(w, basis) -> int:
    """About 10 s of CPU work on 16 cores for the call (measured rates: config 2 ~ 1e5 env/s)."""
    from ace_jl_b200.workloads import algorithmic_work
    fl = algorithmic_work(basis, w.J, w.call, w.nprop)["flops_total"] * (10.0 if w.call == "EF" else 3.0)   # the oracle materialises dA
    n = max(500, min(100_000, 2.0e11 / fl))
    if w.call == "dB":                              # the Jacobian of the sample must fit in host memory: <= 4 GB
        n = max(50, min(n, 4.0e9 / algorithmic_work(basis, w.J, w.call, w.nprop)["out_bytes"]))
    return int(n)


def workload_config(w, nenv, ngpu):
    return {"workload": f"{w.title}, {nenv} environments per GPU", "baseline_config": w.key,
            "envs_per_gpu": nenv, "neighbours": w.J, "nprop": w.nprop, "parallelism": f"env-shard x{ngpu}",
            "l2_policy": "inputs (24*J B/env) and outputs exceed the 126 MB L2: no flush needed"}


def run_reference_arm(args, w):
    """--impl reference: the reference's own CPU path.  ACE.jl is Julia and cannot be installed here
    (no Julia, no registry), so this arm times the line-by-line C restatement in oracle/ -- same
    algorithm incl. the materialised dA matrix and the full (maxL+1)^2 harmonics -- with OpenMP over
    environments standing in for `Threads.@threads` over configurations, on all host cores."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from ace_jl_b200.workloads import build_basis, coefficients
    basis = build_basis(w)
    c = coefficients(w, basis)
    sample = args.cpu_envs or default_cpu_sample(w, basis)
    times, cores = [], 1
    for s in range(args.warmup + args.steps):
        _, cores, dt, passes = cpu_reference_rate(w, basis, c, sample, min_seconds=1.0)      # a step = >= 1 s of passes over the sample
        if s >= args.warmup:
            times.append(dt / passes)
    value = sample * len(times) / sum(times)
    print(json.dumps({
        "impl": "reference", "metric": METRICS[w.call], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(w, args.envs or w.nenv, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} environments x {w.J} neighbours per step (bounded sample of the workload)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="2", help="BASELINE configuration: 1, 1d (evaluate_d), 2, 3, 4a, 4, 4ad / 4d (evaluate_d), 5, 5f")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--envs", type=int, default=0, help="environments per GPU (default: the configuration's)")
    ap.add_argument("--cpu-envs", type=int, default=0, help="environments in the CPU baseline sample (default: ~10 s of work)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs")
    args = ap.parse_args()
    from ace_jl_b200.workloads import WORKLOADS, algorithmic_work, build_basis, coefficients
    if args.config not in WORKLOADS:
        raise SystemExit(f"--config must be one of {sorted(WORKLOADS)}")
    w = WORKLOADS[args.config]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args, w)

    import torch
    import torch.distributed as dist
    import ace_jl_b200 as ace
    from ace_jl_b200.api import measure_dmma_tflops, measure_fp64_tflops

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

    basis = build_basis(w)
    c = coefficients(w, basis)
    model = ace.LinearACEModel(basis, c if w.nprop > 1 else c[:, 0])
    h = model.evaluator.handle
    nenv = args.envs or w.nenv
    J, P = w.J, w.nprop * basis.A2Bmap.ncomp
    R, off, sp = make_inputs(w, basis, nenv, w.seed + 17 * rank)
    t = lambda a: None if a is None else torch.from_numpy(a).to(dev)   # noqa: E731
    batch = ace.B200Batch(t(R), t(off), t(sp))
    nB, ncomp = len(basis), basis.A2Bmap.ncomp
    if w.call == "B":
        out = [torch.empty((nenv, nB, ncomp), dtype=torch.float64 if basis.real else torch.complex128, device=dev)]
        call = lambda b, o: h.eval_B(b, o[0])   # noqa: E731
    elif w.call == "dB":
        dt = torch.float64 if basis.real else torch.complex128
        out = [torch.empty((nenv, nB, ncomp), dtype=dt, device=dev), torch.empty((nenv * J, nB, 3, ncomp), dtype=dt, device=dev)]
        call = lambda b, o: h.eval_dB(b, o[0], o[1])   # noqa: E731
    elif w.call == "E":
        out = [torch.empty((nenv, w.nprop, ncomp), dtype=torch.float64, device=dev)]
        call = lambda b, o: h.energy(b, o[0])   # noqa: E731
    else:
        out = [torch.empty((nenv, w.nprop, ncomp), dtype=torch.float64, device=dev),
               torch.empty((nenv * J, w.nprop, 3, ncomp), dtype=torch.float64, device=dev)]
        call = lambda b, o: h.energy_forces(b, o[0], o[1])   # noqa: E731
    tot = torch.zeros(1, dtype=torch.float64, device=dev)

    def step():
        call(batch, out)
        if w.call not in ("B", "dB"):
            tot.copy_(out[0].sum().reshape(1))
            if world > 1:
                dist.all_reduce(tot)          # the one collective of the path: total energy over all shards

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = measure_fp64_tflops()
    dmma_peak = measure_dmma_tflops()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    stage = {}
    l0 = h.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step()
        for k, v in h.last_stage_ms().items():
            k = "basis" if (w.call in ("B", "dB") and k == "adjoint") else k      # the second launch of a B call is the fused value kernel
            k = "jacobian" if (w.call == "dB" and k == "forces") else k           # ... and what follows it in a dB call: k_dA + k_dB_fused
            stage[k] = stage.get(k, 0.0) + v
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

    # ---- sampled parity against the CPU oracle (outside every timed region): the numbers above are only worth
    # reporting if the kernels computed the right thing at this size
    parity = None
    if rank == 0:
        from ace_jl_b200.descriptor import basis_descriptor
        from oracle import Oracle
        ns = min(nenv, 100 if len(basis.pibasis) > 20000 else 400)
        sel = np.sort(np.random.default_rng(5).choice(nenv, size=ns, replace=False))
        Rs = np.concatenate([R[off[e]:off[e + 1]] for e in sel])
        sps = None if sp is None else np.concatenate([sp[off[e]:off[e + 1]] for e in sel])
        offs = np.arange(ns + 1, dtype=np.int64) * J
        o = Oracle(basis_descriptor(basis, c))
        o.set_threads(len(os.sched_getaffinity(0)))
        rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))   # noqa: E731
        if w.call == "B":
            errs = {"B": rel(out[0].cpu().numpy()[sel], o.eval_B(Rs, offs, sps))}
        elif w.call == "dB":
            Bo, dBo = o.eval_dB(Rs, offs, sps)
            dBg = out[1].cpu().numpy().reshape(nenv, J, *out[1].shape[1:])[sel].reshape(ns * J, *out[1].shape[1:])
            errs = {"B": rel(out[0].cpu().numpy()[sel], Bo), "dB": rel(dBg, dBo)}
        elif w.call == "E":
            errs = {"E": rel(out[0].cpu().numpy()[sel], o.energy(Rs, offs, sps))}
        else:
            Eo, Go = o.energy_forces(Rs, offs, sps)
            G = out[1].cpu().numpy().reshape(nenv, J, *out[1].shape[1:])[sel].reshape(ns * J, *out[1].shape[1:])
            errs = {"E": rel(out[0].cpu().numpy()[sel], Eo), "G": rel(G, Go)}
        parity = {"ok": all(v < 1e-12 for v in errs.values()), "rel_err_vs_oracle": errs, "sampled_envs": int(ns), "tol": 1e-12}

    # ---- end to end through the C ABI with pinned host buffers (the per-environment call, the reference's call shape)
    e2e_env = e2e_struct = None
    e2e_steps = max(1, min(args.steps, 10))
    if not args.no_e2e:
        pin = lambda a: None if a is None else torch.from_numpy(a).pin_memory()   # noqa: E731
        Rh, offh, sph = pin(R), pin(off), pin(sp)
        outh = [torch.empty(o_.shape, dtype=o_.dtype).pin_memory() for o_ in out]
        hb = ace.B200Batch(Rh.numpy(), offh.numpy(), None if sph is None else sph.numpy())
        outn = [o_.numpy() for o_ in outh]
        call(hb, outn)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            call(hb, outn)                       # returns after the D2H copies have completed
            _ = outh[0].view(-1)[0].item()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert np.array_equal(outn[0], out[0].cpu().numpy()), "host-buffer path and device-resident path disagree"
        e2e_env = {"value": world * nenv * e2e_steps / float(dt.item()), "unit": UNIT,
                   "h2d_bytes_per_step": int(R.nbytes + off.nbytes + (0 if sp is None else sp.nbytes)),
                   "d2h_bytes_per_step": int(sum(o_.numel() * o_.element_size() for o_ in outh)), "steps": e2e_steps,
                   "call": {"EF": "aceb200_energy_forces", "E": "aceb200_energy", "B": "aceb200_eval_B", "dB": "aceb200_eval_dB"}[w.call],
                   "timer": "host wall clock around the C-ABI calls (they return after the D2H copy)"}

    # ---- end to end through the caller-side entry: a whole periodic structure with its neighbour
    # list in pinned host memory -> site energies, atomic forces and the virial back in pinned host memory.  The
    # environments are built and the forces assembled on the device, so the pair gradients never cross PCIe.
    if w.call == "EF" and w.nspecies == 0 and ncomp == 1 and not args.no_e2e:
        from ace_jl_b200.structure import B200Structure
        from ace_jl_b200.utils import fcc_structure, philox
        ncell = max(2, round((nenv / 4.0) ** (1.0 / 3.0)))
        from ace_jl_b200.structure import pack_neighbours
        sX, scell, sfirst, snbr, simg = fcc_structure(philox(w.seed + 99 + rank), ncell)
        snbr = pack_neighbours(snbr, simg)     # (j, S) in one 32-bit word per pair (ACEB200_NBR_PACKED): 4 B instead of 7 B over PCIe
        pX, pfirst, pnbr = (torch.from_numpy(a).pin_memory() for a in (sX, sfirst, snbr))
        st = B200Structure(pX.numpy(), pfirst.numpy(), pnbr.numpy(), None, scell, packed=True)
        sE = torch.empty((st.natoms, w.nprop, 1), dtype=torch.float64).pin_memory()
        sF = torch.empty((st.natoms, w.nprop, 3, 1), dtype=torch.float64).pin_memory()
        sW = torch.empty((w.nprop, 3, 3), dtype=torch.float64).pin_memory()
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
        fsum = float(np.abs(sF.numpy().sum(axis=0)).max() / np.abs(sF.numpy()).max())
        assert fsum < 1e-9, "forces of a periodic structure must sum to zero"
        e2e_struct = {"value": world * st.natoms * e2e_steps / float(dts.item()), "unit": UNIT,
                      "h2d_bytes_per_step": int(sX.nbytes + sfirst.nbytes + snbr.nbytes),
                      "d2h_bytes_per_step": int(sE.numel() * 8 + sF.numel() * 8 + sW.numel() * 8),
                      "steps": e2e_steps, "atoms_per_gpu": st.natoms, "pairs_per_gpu": st.npairs,
                      "call": "aceb200_structure_energy_forces",
                      "workload": "jittered periodic FCC crystal, 42 neighbours per atom inside rcut, positions + neighbour list "
                                  "(i, j, S; j and S packed in one 32-bit word per pair) in pinned host memory -> site energies, atomic forces, virial in pinned host memory "
                                  "(same model; environments built and forces assembled on the device)",
                      "timer": "host wall clock around the C-ABI calls (they return after the D2H copy)"}

    if rank == 0:
        work = algorithmic_work(basis, J, w.call, w.nprop)
        flops = work["flops"]
        per_launch_ms = {k: v / args.steps for k, v in stage.items() if v > 0.0}
        kernel_names = {"pool": "k_pool", "adjoint": "k_adjoint_stream", "forces": "k_forces", "product": "k_basis_stream", "coupling": "k_basis_stream"}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        step_ms = total_ms / args.steps
        hbm_ach = value / world * work["bytes"] / 1e9
        whole_tflops = work["flops_total"] * nenv / (step_ms * 1e-3) / 1e12
        stage_flops = dict(flops)
        if "coupling" in stage_flops:               # the fused value kernel does products + coupling in one launch
            stage_flops["basis"] = stage_flops.pop("product") + stage_flops.pop("coupling")
            kernel_names["basis"] = "k_basis_stream"
        if "jacobian" in stage_flops:               # the Jacobian stage: canonical dA, then W_e . dA per environment
            stage_flops["basis"] = stage_flops.pop("product_B") + stage_flops.pop("coupling_B")
            kernel_names["basis"] = "k_basis_stream"
            kernel_names["jacobian"] = "k_dA + k_dB_env"
        common = {k: v for k, v in per_launch_ms.items() if k in stage_flops}
        dom = max(common, key=common.get) if common else None
        fp64_roof = None
        if dom:
            ach = stage_flops[dom] * nenv / (common[dom] * 1e-3) / 1e12
            fp64_roof = {"bound": "fp64", "kernel": kernel_names.get(dom, dom), "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach / fp64_peak if fp64_peak else None}
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            tk = tr.get(f"config{w.key}", {}).get(kernel_names.get(dom, ""), None)
            if tk:
                traffic = (tk["dram_read_bytes"] + tk["dram_write_bytes"]) * nenv / tk["envs_per_launch"]
        except Exception:
            pass
        hbm_roof = {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                    "algorithmic_bytes_per_env": work["bytes"], "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"}
        # which roofline bounds the call: whichever fraction is larger (FP64 for the model calls, output
        # bandwidth for the vector- / matrix-valued basis)
        whole_frac = whole_tflops / fp64_peak if fp64_peak else 0.0
        if hbm_roof["frac"] > whole_frac:
            roofline = dict(hbm_roof)
            roofline["fp64"] = fp64_roof
        else:
            roofline = dict(fp64_roof or {"bound": "fp64", "achieved": whole_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": whole_frac})
            roofline["hbm"] = hbm_roof
        roofline.update({
            "traffic": traffic,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, profiles/r2_traffic.json",
            "peak_source": "FP64 FMA probe run in this process (MEASURED_PEAKS.json holds no FP64 figure)",
            "fp64_dmma_peak_tflops": dmma_peak,
            "algorithmic_flops_per_env": flops, "ms_per_launch": per_launch_ms,
            "whole_step_tflops": whole_tflops, "whole_step_frac": whole_frac})
        primary = e2e_struct or e2e_env
        line = {
            "metric": METRICS[w.call], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(w, nenv, world),
            "roofline": roofline, "parity": parity, "gpu_launches": int(launches), "clocks": clocks,
        }
        if primary:
            line["e2e"] = primary
        if e2e_struct and e2e_env:
            line["e2e_per_environment"] = e2e_env
        if not args.no_cpu and world == 1:          # the CPU baseline is a rank-0, N = 1 leg
            sample = args.cpu_envs or default_cpu_sample(w, basis)
            rate, cores, dtc, passes = cpu_reference_rate(w, basis, c, sample)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} environments x {J} neighbours, {passes} passes, {dtc:.1f} s, restated reference "
                                              "(C + OpenMP over environments, materialised dA like src/evaluator.jl:169)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
