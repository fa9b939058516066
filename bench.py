#!/usr/bin/env python
"""bench.py -- throughput of the element-block hot path on B200 (contract: see the task brief).

A *step* is one global Newton iteration over every owned element block of the workload:
the strain-recovery / material-update / internal-force pass (w3d_eps_sig_ifv) followed by the
element tangent-stiffness pass (w3d_tanstf).  Workload: BASELINE.json configs[3], the synthetic
structured 2M-element hex8 mesh with mm01 (geonl, B-bar), state at n prepared by one converged
load step so that about half of the Gauss points are plastic in the timed iteration.

  value   Gauss-point updates/s with u, du and all state resident in HBM (whole job, all ranks)
  e2e     same metric through the host-buffer C-ABI calls: H2D of u, du; D2H of ifv and of the
          stiffness ASSEMBLED on the device (w3d_assemble: k_diag, k_coeffs); e2e_element_blocks
          is the variant that returns the estiff(300,span) blocks instead.  At N > 1 every rank
          returns the assembled rows of its own domain (w3d_assemble_rows)
  residual  one w3d_residual call (upres + cvtest residual loop on the device), beside the step
  roofline, cpu_baseline, clocks: as the brief defines them.

`--impl reference` times the CPU restatement of the reference path (oracle/, OpenMP over blocks
like the reference) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dims, material kind, CPU sample dims)
    "hex8_2m_mm01": ((200, 100, 100), "mm01", (64, 50, 50)),   # BASELINE.json configs[3]
    "hex8_256k_mm01": ((64, 64, 64), "mm01", (64, 50, 50)),    # smaller variant for quick checks
    # SURVEY 8(d) variants of the headline workload: every point elastic; mixed hardening (beta 0.5)
    "hex8_2m_mm01_elastic": ((200, 100, 100), "mm01", (64, 50, 50)),
    "hex8_2m_mm01_beta05": ((200, 100, 100), "mm01", (64, 50, 50)),
    "hex8_1m_mm10": ((100, 100, 100), "mm10", (32, 25, 25)),   # BASELINE.json configs[4]: CP polycrystal
    "hex8_64k_mm10": ((40, 40, 40), "mm10", (32, 25, 25)),
    # the same with geometric hardening from the lattice curvature (k_0 = 10 as in verification/test74):
    # the k_mm10f instantiation that also keeps tau_l of the 12 systems per thread
    "hex8_64k_mm10_k0": ((40, 40, 40), "mm10", (32, 25, 25)),
    "hex8_1m_mm10_k0": ((100, 100, 100), "mm10", (32, 25, 25)),
    # mm05 cyclic plasticity, both options (the GP one with the reference's own deck values, ksi)
    "hex8_256k_mm05_fa": ((64, 64, 64), "mm05_fa", (32, 25, 25)),
    "hex8_256k_mm05_gp": ((64, 64, 64), "mm05_gp", (32, 25, 25)),
    # BASELINE.json configs[1]: the reference's fracture_models/seb deck with mm01 (17 318 hex8, the
    # deck's 137 user blocks and 2 domains, consistent q off); mesh from tests/golden/seb_deck.npz
    "seb_mm01": (None, "seb", None),
    # the same deck AS SHIPPED: material `mises e 200e9 nu 0.3 yld_pt 400e6 n_power 10` = mm03
    "seb_mm03": (None, "seb", None),
}
# mm10 workload: fcc (12 slip systems), Voce hardening with the property values of
# verification/test82/voche_model/warp3d_voche.inp:6-10, one crystal per element, Kocks angles
# from default_rng(20261017)
MM10_CRYSTAL = dict(e=150000.0, nu=0.28, harden_n=20.0, voche_m=5.0, theta_o=400.0, tau_y=300.0, tau_v=1200.0)


def make_workload(name, dims=None):
    """Returns (model, du_per_step, nprep): `nprep` converged load steps of increment du prepare
    the state at n; the timed Newton iteration is iteration 1 of step nprep+1."""
    from warp3d_b200 import (Material, attach_crystals, build_structured_hex8, graded_stretch,
                             make_crystal_type)
    wdims, kind, _ = WORKLOADS[name]
    dims = dims or wdims
    if kind == "seb":
        from warp3d_b200 import seb_bend_field, seb_model
        m, cons = seb_model(material="mises" if name == "seb_mm03" else "bilinear")
        m.deck_constraints = cons
        # one converged step of the bending field at 0.75 of the first-yield scale, then the timed
        # iteration adds the same again: ~90 % of the points are plastic in it
        return m, seb_bend_field(m, 0.6), 1
    if kind == "mm01":
        if name.endswith("_beta05"):
            m = build_structured_hex8(*dims, materials=[Material(beta=0.5)])
        else:
            m = build_structured_hex8(*dims)
        # elastic variant: 0.2 x the stretch keeps the largest strain (1.2e-3 after two steps)
        # under the yield strain 2e-3
        return m, (0.2 if name.endswith("_elastic") else 0.5) * graded_stretch(m), 1
    if kind.startswith("mm05"):
        from warp3d_b200 import cyclic_fa_material, cyclic_gp_material
        mat = cyclic_fa_material() if kind == "mm05_fa" else cyclic_gp_material()
        m = build_structured_hex8(*dims, materials=[mat])
        return m, 0.5 * graded_stretch(m), 1      # yield strain 2e-3 for both, as for mm01
    m = build_structured_hex8(*dims, materials=[Material(mat_type=10)], dt=1.0)
    attach_crystals(m, make_crystal_type(**MM10_CRYSTAL, **(dict(k_o=10.0) if name.endswith("_k0") else {})),
                    ncrystals=1)
    return m, 0.25 * graded_stretch(m, eps_lo=3.0e-3, eps_hi=9.0e-3), 2


def workload_description(name):
    kind = WORKLOADS[name][1]
    if kind == "seb" and name == "seb_mm03":
        # 15 history words instead of 11, and the three-launch structure: 2 x 128 B of scratch
        return dict(material="mm03 mises, the deck's own card (e 200e9 nu 0.3 yld_pt 400e6 n_power 10), "
                             "fracture_models/seb mesh and blocking, consistent q off",
                    bytes_stress=776.0 + 64.0 + 256.0, bytes_tanstf=708.0)
    if kind == "seb":
        return dict(material="mm01 bilinear (deck card switched from mises: e 200e9 nu 0.3 yld_pt 400e6 tan_e 2e9), "
                             "fracture_models/seb mesh and blocking, consistent q off",
                    bytes_stress=776.0, bytes_tanstf=708.0, flops_stress=3225.0, flops_tanstf=4540.0)
    if kind.startswith("mm05"):
        # as mm01 with 14 history words per point instead of 11 (3 x 8 B more each way)
        return dict(material={"mm05_fa": "mm05 cyclic, nonlinear_hardening (q_u 100e6, b_u 20, h_u 10e9, gamma_u 50)",
                              "mm05_gp": "mm05 cyclic, generalized_plasticity (verification/test39 values, ksi)"}[kind],
                    bytes_stress=776.0 + 48.0, bytes_tanstf=708.0)
    if kind == "mm01":
        # FP64 flops per Gauss point: 2 x DFMA + DMUL + DADD executed (predicated-on thread
        # instructions / Gauss points) from the per-instruction counts of the round-2 ncu captures
        # of the two kernels on the 50 %-plastic workload (profiles/r02_summary.md "instruction
        # mix": stress kernel 1 264 DFMA + 497 DMUL + 201 DADD, stiffness kernel 1 974 + 424 + 170)
        return dict(material="mm01 bilinear (E=200e9 as real*4, nu=0.3, yld=400e6, tan_e=E/100)"
                             + (", beta 0.5 (mixed hardening)" if name.endswith("_beta05") else "")
                             + (", stretch scaled to stay elastic" if name.endswith("_elastic") else ""),
                    bytes_stress=776.0, bytes_tanstf=708.0, flops_stress=3225.0, flops_tanstf=4540.0,
                    # measured DRAM bytes per Gauss point (dram__bytes_read + write of one `ncu --set
                    # full` launch on hex8_256k_mm01 / 2 097 152 points: round-2 captures,
                    # profiles/r02_stress_256k_details.csv, profiles/r02_tanstf_256k_v2_details.csv)
                    dram_stress=(531.654400e6 + 1011.120e6) / 2097152.0,
                    dram_tanstf=(674.907648e6 + 582.138112e6) / 2097152.0)
    # per Gauss point: history 2 x 158 x 8, urcs 2 x 72, eps 2 x 48, rot 72, scratch 2 x 128,
    # crystal properties 1448/8, element vectors 120  (DESIGN.md "mm10")
    return dict(material="mm10 crystal plasticity: fcc 12 slip, Voce (test82/voche_model values), 1 crystal/element",
                bytes_stress=2.0 * 158 * 8 + 144 + 96 + 72 + 256 + 181 + 120, bytes_tanstf=72 + 72 + 288 + 396.0)
METRIC = "gauss_point_updates_per_s_per_newton_iteration"
UNIT = "GP/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clocks and throttle reasons DURING the timed region, polled through NVML every few ms
    (same quantities as the profiling recipe's nvidia-smi line; nvidia-smi -lms is too coarse for
    a timed region of ~100 ms)."""

    def __init__(self, gpu_index, period_s=0.004):
        import threading
        self.samples, self.reasons, self.smax, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._ok = False
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(gpu_index)
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self._ok = True
        except Exception:
            return
        self.period = period_s
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=self.smax, reasons=[])
        if not self._ok:
            return out
        self._stop.set()
        self.t.join(timeout=2)
        if self.samples:
            out.update(sm_mhz=float(np.median(self.samples)), sm_min_mhz=float(min(self.samples)),
                       reasons=sorted(self.reasons), samples=len(self.samples),
                       power_w_max=float(max(self.power)) if self.power else None)
        return out


CPU_BUILD = "g++ -O3 -march=x86-64-v3 -ffp-contract=fast (oracle/Makefile: libw3d_oracle_fast.so)"


def prepare_cpu_case(workload, nthreads):
    """Same mesh family, material and load path as the GPU workload, on the oracle.  The timed
    CPU legs use the -O3 / FMA build of the restatement (the reference's own build is ifort -O3);
    the parity tests never do."""
    if not os.environ.get("W3D_ORACLE_STATS"):
        os.environ["W3D_ORACLE_FAST"] = "1"
    from oracle.oracle_py import Oracle
    from warp3d_b200 import new_state
    m, du, nprep = make_workload(workload, WORKLOADS[workload][2])   # seb: the whole deck
    orc = Oracle(m, nthreads=nthreads)
    st = new_state(m)
    u = np.zeros(m.nodof)
    for step in range(1, nprep + 1):
        orc.eps_sig_ifv(st, step, 0, u, du)
        orc.eps_sig_ifv(st, step, 1, u, du)
        orc.update(st); orc.adaptive_save(st)
        u = u + du
    orc.eps_sig_ifv(st, nprep + 1, 0, u, du)
    return m, orc, st, u, du, nprep + 1


def cpu_pass(m, orc, st, u, du, step):
    t0 = time.perf_counter()
    orc.eps_sig_ifv(st, step, 1, u, du)
    t1 = time.perf_counter()
    orc.tanstf(st, step, 1, u, du)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def cpu_baseline(workload, nrep=3):
    cores = os.cpu_count() or 1
    CPU_SAMPLE = WORKLOADS[workload][2]
    m, orc, st, u, du, step = prepare_cpu_case(workload, cores)
    cpu_pass(m, orc, st, u, du, step)
    ts = [cpu_pass(m, orc, st, u, du, step) for _ in range(nrep)]
    t_s = float(np.mean([a for a, _ in ts])); t_k = float(np.mean([b for _, b in ts]))
    ngp = m.noelem * 8
    what = (f"{CPU_SAMPLE[0]}x{CPU_SAMPLE[1]}x{CPU_SAMPLE[2]} sub-box of the same mesh, material and load path"
            if CPU_SAMPLE else "the whole deck")
    return dict(value=ngp / (t_s + t_k), unit=UNIT, cores=cores, kind="port", build=CPU_BUILD,
                sample=f"{m.noelem} elements ({what}), {nrep} passes, oracle/ C++ restatement with OpenMP over blocks",
                stress_gp_per_s=ngp / t_s, stiffness_elem_per_s=m.noelem / t_k)


def mm10_operation_counts(workload, dims=(16, 16, 16)):
    """Per-Gauss-point averages of the mm10 solver's operation counters (slip evaluations,
    Jacobians, dense solves ...) for this workload's material and load path, from the CPU
    restatement built with -DW3D_MM10_STATS, run in a child process on a small sub-box."""
    code = ("import json, sys; sys.path.insert(0, %r); import bench; from oracle import oracle_py\n"
            "bench.WORKLOADS[%r] = bench.WORKLOADS[%r][:2] + (%r,)\n"
            "m, orc, st, u, du, step = bench.prepare_cpu_case(%r, 4)\n"
            "oracle_py.mm10_stats(reset=True); orc.eps_sig_ifv(st, step, 1, u, du)\n"
            "s = oracle_py.mm10_stats(); print(json.dumps({k: v / (m.noelem * 8.0) for k, v in s.items()}))\n"
            % (ROOT, workload, workload, tuple(dims), workload))
    env = dict(os.environ, W3D_ORACLE_STATS="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True)
    return json.loads(out.stdout.strip().split("\n")[-1])


def mm10_flops_per_gp(ops, nslip=12, exponent=19):
    """Algorithmic FP64 flops of one mm10 Gauss-point update (FMA = 2), from the per-operation
    counts of the algorithm as implemented in w3d_mm10_core.h (DESIGN.md 3.3 lists them):
    one slip evaluation = residual R1 (+R2), one Jacobian = formJ, one solve = 7x7 LU + line-search
    bookkeeping, plus the once-per-crystal part (set-up, tangent, output terms, rotation)."""
    powering = exponent.bit_length() + bin(exponent).count("1")          # multiplies of x^e
    per_eval = 72 + nslip * (11 + 1 + powering + 2 + 2 + 12 + 6) + 66 + 15 + 24 + 6 + 66 + 24 + nslip * 5
    per_jac = nslip * 230 + 432 + 168 + 80
    per_solve = 203 + 84 + 14 + 98 + 30
    once = 300 + 600 + nslip * 150 + nslip * 110 + 200 + 170 + 200 + 150
    return (ops["slip_evals"] * per_eval + ops["jacobians"] * per_jac + ops["solves"] * per_solve
            + ops["crystal_solves"] * once)


def reference_arm(args):
    """The reference's own CPU algorithm for the path (oracle port; the Fortran cannot be built
    here: no Fortran compiler, needs ifort+MKL), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    CPU_SAMPLE = WORKLOADS[args.workload][2]
    m, orc, st, u, du, step = prepare_cpu_case(args.workload, cores)
    for _ in range(args.warmup):
        cpu_pass(m, orc, st, u, du, step)
    t0 = time.perf_counter()
    ts = [cpu_pass(m, orc, st, u, du, step) for _ in range(args.steps)]
    total = time.perf_counter() - t0
    ngp = m.noelem * 8
    value = ngp * args.steps / total
    t_s = float(np.mean([a for a, _ in ts])); t_k = float(np.mean([b for _, b in ts]))
    what = (f"{CPU_SAMPLE[0]}x{CPU_SAMPLE[1]}x{CPU_SAMPLE[2]} sub-box of the {args.workload} workload"
            if CPU_SAMPLE else f"the whole {args.workload} deck")
    sample = f"{m.noelem} elements per step ({what}), oracle/ port of the reference path, OpenMP over blocks"
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * total / args.steps, higher_is_better=True,
                scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=args.workload, cpu_sample_elements=m.noelem,
                            material=workload_description(args.workload)["material"], formulation="geonl+bbar"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", build=CPU_BUILD, sample=sample,
                                  stress_gp_per_s=ngp / t_s, stiffness_elem_per_s=m.noelem / t_k),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    emit(line)


def engine_arm(args):
    import torch
    import torch.distributed as dist
    from warp3d_b200.distributed import DomainEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own log (whatever NCCL_DEBUG the harness set,
        # e.g. INFO for its communicator check) goes to stderr unless the harness routed it itself
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    model, du_step, nprep = make_workload(args.workload)
    wd = workload_description(args.workload)
    TS = nprep + 1                      # the timed load step
    # seb: the deck's own domain column (2 domains) when run on 2 GPUs, slabs of blocks otherwise
    # exchange step: inside the engine (C ABI, NCCL send/recv between neighbouring slabs); at N > 1
    # rank 0 alone uploads u, du (NVLink broadcast) and every rank returns what it owns
    de = DomainEngine(model, rank, world, device=local,
                      assign_slabs=not (WORKLOADS[args.workload][1] == "seb" and world == 2),
                      exchange=args.exchange, bcast_inputs=world > 1)
    eng = de.engine
    nodof, nel_own = model.nodof, eng.n_owned_elems
    ngp_total = model.noelem * 8

    def pinned(n):
        return torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()

    u_h, du_h, ifv_h = pinned(nodof), pinned(nodof), pinned(nodof)
    est_h = pinned(nel_own * 300)

    # ---- state preparation: nprep converged load steps, roll over; then iteration 0 of step TS
    u_h[:] = 0.0; du_h[:] = du_step
    for step in range(1, nprep + 1):
        de.drive_eps_sig_internal_forces(step, 0, u_h, du_h, ifv_h)
        de.drive_eps_sig_internal_forces(step, 1, u_h, du_h, ifv_h)
        de.update(); de.adaptive_save()
        u_h[:] = u_h + du_step
    de.drive_eps_sig_internal_forces(TS, 0, u_h, du_h, ifv_h)
    _, sc = de.drive_eps_sig_internal_forces(TS, 1, u_h, du_h, ifv_h)
    if sc["material_cut_step"]:
        raise SystemExit("bench.py: the workload asked for a step cut; timing would be invalid")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    S = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))

    # ---- resident timing: K steps bracketed by barrier + synchronize, CUDA events on the stream
    for _ in range(max(args.warmup, 3)):
        de.eps_sig_ifv_resident(TS, 1, exchange=args.exchange); de.tanstf_resident(TS, 1)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    l0 = eng.launch_count
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    ev[0].record(S)
    for k in range(args.steps):
        de.eps_sig_ifv_resident(TS, 1, exchange=args.exchange)
        ev[2 * k + 1].record(S)
        de.tanstf_resident(TS, 1)
        ev[2 * k + 2].record(S)
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    t_stress = float(np.mean([ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)]))
    t_tan = float(np.mean([ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(args.steps)]))

    # ---- the other exchange mode as a labelled variant (N > 1): same K steps, same bracketing
    alt_ms, alt_mode = None, None
    if world > 1:
        alt_mode = "full" if args.exchange != "full" else "neighbours"
        eng.comm_set_options(alt_mode, True)
        for _ in range(3):
            de.eps_sig_ifv_resident(TS, 1); de.tanstf_resident(TS, 1)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(S)
        for k in range(args.steps):
            de.eps_sig_ifv_resident(TS, 1); de.tanstf_resident(TS, 1)
        b1.record(S)
        barrier()
        alt_ms = b0.elapsed_time(b1) / args.steps
        eng.comm_set_options(args.exchange, True)

    # ---- end-to-end timing through the host-buffer calls (pinned host memory)
    OWN = world > 1          # a rank copies back the dofs it owns, not the whole vector
    own_first, own_n = de.owned_range()
    for _ in range(2):
        de.drive_eps_sig_internal_forces(TS, 1, u_h, du_h, ifv_h, owned_only=OWN); de.tanstf(TS, 1, est_h)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nrep_e2e = max(2, min(args.steps, 5))
    e0.record(S)
    w0 = time.perf_counter()
    for _ in range(nrep_e2e):
        de.drive_eps_sig_internal_forces(TS, 1, u_h, du_h, ifv_h, owned_only=OWN)
        de.tanstf(TS, 1, est_h)
    e1.record(S)
    barrier()
    e2e_wall_ms = 1e3 * (time.perf_counter() - w0) / nrep_e2e
    e2e_ms = e0.elapsed_time(e1) / nrep_e2e

    # ---- the same iteration with the stiffnesses assembled on the device (assemble_code.f): what
    #      crosses to the host is the sparse upper triangle the solver consumes, not the element
    #      blocks.  x = xmin face clamped.  With several domains every rank returns the rows of the
    #      nodes its own elements touch (interface rows come from two ranks and add on the host).
    cst = np.zeros(nodof, dtype=np.int32)
    if getattr(model, "deck_constraints", None):
        cst[model.deck_constraints["dofs"]] = 1        # the deck's own constraints
    else:
        xs = model.c.reshape(-1, 3)[:, 0]
        cst.reshape(-1, 3)[xs <= xs.min() + 1e-9] = 1
    ts0 = time.perf_counter()
    st_ = eng.assemble_setup(cst)
    setup_s = time.perf_counter() - ts0
    kd_h, kc_h = pinned(st_["n_eqns"]), pinned(max(st_["ncoeff"], 1))
    for _ in range(2):
        de.drive_eps_sig_internal_forces(TS, 1, u_h, du_h, ifv_h, owned_only=OWN)
        eng.tanstf_resident(TS, 1); eng.assemble(kd_h, kc_h)
    barrier()
    a0, a1, a2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a0.record(S)
    for _ in range(nrep_e2e):
        de.drive_eps_sig_internal_forces(TS, 1, u_h, du_h, ifv_h, owned_only=OWN)
        eng.tanstf_resident(TS, 1); eng.assemble(kd_h, kc_h)
    a1.record(S)
    for _ in range(nrep_e2e):
        eng.assemble(to_host=False)
    a2.record(S)
    barrier()
    asm = dict(ms=a0.elapsed_time(a1) / nrep_e2e, kernel_ms=a1.elapsed_time(a2) / nrep_e2e,
               neqns=st_["neqns"], ncoeff=st_["ncoeff"], setup_s=setup_s,
               d2h=8 * (own_n if OWN else nodof) + 8 * (st_["n_eqns"] + st_["ncoeff"]))

    # structure residual + convergence norms on the device (upres.f, cvtest.f:62-78): the internal
    # forces stay in HBM, 7 scalars come back.  Not part of the timed step; reported beside it.
    eng.residual_set_loads(cst, pbar=np.ones(nodof), load=np.ones(nodof))
    for _ in range(3):
        eng.residual(1, to_host=False)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(S)
    for _ in range(20):
        eng.residual(1, to_host=False)
    r1.record(S)
    barrier()
    asm["residual_ms"] = r0.elapsed_time(r1) / 20

    t = torch.tensor([total_ms, t_stress, t_tan, e2e_ms, e2e_wall_ms, float(launches), asm["ms"], asm["kernel_ms"],
                      float(asm["d2h"]), float(asm["ncoeff"]), alt_ms or 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        tl = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches = int(tl[5].item())
        asm["d2h"], asm["ncoeff"] = int(tl[8].item()), int(tl[9].item())     # bytes / terms over all ranks
    total_ms, t_stress, t_tan, e2e_ms, e2e_wall_ms = [float(x) for x in t[:5].tolist()]
    asm["ms"], asm["kernel_ms"] = float(t[6].item()), float(t[7].item())     # max over ranks
    alt_ms = float(t[10].item()) or None

    if rank == 0:
        peak, peak_kind = measured_peaks()
        ms_per_step = total_ms / args.steps
        value = ngp_total / (ms_per_step * 1e-3)
        # dominant kernel = the one with the longer average launch; per-launch units = owned GPs
        ngp_own = nel_own * 8
        sname = ("k_eps_sig_ifv<phase 1,2> + k_mm10f" if WORKLOADS[args.workload][1] == "mm10" else
                 "k_eps_sig_ifv<phase 1,2> + k_mm03" if args.workload == "seb_mm03" else "k_eps_sig_ifv")
        k_name, k_ms, k_bytes = (("k_tanstf", t_tan, wd["bytes_tanstf"]) if t_tan >= t_stress
                                 else (sname, t_stress, wd["bytes_stress"]))
        ach = ngp_own * k_bytes / (k_ms * 1e-3) / 1e9
        other = (sname, t_stress, wd["bytes_stress"]) if k_name == "k_tanstf" else ("k_tanstf", t_tan, wd["bytes_tanstf"])
        fp64_peak = None
        pk = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
        if os.path.exists(pk):
            with open(pk) as f:
                fp64_peak = json.load(f).get("dfma_tflops")
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
            data="synthetic",
            config=dict(workload=args.workload, elements=model.noelem, gauss_points=ngp_total, blocks=model.nelblk,
                        nodof=nodof, material=wd["material"],
                        formulation="hex8 l3disop 2x2x2, geonl, bbar",
                        l2=("inputs_exceed_l2 (state >> 126 MB)" if model.noelem > 100000 else
                            "state fits L2 (small deck): launch-latency bound, not a roofline case"),
                        partition=(f"{world} contiguous block slabs; exchange inside the engine (C ABI): "
                                   + ("NCCL send/recv of the shared dofs between neighbouring domains"
                                      if args.exchange != "full" else "ncclAllReduce of the whole ifv vector")
                                   + " + one 5-double all-reduce of the scalars / cut flag; e2e: rank 0 uploads u, du "
                                     "(ncclBroadcast), every rank returns the dofs and rows it owns")
                        if world > 1 else "single GPU",
                        step="w3d_eps_sig_ifv + w3d_tanstf (one Newton iteration)"),
            stress_pass=dict(ms=t_stress, gp_updates_per_s=ngp_total / (t_stress * 1e-3)),
            stiffness_pass=dict(ms=t_tan, element_stiffnesses_per_s=model.noelem / (t_tan * 1e-3)),
            roofline=dict(bound="hbm", kernel=k_name, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak,
                          peak_kind=f"{peak_kind} copy bandwidth (MEASURED_PEAKS.json)",
                          traffic=(ngp_own * wd["dram_tanstf" if k_name == "k_tanstf" else "dram_stress"]
                                   if "dram_stress" in wd else None),
                          traffic_note="bytes per launch = DRAM bytes per point of one ncu --set full launch "
                                       "(256k-element mesh) x points of this launch",
                          bytes_per_gp=k_bytes, launch_ms=k_ms,
                          other_kernel=dict(kernel=other[0], launch_ms=other[1],
                                            achieved=ngp_own * other[2] / (other[1] * 1e-3) / 1e9,
                                            frac=ngp_own * other[2] / (other[1] * 1e-3) / 1e9 / peak),
                          fp64_peak_tflops=fp64_peak),
            e2e=dict(value=ngp_total / (e2e_ms * 1e-3), unit=UNIT, ms_per_step=e2e_ms, wall_ms_per_step=e2e_wall_ms,
                     h2d_bytes_per_step=2 * 8 * nodof, d2h_bytes_per_step=8 * nodof + 8 * 300 * model.noelem,
                     note="pinned host buffers; u,du in; ifv and estiff(300,span) blocks out"),
            gpu_launches=launches, clocks=clocks,
            exchange_variant=(dict(mode=alt_mode, ms_per_step=alt_ms, value=ngp_total / (alt_ms * 1e-3), unit=UNIT,
                                   note="same K steps with the other exchange mode (full = ncclAllReduce of the whole "
                                        "vector, the reference's wmpi_reduce_vec result on every rank)")
                              if alt_ms else None),
            plastic_check=dict(plastic_work=sc["plastic_work"], internal_energy=sc["internal_energy"]),
        )
        if asm:
            # headline end-to-end number = the iteration as a solver-facing caller runs it: ifv and
            # the ASSEMBLED sparse stiffness come back; the element-block variant stays alongside
            line["e2e_element_blocks"] = line["e2e"]
            line["e2e"] = dict(
                value=ngp_total / (asm["ms"] * 1e-3), unit=UNIT, ms_per_step=asm["ms"],
                h2d_bytes_per_step=2 * 8 * nodof, d2h_bytes_per_step=asm["d2h"],
                assemble_kernel_ms=asm["kernel_ms"], neqns=asm["neqns"], ncoeff=asm["ncoeff"],
                structure_setup_s=asm["setup_s"],
                note="pinned host buffers; u,du in; ifv, k_diag(neqns) and k_coeffs(ncoeff) out: w3d_eps_sig_ifv + "
                     "w3d_tanstf(estiff=NULL) + w3d_assemble (assemble_code.f on the device, x=xmin face clamped); "
                     "sparsity set-up once per constraint set, outside the timed region"
                     + ("; every rank returns the rows of its own nodes (interface rows from two ranks add on "
                        "the host), ncoeff = terms over all ranks" if world > 1 else ""))
            rb = 44 * nodof      # cstmap 4 + pbar, load, du, ifv in + res out, 8 each
            line["residual"] = dict(
                call_ms=asm["residual_ms"], bytes=rb, achieved=rb / (asm["residual_ms"] * 1e-3) / 1e9, unit="GB/s",
                frac=rb / (asm["residual_ms"] * 1e-3) / 1e9 / peak,
                note="w3d_residual (upres + cvtest residual loop), residual left on the device; the call "
                     "includes the read-back of the 7 scalars (one stream synchronisation)")
        if "flops_stress" in wd and fp64_peak:
            # both mm01 kernels sit near the FP64 ridge (4.2 and 6.4 flop/B against 36.4 TFLOP/s /
            # 6.55 TB/s = 5.6 flop/B), and a DMUL or DADD occupies the pipe like a DFMA: with 1 962
            # and 2 567 FP64 instructions per point the pipe floors are 1.7 and 2.3 ms for 16 M
            # points, beside HBM floors of 1.9 and 1.7 ms.  State them against the DFMA peak as well.
            fs = ngp_own * wd["flops_stress"] / (t_stress * 1e-3) / 1e12
            ft = ngp_own * wd["flops_tanstf"] / (t_tan * 1e-3) / 1e12
            line["roofline"]["fp64"] = dict(
                bound="fp64", unit="TFLOP/s", peak=fp64_peak,
                peak_kind="measured DFMA throughput (tools/fp64_peak.cu, profiles/r01_fp64_peak.json)",
                k_eps_sig_ifv=dict(achieved=fs, frac=fs / fp64_peak, flops_per_gp=wd["flops_stress"], launch_ms=t_stress),
                k_tanstf=dict(achieved=ft, frac=ft / fp64_peak, flops_per_gp=wd["flops_tanstf"], launch_ms=t_tan))
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.workload)
            if WORKLOADS[args.workload][1] == "mm10" and fp64_peak:
                # mm10 is bound by the FP64 pipe, not HBM: state the algorithmic flops too
                ops = mm10_operation_counts(args.workload)
                fl = mm10_flops_per_gp(ops)
                tf = ngp_own * fl / (t_stress * 1e-3) / 1e12
                line["roofline"]["fp64"] = dict(
                    bound="fp64", achieved=tf, peak=fp64_peak, unit="TFLOP/s", frac=tf / fp64_peak,
                    flops_per_gp=fl, operations_per_gp=ops, launch_ms=t_stress,
                    peak_kind="measured DFMA throughput (tools/fp64_peak.cu, profiles/r01_fp64_peak.json)",
                    note="algorithmic flops = per-operation counts of the algorithm as implemented "
                         "(DESIGN.md 3.3) x operation counters of the CPU restatement on a sub-box")
    de.close()
    # ---- BASELINE.json configs[4] (1 M-element mm10 polycrystal) on the same ranks: a sub-object of
    #      the same line, so the 1/2/4/8-GPU runs of the default command carry its scaling curve too
    second = None
    if args.workload == "hex8_2m_mm01" and not args.no_mm10:
        try:
            second = config5_leg(args, rank, world, local)
        except Exception as ex:          # never lose the headline line to the secondary leg
            second = dict(workload="hex8_1m_mm10", error=f"{type(ex).__name__}: {ex}")
    if rank == 0:
        if second is not None:
            line["mm10"] = second
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def config5_leg(args, rank, world, local):
    """hex8_1m_mm10 (BASELINE.json configs[4]) with the blocks partitioned over the same ranks:
    K timed Newton iterations (stress pass = kinematics + k_mm10f + internal force, then tangent),
    bracketed like the headline workload; max over ranks."""
    import torch
    import torch.distributed as dist
    from warp3d_b200.distributed import DomainEngine
    name = "hex8_1m_mm10"
    model, du_step, nprep = make_workload(name)
    wd = workload_description(name)
    TS = nprep + 1
    de = DomainEngine(model, rank, world, device=local, assign_slabs=True, exchange=args.exchange,
                      bcast_inputs=world > 1)
    eng = de.engine
    nodof = model.nodof
    u_h = torch.zeros(nodof, dtype=torch.float64, pin_memory=True).numpy()
    du_h = torch.empty(nodof, dtype=torch.float64, pin_memory=True).numpy()
    ifv_h = torch.empty(nodof, dtype=torch.float64, pin_memory=True).numpy()
    du_h[:] = du_step
    for step in range(1, nprep + 1):
        de.drive_eps_sig_internal_forces(step, 0, u_h, du_h, ifv_h)
        de.drive_eps_sig_internal_forces(step, 1, u_h, du_h, ifv_h)
        de.update(); de.adaptive_save()
        u_h[:] = u_h + du_step
    de.drive_eps_sig_internal_forces(TS, 0, u_h, du_h, ifv_h)
    _, sc = de.drive_eps_sig_internal_forces(TS, 1, u_h, du_h, ifv_h)
    if sc["material_cut_step"]:
        de.close()
        return dict(workload=name, error="the workload asked for a step cut")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    S = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))
    K = max(3, min(args.steps, 10))
    for _ in range(3):
        de.eps_sig_ifv_resident(TS, 1, exchange=args.exchange); de.tanstf_resident(TS, 1)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K + 1)]
    l0 = eng.launch_count
    ev[0].record(S)
    for k in range(K):
        de.eps_sig_ifv_resident(TS, 1, exchange=args.exchange)
        ev[2 * k + 1].record(S)
        de.tanstf_resident(TS, 1)
        ev[2 * k + 2].record(S)
    barrier()
    launches = eng.launch_count - l0
    t = torch.tensor([ev[0].elapsed_time(ev[-1]) / K,
                      float(np.mean([ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(K)])),
                      float(np.mean([ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(K)]))],
                     dtype=torch.float64, device="cuda")
    nl = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    ms, t_stress, t_tan = [float(x) for x in t.tolist()]
    nel_own = eng.n_owned_elems
    de.close()
    ngp = model.noelem * 8
    out = dict(workload=name, config="BASELINE.json configs[4]", elements=model.noelem, gauss_points=ngp,
               material=wd["material"], n_gpus=world, steps=K, warmup=3, ms_per_step=ms,
               value=ngp / (ms * 1e-3), unit=UNIT, scaling="strong",
               stress_pass=dict(ms=t_stress, gp_updates_per_s=ngp / (t_stress * 1e-3),
                                kernels="k_eps_sig_ifv<10,phase 1> + k_mm10f + k_eps_sig_ifv<10,phase 2>"),
               stiffness_pass=dict(ms=t_tan), gpu_launches=int(nl.item()),
               plastic_check=dict(plastic_work=sc["plastic_work"], internal_energy=sc["internal_energy"]))
    pk = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
    if rank == 0 and world == 1 and not args.no_cpu and os.path.exists(pk):
        with open(pk) as f:
            fp64_peak = json.load(f).get("dfma_tflops")
        ops = mm10_operation_counts(name)
        fl = mm10_flops_per_gp(ops)
        tf = nel_own * 8 * fl / (t_stress * 1e-3) / 1e12
        out["roofline"] = dict(bound="fp64", achieved=tf, peak=fp64_peak, unit="TFLOP/s", frac=tf / fp64_peak,
                               flops_per_gp=fl, launch_ms=t_stress,
                               peak_kind="measured DFMA throughput of tools/fp64_peak.cu (own microbenchmark, "
                                         "profiles/r01_fp64_peak.json); MEASURED_PEAKS.json has no FP64 figure",
                               note="algorithmic flops = per-operation counts of the algorithm as implemented x "
                                    "operation counters of the CPU restatement on a sub-box")
    return out


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries below us write to file descriptor 1 on their
    own (NCCL prints its version banner there whenever the harness sets NCCL_DEBUG, whatever
    NCCL_DEBUG_FILE says): keep a private duplicate of the real stdout for the result line and
    point descriptor 1 at stderr for everybody else, child processes included."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="hex8_2m_mm01", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-mm10", action="store_true",
                    help="skip the hex8_1m_mm10 sub-object (BASELINE.json configs[4]) of the default workload")
    ap.add_argument("--exchange", default="neighbours", choices=["neighbours", "full"],
                    help="N>1: send/recv of the shared dofs between neighbouring domains, or all-reduce of the whole ifv "
                         "vector; the other mode is timed as well and reported as exchange_variant")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        reference_arm(args)
    else:
        engine_arm(args)


if __name__ == "__main__":
    main()
