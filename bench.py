#!/usr/bin/env python
"""bench.py -- CG GFLOP/s and HBM roofline of the B200-native HPCCG hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload weak512|c2|c3|strong] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A STEP is one complete HPCCG() solve (HPCCG.cpp:312-402): max_iter = 150 -> 149 CG iterations of
{p = r + beta p, [halo], Ap = A p, p.Ap, x += alpha p, r -= alpha Ap, r.r} plus the reference's set-up
(p = x, Ap = A p, r = b - Ap, r.r) on the synthetic matrix generate_matrix defines (generate_matrix.cpp:196-307).

Workloads (BASELINE.json configs):
  weak512 (default) configs[3]: 27-pt, local 512x512x512 per GPU, ranks stacked in z, weak scaling. At N = 1 this is
                    the configuration the north_star target is quoted on ("CG loop >= 80 % of HBM roofline on
                    1 B200 at 512^3 per GPU").
  c2                configs[1]: 27-pt, local 256^3 (also reported inside the default line under "also")
  c3                configs[2]: 7-pt, local 512^3
  strong            configs[4]: 27-pt, global 512x512x1024 split in z over N GPUs (strong scaling)

Printed (rank 0, ONE JSON line): value = whole-job CG GFLOP/s in HPCCG's own accounting (main.cpp:217-227:
niters * (4 + 6 + 2*27) * total_nrow flops) with b, x and the matrix resident in HBM when the timed region starts;
e2e = the same metric through the reference-named call HPCCG(A, b, x, ...) with HOST b and x (pinned), host->device
and device->host copies inside the timed region; roofline = the dominant kernel (fused SpMV + p.Ap) against the
measured HBM copy peak; cpu_baseline = the reference's own OpenMP build (oracle/_ref, compiled from the unmodified
reference sources) on this box's host cores on a bounded sample.

Also in the default line (each leg short; the whole run stays within a few minutes per N):
  N = 1 : also (configs[1] 256^3 on the GPU AND on the reference's OpenMP build: one ratio on an identical problem),
          also_c3 (configs[2] 7-pt 512^3), also_c1 (configs[0] 20x30x10, ms per solve), also_strong (configs[4] on one GPU =
          the same-lease base of the strong-scaling curve), also_pattern_format (opt-in pattern-coded mirror)
  N > 1 : parity (global 256^3 split in z over the N ranks through BOTH data planes -- peer memory inside the kernels and
          NCCL between them -- residual history against tests/golden/golden_256.json, the real reference's serial run;
          the run FAILS, rc 3, above 1e-8) and also_strong (configs[4] split over the N ranks)
  config.comm says which data plane carried the halos and scalar sums of the timed solves.

--impl reference times ONLY that CPU reference (rank 0 alone under torchrun) and prints the same line shape; its
`config` repeats the GPU arm's (the contract), `sample_config` + `same_config: false` say what the bounded sample really
was, and `also` holds the reference on the full configs[1] problem (256^3, 149 iterations) for the same-config ratio.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "cg_gflops"
UNIT = "GFLOP/s"
FLOPS_PER_ROW_ITER = 64.0  # 4 (2 ddot) + 6 (3 waxpby) + 2*27 (sparsemv), main.cpp:217-227
MAX_ITER = 150             # -> 149 iterations (HPCCG.cpp:358)

WORKLOADS = {
    # name: (nx, ny, nz_local or None, stencil, scaling, global_nz for strong)
    "weak512": dict(nx=512, ny=512, nz=512, stencil=27, scaling="weak"),
    "c1": dict(nx=20, ny=30, nz=10, stencil=27, scaling="weak"),
    "c2": dict(nx=256, ny=256, nz=256, stencil=27, scaling="weak"),
    "c3": dict(nx=512, ny=512, nz=512, stencil=7, scaling="weak"),
    "strong": dict(nx=512, ny=512, nz_global=1024, stencil=27, scaling="strong"),
}


def bytes_per_row_iter(stencil: int, fmt: str = "sell", eager_x: bool = False) -> dict:
    """Algorithmic HBM bytes per local row per CG iteration (SURVEY.md 8d / DESIGN.md): matrix streamed once
    (8 B value + 4 B column id per slot; one 2-byte pattern id per row in the opt-in pattern format), p gathered once, Ap written;
    x,p,r,Ap read + x,r written; r,p read + p written."""
    spmv = (2 if fmt == "pattern" else stencil * 12) + 16
    if eager_x:  # x += alpha p in the kernel after the SpMV: read x,p,r,Ap + write x,r ; then read r,p + write p
        return {"spmv_dot": spmv, "update_xr_dot": 48, "p_update": 24, "iteration": spmv + 72}
    # default: the x update rides in the next p-update: read r,Ap + write r ; then read x,r,p + write x,p
    return {"spmv_dot": spmv, "update_xr_dot": 24, "p_update": 40, "iteration": spmv + 64}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="weak512", choices=sorted(WORKLOADS))
    ap.add_argument("--nx", type=int)
    ap.add_argument("--ny", type=int)
    ap.add_argument("--nz", type=int, help="local nz per GPU (overrides the workload)")
    ap.add_argument("--max-iter", type=int, default=MAX_ITER)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary 256^3 (configs[1]) measurement")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="multi-GPU: halo exchange not overlapped (A/B)")
    ap.add_argument("--unfused", action="store_true", help="literal reference kernel sequence (A/B)")
    ap.add_argument("--eager-x", action="store_true", help="x += alpha p right after the SpMV instead of deferred into the p-update (A/B)")
    ap.add_argument("--format", default=os.environ.get("HPCCG_BENCH_FORMAT", "sell"), choices=["sell", "pattern"],
                    help="device-mirror format: sell = SELL-128 values + int32 columns (north-star layout, default); "
                         "pattern = lossless 16-bit row-pattern ids (SURVEY.md 8 f3)")
    ap.add_argument("--cpu-iters", type=int, default=30, help="CG iterations of the CPU reference sample per step")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the 256^3 parity block")
    ap.add_argument("--no-strong", action="store_true", help="skip the also_strong leg (configs[4])")
    ap.add_argument("--no-ref-c2", action="store_true", help="skip the reference's full 256^3 run (same-config ratio)")
    ap.add_argument("--side-steps", type=int, default=4, help="timed steps of the secondary legs (also_*)")
    return ap.parse_args()


def resolve_workload(args, size: int) -> dict:
    w = dict(WORKLOADS[args.workload])
    w["name"] = args.workload
    if "nz_global" in w:
        if w["nz_global"] % size:
            raise SystemExit(f"strong workload: {w['nz_global']} planes do not split over {size} GPUs")
        w["nz"] = w["nz_global"] // size
    if args.nx:
        w["nx"] = args.nx
    if args.ny:
        w["ny"] = args.ny
    if args.nz:
        w["nz"] = args.nz
    if args.nx or args.ny or args.nz:
        w["name"] += "-custom"
    return w


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured: MEASURED_PEAKS.json hbm_gbs (bf16 copy, read+write)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback: B200_PROFILING.md 6.65 TB/s (MEASURED_PEAKS.json absent)"


def ncu_traffic(workload: str):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full summary, if one matches."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(workload)
        except Exception:  # noqa: BLE001
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


# =====================================================================================================================
# CPU reference leg (oracle/_ref = the reference's own sources; the one place bench.py may execute oracle/)
# =====================================================================================================================
def cpu_sample_dims(w: dict) -> tuple[int, int, int]:
    """A z-slab of the workload the unmodified reference can hold: it overflows `int local_nnz = 27*local_nrow`
    beyond 430^3 (generate_matrix.cpp:223) and needs 720 B/row of host memory (README.md:92-105)."""
    nz = max(1, min(w["nz"], (16 * 1024 * 1024) // (w["nx"] * w["ny"])))
    return w["nx"], w["ny"], nz


def _ref_variant():
    sys.path.insert(0, str(ROOT / "oracle"))
    import refwrap  # the checker / baseline; never on the product path
    if refwrap.available("omp"):
        return refwrap, "omp", "reference"
    if refwrap.available("serial"):
        return refwrap, "serial", "reference"
    return refwrap, "oracle", "port"


def _cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_cpu_dims(nx, ny, nz, stencil, iters, steps, warmup, variant=None):
    """Times the reference's HPCCG() on an nx*ny*nz block. Returns (gflops list, seconds list, cores, variant, gen_s, last)."""
    refwrap, best, kind = _ref_variant()
    variant = variant or best
    n = nx * ny * nz
    t0 = time.time()
    R = refwrap.RefWorld(nx, ny, nz, size=1, stencil=stencil, variant=variant)
    t_gen = time.time() - t0
    cores = R.threads() if variant == "omp" else 1
    rates, secs, last = [], [], None
    for i in range(warmup + steps):
        last = R.solve(iters + 1, 0.0, hist=False, want_x=False)
        if i >= warmup:
            rates.append(last["niters"] * FLOPS_PER_ROW_ITER * n / last["times"][0] / 1e9)
            secs.append(float(last["times"][0]))
    R.close()
    return rates, secs, cores, variant, t_gen, last


def run_cpu_reference(w: dict, iters: int, steps: int, warmup: int):
    """Times the reference's HPCCG() (OpenMP build, all host threads) on the sample. Returns (gflops list, info)."""
    _, _, kind = _ref_variant()
    nx, ny, nz = cpu_sample_dims(w)
    rates, secs, cores, variant, t_gen, _ = run_cpu_dims(nx, ny, nz, w["stencil"], iters, steps, warmup)
    info = {"kind": kind, "cores": cores, "variant": variant, "nproc": os.cpu_count(), "cpu_model": _cpu_model(),
            "flags": "-O3 -funroll-all-loops -malign-double -fopenmp -DUSING_OMP -DWALL (MakefileOMP:83,103,117,133)",
            "sample": f"{nx}x{ny}x{nz} z-slab of the workload ({w['stencil']}-pt), {iters} CG iterations per step, "
                      f"{steps} steps after {warmup} warm-up; reference HPCCG() times[0]; matrix generation {t_gen:.1f} s not timed",
            "sample_config": {"nx": nx, "ny": ny, "nz": nz, "stencil": w["stencil"], "cg_iterations": iters,
                              "rows": nx * ny * nz},
            "seconds_per_step": secs}
    return rates, info


def run_cpu_c2(full_iters: int = MAX_ITER - 1, one_thread_iters: int = 8):
    """The reference on configs[1] exactly as the GPU arm runs it (27-pt 256^3, 149 iterations, all host threads), plus the
    reference's serial build (1 thread) on the same matrix for a few iterations."""
    c2 = WORKLOADS["c2"]
    out = {"workload": "configs[1]: 27-pt 256x256x256, 150 CG iterations"}
    rates, secs, cores, variant, t_gen, last = run_cpu_dims(c2["nx"], c2["ny"], c2["nz"], 27, full_iters, 1, 0)
    out.update({"value": rates[0], "unit": UNIT, "cores": cores, "variant": variant, "seconds_per_step": secs[0],
                "niters": last["niters"], "normr": last["normr"], "generation_s": round(t_gen, 1),
                "gbs_at_412_bytes_per_row": 412.0 * c2["nx"] * c2["ny"] * c2["nz"] * last["niters"] / secs[0] / 1e9})
    refwrap, _, _ = _ref_variant()
    if refwrap.available("serial") and one_thread_iters > 0:
        r1, s1, _, _, _, l1 = run_cpu_dims(c2["nx"], c2["ny"], c2["nz"], 27, one_thread_iters, 1, 0, variant="serial")
        out["one_thread"] = {"value": r1[0], "unit": UNIT, "cores": 1, "variant": "serial", "cg_iterations": l1["niters"],
                             "seconds": s1[0]}
    return out


def run_cpu_c1():
    """configs[0]: test_HPCCG 20 30 10 serial, 150 iterations -- the reference's own CPU-runnable case."""
    refwrap, best, _ = _ref_variant()
    out = {}
    for variant in ("serial", "omp"):
        if not refwrap.available(variant):
            continue
        rates, secs, cores, _, _, last = run_cpu_dims(20, 30, 10, 27, MAX_ITER - 1, 5, 2, variant=variant)
        out[variant] = {"ms_per_solve": 1e3 * statistics.median(secs), "value": statistics.median(rates), "unit": UNIT,
                        "cores": cores, "niters": last["niters"]}
    return out


def reference_arm(args, w: dict, config: dict, rank: int, size: int):
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers unless the user set it; the reference arm is meant to use all
    # the host threads it can (libgomp reads the variable when the reference library is loaded, i.e. later)
    if "TORCHELASTIC_RUN_ID" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    # a bounded sample: the driver passes the GPU arm's --steps/--warmup; keep the run within a few minutes
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    rates, info = run_cpu_reference(w, args.cpu_iters, steps, warmup)
    value = statistics.mean(rates)
    ms = 1e3 * statistics.mean(info["seconds_per_step"])
    info_out = dict(info)
    info_out["value"] = value
    info_out["unit"] = UNIT
    sc = info["sample_config"]
    same = (sc["nx"], sc["ny"], sc["nz"], sc["cg_iterations"]) == (w["nx"], w["ny"], w["nz"] * size, args.max_iter - 1)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "sample_config": sc, "same_config": same,
            "note": ("`config` names the GPU arm's workload (the contract); what was TIMED here is `sample_config`: one host, "
                     f"{info['cores']} threads, a z-slab the unmodified reference can hold (its `int local_nnz = 27*local_nrow` overflows "
                     "beyond 430^3, generate_matrix.cpp:223) and a bounded iteration count -- a RATE on a smaller problem of the "
                     "same stencil, not a time on the same problem" +
                     (f"; at N = {args.gpus} the GPU arm's value is the aggregate of {args.gpus} GPUs against this ONE host" if args.gpus > 1 else "") +
                     "; `also` is the reference on the identical configs[1] problem"),
            "cpu_baseline": info_out,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    if not args.no_ref_c2 and not (args.nx or args.ny or args.nz):
        try:
            line["also"] = run_cpu_c2()
            line["also_c1"] = run_cpu_c1()
        except Exception as e:  # noqa: BLE001
            line["also"] = {"failed": str(e)[:200]}
    print(json.dumps(line), flush=True)


# =====================================================================================================================
# GPU arm
# =====================================================================================================================
def gpu_arm(args, w: dict, config: dict, rank: int, size: int, local_rank: int):
    import numpy as np
    import torch
    import torch.distributed as dist
    import hpccg_pkg
    H = hpccg_pkg.load()  # ImportError if libhpccg_b200.so is missing: there is no fallback

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    H.set_print(False)
    gloo = None
    if size > 1:
        from hpccg_sycl_b200 import dist as hdist
        _, _, gloo = hdist.init_process_group_context(use_nccl=True)
    else:
        H.set_rank(0, 1)

    def barrier():
        if size > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if size == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def comm_name(m) -> str:
        if size == 1:
            return "none (single rank)"
        c = m.comm()
        if c["peer"]:
            return ("peer memory inside the kernels (NVLink P2P stores + in-kernel scalar all-reduce)" +
                    (", halo put fused into the p-update kernel" if c["fused_put"] else ", stand-alone halo put kernel"))
        return "NCCL between the kernels (send/recv halos overlapped with the interior SpMV, 1-double all-gathers)"

    def measure(nx, ny, nz, stencil, steps, warmup, do_e2e, max_iter, fmt=None, pageable=False):
        """One workload: device-resident value, per-kernel roofline numbers, e2e through HPCCG()."""
        fmt = fmt or args.format
        n = nx * ny * nz
        # host row arrays exist only where the reference itself could hold them (27 n < 2^31 and a sane footprint)
        host_rows = (27 * n < 2 ** 31) and n <= 32 * 1024 * 1024 and size == 1
        H.set_options(stencil, host_rows)
        H.set_matrix_format(fmt)
        t0 = time.time()
        A = H.generate_matrix(nx, ny, nz)
        if size > 1:
            H.make_local_matrix(A)
        m = A.device()  # the ELL mirror in HBM, built outside every timed region
        t_setup = time.time() - t0
        info = m.info()
        b = torch.from_numpy(A.b).to(dev)
        x = torch.zeros(n, dtype=torch.float64, device=dev)
        flags = H.SOLVE_TIMERS if hasattr(H, "SOLVE_TIMERS") else 4
        if args.no_overlap:
            flags |= 2
        if args.unfused:
            flags |= 1
        if args.eager_x:
            flags |= 32
        bpr = bytes_per_row_iter(stencil, fmt, args.eager_x or args.unfused)

        def step(acc=None):
            x.zero_()
            out = H.dev.cg_solve(m, b, x, max_iter, 0.0, flags=flags, want_hist=False, want_times=True)
            if acc is not None:
                acc["loop_ms"] += out["loop_ms"]
                acc["iters"] += out["niters"]
                for i, k in ((7, "spmv_dot"), (8, "update_xr_dot"), (9, "p_update"), (4, "allreduce"), (5, "exchange")):
                    acc[k] += out["times"][i]
            return out

        for _ in range(warmup):
            step()
        acc = {k: 0.0 for k in ("loop_ms", "iters", "spmv_dot", "update_xr_dot", "p_update", "allreduce", "exchange")}
        sampler = ClockSampler(local_rank)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if rank == 0:
            sampler.start()
        launches0 = H.launch_count()
        ev0.record()
        for _ in range(steps):
            last = step(acc)
        ev1.record()
        barrier()
        launches = H.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        ms_total = max_over_ranks(ev0.elapsed_time(ev1))
        niters = last["niters"]
        err = float((x - 1.0).abs().max().item())  # known answer x -> 1 (generate_matrix.cpp:284-286), outside the timed region
        err = max_over_ranks(err)
        n_total = n * size
        flops = steps * niters * FLOPS_PER_ROW_ITER * n_total
        res = {
            "n_local": n, "n_total": n_total, "niters": niters, "normr": last["normr"], "x_max_err": err,
            "ms_per_step": ms_total / steps, "value": flops / (ms_total * 1e-3) / 1e9, "launches": launches, "clocks": clocks,
            "setup_s": t_setup, "ell_bytes": m.bytes(), "slots": info["slots"], "format": m.format(), "comm": comm_name(m),
            "steps": steps, "warmup": warmup,
        }
        # per-kernel CUDA-event sums (this rank), on the stream the kernels run on
        it = max(acc["iters"], 1)
        kern = {}
        for k in ("spmv_dot", "update_xr_dot", "p_update"):
            sec = max_over_ranks(acc[k]) / it
            kern[k] = {"ms": sec * 1e3, "bytes": bpr[k] * n, "gbs": bpr[k] * n / sec / 1e9 if sec > 0 else None}
        loop_s = max_over_ranks(acc["loop_ms"]) * 1e-3 / it
        kern["iteration"] = {"ms": loop_s * 1e3, "bytes": bpr["iteration"] * n, "gbs": bpr["iteration"] * n / loop_s / 1e9}
        if size > 1:
            kern["allreduce_ms"] = max_over_ranks(acc["allreduce"]) / it * 1e3
            kern["exchange_ms"] = max_over_ranks(acc["exchange"]) / it * 1e3
            kern["launches_per_iteration"] = launches / max(steps * niters, 1)
        res["kernels"] = kern

        if do_e2e:
            # the reference-facing call with HOST buffers: b = generate_matrix's own (page-locked) vector, x = one
            # page-locked zero initial guess per step, prepared before the region (HPCCG() updates x in place)
            bh = A.b
            nbuf = max(1, min(steps, int(4e9 // (8 * n))))  # at most ~4 GB of page-locked initial guesses per rank
            xbufs = torch.zeros((nbuf, n), dtype=torch.float64, pin_memory=True)
            xh = A.x
            xh[:] = 0.0
            H.HPCCG(A, bh, xh, max_iter, 0.0)  # warm-up (allocates the staging buffers of the mirror)
            # Every step is bracketed by its own event pair (H2D of b and x, the solve, D2H of x inside); the sum over
            # the steps is the e2e time.  Between the brackets a reused guess buffer is reset to zero -- that is the
            # caller preparing the next step's input, not part of the call.
            e_ms_local = 0.0
            e_iters = 0
            for s_ in range(steps):
                xs = xbufs[s_ % nbuf].numpy()
                if s_ >= nbuf:
                    xs[:] = 0.0
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                e0.record()
                it_, normr_, _, _ = H.HPCCG(A, bh, xs, max_iter, 0.0)
                e1.record()
                torch.cuda.synchronize()
                e_ms_local += e0.elapsed_time(e1)
                e_iters += it_
            barrier()
            e_ms = max_over_ranks(e_ms_local)
            res["e2e"] = {"value": e_iters * FLOPS_PER_ROW_ITER * n_total / (e_ms * 1e-3) / 1e9, "unit": UNIT,
                          "h2d_bytes_per_step": 16 * n * size, "d2h_bytes_per_step": (8 * n + 8 * max_iter + 64) * size,
                          "ms_per_step": e_ms / steps, "x_max_err": max_over_ranks(float((xbufs[(steps - 1) % nbuf] - 1.0).abs().max())),
                          "host_buffers": "page-locked (generate_matrix's own b; cudaHostAlloc'd x)",
                          "copy_floor_ms": None}
            # what the three copies alone cost at the measured link rate: the part of the gap no ordering can remove
            # (x is needed before the first kernel, b before r = b - Ap, and the final x exists only after the last iteration)
            del xbufs
            if pageable:
                # the vectors the reference's generate_matrix hands out are plain new double[] (generate_matrix.cpp:233-235):
                # same call with PAGEABLE host memory (the driver stages such copies through its own pinned buffers)
                bp = np.array(bh, copy=True)
                xp = np.zeros(n)
                H.HPCCG(A, bp, xp, max_iter, 0.0)
                p_ms, p_it = 0.0, 0
                for _ in range(2):
                    xp[:] = 0.0
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    barrier()
                    e0.record()
                    it_, _, _, _ = H.HPCCG(A, bp, xp, max_iter, 0.0)
                    e1.record()
                    torch.cuda.synchronize()
                    p_ms += e0.elapsed_time(e1)
                    p_it += it_
                p_ms = max_over_ranks(p_ms)
                res["e2e"]["pageable"] = {"value": p_it * FLOPS_PER_ROW_ITER * n_total / (p_ms * 1e-3) / 1e9, "unit": UNIT,
                                          "ms_per_step": p_ms / 2, "steps": 2, "x_max_err": float(np.abs(xp - 1.0).max())}
        A.destroy()
        del b, x
        torch.cuda.empty_cache()
        return res

    def small_solves(nx, ny, nz, reps=40):
        """Launch-bound sizes (configs[0], 20x30x10): milliseconds per complete solve."""
        H.set_options(27, True)
        H.set_matrix_format("sell")
        A = H.generate_matrix(nx, ny, nz)
        m = A.device()
        n = nx * ny * nz
        out = {"workload": f"configs[0]: 27-pt {nx}x{ny}x{nz}, 150 CG iterations", "rows": n}
        # (1) the reference-named call with host vectors (wall clock; the call is synchronous)
        xh = A.x
        walls = []
        for i in range(reps + 3):
            xh[:] = 0.0
            t0 = time.perf_counter()
            it_, nr_, _, hist = H.HPCCG(A, A.b, xh, MAX_ITER, 0.0)
            t1 = time.perf_counter()
            if i >= 3:
                walls.append((t1 - t0) * 1e3)
        out["hpccg_host_call_ms"] = statistics.median(walls)
        out["niters"] = it_
        out["normr"] = nr_
        out["x_max_err"] = float(np.abs(xh - 1.0).max())
        # (2) device-resident vectors: graph replay and (where built) the single persistent kernel
        b = torch.from_numpy(A.b).to(dev)
        x = torch.zeros(n, dtype=torch.float64, device=dev)
        modes = [("graph_replay", 16)]
        if hasattr(H, "SOLVE_PERSISTENT"):
            modes.append(("persistent_kernel", H.SOLVE_PERSISTENT))
        for name, fl in modes:
            devms, wl = [], []
            for i in range(reps + 3):
                x.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                o = H.dev.cg_solve(m, b, x, MAX_ITER, 0.0, flags=fl, want_hist=False)
                t1 = time.perf_counter()
                if i >= 3:
                    devms.append(o["loop_ms"])
                    wl.append((t1 - t0) * 1e3)
            out[name] = {"device_ms": statistics.median(devms), "call_ms": statistics.median(wl), "niters": o["niters"],
                         "normr": o["normr"], "x_max_err": float((x - 1.0).abs().max().item())}
        best = min(v["call_ms"] for k, v in out.items() if isinstance(v, dict) and "call_ms" in v)
        out["ms_per_solve"] = best
        out["value"] = out["niters"] * FLOPS_PER_ROW_ITER * n / (best * 1e-3) / 1e9
        out["unit"] = UNIT
        A.destroy()
        return out

    def parity_block():
        """N > 1: BASELINE configs[1]'s global problem (256^3) split in z over the ranks, through the reference-named call
        HPCCG(A, b, x, ...) on both data planes, against the committed residual history of the real reference."""
        gpath = ROOT / "tests" / "golden" / "golden_256.json"
        if not gpath.exists() or 256 % size:
            return {"skipped": "no golden_256.json or 256 planes do not split over the ranks"}
        g = json.loads(gpath.read_text())
        ref = np.array([float.fromhex(v) for v in g["hist"]])
        out = {"problem": f"27-pt global 256x256x256 = {size} ranks x 256x256x{256 // size}, max_iter 150",
               "oracle": "tests/golden/golden_256.json (unmodified reference, serial build, tests/golden/make_golden.py)",
               "bar": 1e-8, "planes": {}}
        ok = True
        for comm in ("p2p", "nccl"):
            os.environ["HPCCG_B200_COMM"] = comm  # read when the mirror's peer link is negotiated (first solve)
            H.set_options(27, False)
            H.set_matrix_format("sell")
            A = H.generate_matrix(256, 256, 256 // size)
            H.make_local_matrix(A)
            xh = A.x
            xh[:] = 0.0
            it_, nr_, _, hist = H.HPCCG(A, A.b, xh, MAX_ITER, 0.0)
            ran = ~np.isnan(hist) & ~np.isnan(ref)
            rel = np.abs(hist[ran] - ref[ran]) / ref[ran]
            regular = ran & (ref >= 1e-10 * ref[0])
            rel_reg = np.abs(hist[regular] - ref[regular]) / ref[regular]
            rec = {"comm": comm_name(A.device()), "niters": it_, "normr": nr_, "ref_normr": float(ref[g["niters"]]),
                   "worst_rel": max_over_ranks(float(rel.max())), "worst_rel_regular": max_over_ranks(float(rel_reg.max())),
                   "iterations_compared": int(ran.sum()), "x_max_err": max_over_ranks(float(np.abs(xh - 1.0).max()))}
            rec["ok"] = bool(it_ == g["niters"] and rec["worst_rel_regular"] <= 1e-8 and rec["x_max_err"] <= 1e-12)
            ok = ok and rec["ok"]
            out["planes"][comm] = rec
            A.destroy()
            torch.cuda.empty_cache()
        os.environ.pop("HPCCG_B200_COMM", None)
        out["worst_rel"] = max(v["worst_rel"] for v in out["planes"].values())
        out["ok"] = ok
        return out

    side = max(1, min(args.side_steps, args.steps))
    main = measure(w["nx"], w["ny"], w["nz"], w["stencil"], args.steps, args.warmup, not args.no_e2e, args.max_iter,
                   pageable=(size == 1 and not args.no_also))
    default_run = w["name"] == "weak512" and not args.no_also
    also = None
    if size == 1 and default_run:
        c2 = WORKLOADS["c2"]
        also = measure(c2["nx"], c2["ny"], c2["nz"], c2["stencil"], max(args.steps, 3), max(args.warmup, 3), not args.no_e2e,
                       args.max_iter)

    also_pattern = None
    if size == 1 and default_run and args.format == "sell":
        # the opt-in pattern-coded mirror on the same workload: same results bit for bit, fewer bytes
        also_pattern = measure(w["nx"], w["ny"], w["nz"], w["stencil"], side, 3, not args.no_e2e, args.max_iter, fmt="pattern")
    H.set_matrix_format(args.format)

    also_c3 = also_c1 = also_strong = parity = None
    if size == 1 and default_run and args.format == "sell":
        c3 = WORKLOADS["c3"]
        also_c3 = measure(c3["nx"], c3["ny"], c3["nz"], c3["stencil"], side, 3, False, args.max_iter)
        also_c1 = small_solves(20, 30, 10)
    if default_run and args.format == "sell" and not args.no_strong and 1024 % size == 0:
        # configs[4] on THIS lease: at N = 1 the base of the strong-scaling curve, at N > 1 its points
        st = WORKLOADS["strong"]
        also_strong = measure(st["nx"], st["ny"], st["nz_global"] // size, st["stencil"], side, 3, False, args.max_iter)
    if size > 1 and default_run and not args.no_parity:
        parity = parity_block()

    cpu = None
    if rank == 0 and size == 1 and not args.no_cpu_baseline:
        try:
            rates, info = run_cpu_reference(w, args.cpu_iters, 1, 1)
            cpu = dict(info)
            cpu["value"] = statistics.mean(rates)
            cpu["unit"] = UNIT
            if default_run and not args.no_ref_c2:
                cpu["c2"] = run_cpu_c2()
                cpu["c1"] = run_cpu_c1()
        except Exception as e:  # noqa: BLE001 - the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}

    if size > 1:
        from hpccg_sycl_b200 import dist as hdist
        hdist.finalize()
    if rank != 0:
        if parity and not parity.get("ok", True):
            sys.exit(3)
        return
    peak, peak_src = measured_peak()
    k = main["kernels"]["spmv_dot"]
    roofline = {"bound": "hbm", "achieved": k["gbs"], "peak": peak, "unit": "GB/s", "frac": (k["gbs"] or 0) / peak,
                "traffic": ncu_traffic(w["name"] + ("" if args.format == "sell" else "-" + args.format)), "kernel": (f"spmv_pattern_kernel<{main['slots']},true>" if main["format"]["format"] == 1 else
                           f"spmv_sell_tma_kernel<{main['slots']},...,true>" if main["slots"] in (7, 27) and os.environ.get("HPCCG_B200_SPMV") != "reg"
                           else f"spmv_ell_kernel<{main['slots']},2,true>") + " (fused SpMV + p.Ap)",
                "algorithmic_bytes_per_launch": k["bytes"], "ms_per_launch": k["ms"], "peak_source": peak_src,
                "loop": {"bytes_per_iteration": main["kernels"]["iteration"]["bytes"],
                         "ms_per_iteration": main["kernels"]["iteration"]["ms"], "achieved": main["kernels"]["iteration"]["gbs"],
                         "frac": main["kernels"]["iteration"]["gbs"] / peak, "frac_of_8TBs_nominal": main["kernels"]["iteration"]["gbs"] / 8000.0},
                "kernels": main["kernels"]}
    config = dict(config)
    config["comm"] = main["comm"]
    if args.numa is not None:
        config["host_placement"] = args.numa
    if main.get("e2e"):
        # the three host<->device copies of a step at the link rate this run achieved end to end (gap / bytes): reported
        # so that the e2e / value gap can be read against the bytes that must cross PCIe whatever the ordering
        gap_ms = main["e2e"]["ms_per_step"] - main["ms_per_step"]
        main["e2e"].pop("copy_floor_ms", None)
        main["e2e"]["gap_ms_per_step"] = gap_ms
        main["e2e"]["gap_note"] = ("24 B/row must cross PCIe per step (x up before the first kernel, b up before r = b - Ap, x down after "
                                   "the last iteration); the set-up SpMV runs under b's upload and x returns in chunks behind x_fixup")
    line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "roofline": roofline, "cpu_baseline": cpu,
            "e2e": main.get("e2e"), "gpu_launches": main["launches"], "clocks": main["clocks"],
            "check": {"niters": main["niters"], "normr": main["normr"], "x_max_err": main["x_max_err"]},
            "setup_s": main["setup_s"]}
    if also:
        ka = also["kernels"]
        line["also"] = {"workload": "configs[1]: 27-pt 256x256x256, 150 CG iterations", "value": also["value"], "unit": UNIT,
                        "ms_per_step": also["ms_per_step"], "e2e": also.get("e2e"),
                        "roofline": {"achieved": ka["spmv_dot"]["gbs"], "frac": ka["spmv_dot"]["gbs"] / peak,
                                     "loop_achieved": ka["iteration"]["gbs"], "loop_frac": ka["iteration"]["gbs"] / peak,
                                     "kernels": ka},
                        "traffic": ncu_traffic("c2"),
                        "check": {"niters": also["niters"], "normr": also["normr"], "x_max_err": also["x_max_err"]}}
        if cpu and isinstance(cpu.get("c2"), dict) and cpu["c2"].get("value"):
            r = cpu["c2"]
            line["also"]["reference"] = r
            line["also"]["same_config_ratio"] = {
                "gpu_value_over_reference": also["value"] / r["value"],
                "gpu_e2e_over_reference": (also["e2e"]["value"] / r["value"]) if also.get("e2e") else None,
                "what": "identical problem on both sides: 27-pt 256^3, 149 CG iterations, HPCCG accounting; the reference's own "
                        f"OpenMP build on {r['cores']} host threads of this box"}
    if also_pattern:
        kp = also_pattern["kernels"]
        line["also_pattern_format"] = {
            "what": "same workload with the opt-in pattern-coded mirror (hpccg_dev_matrix_compress): one 16-bit pattern id per "
                    "row instead of 12 B per stored entry; bit-identical SpMV",
            "value": also_pattern["value"], "unit": UNIT, "ms_per_step": also_pattern["ms_per_step"], "e2e": also_pattern.get("e2e"),
            "steps": also_pattern["steps"],
            "bytes_per_row_iteration": bytes_per_row_iter(w["stencil"], "pattern", args.eager_x or args.unfused)["iteration"],
            "mirror_bytes": also_pattern["ell_bytes"], "patterns": also_pattern["format"]["patterns"],
            "kernels": kp, "loop_frac_of_peak": kp["iteration"]["gbs"] / peak, "traffic": ncu_traffic("weak512-pattern"),
            "check": {"niters": also_pattern["niters"], "normr": also_pattern["normr"], "x_max_err": also_pattern["x_max_err"]}}
    if also_c3:
        k3 = also_c3["kernels"]
        line["also_c3"] = {"workload": "configs[2]: 7-pt 512x512x512, 150 CG iterations (HPCCG accounting: 64 flops/row/iteration as the "
                                       "reference reports for any stencil)", "value": also_c3["value"], "unit": UNIT,
                           "ms_per_step": also_c3["ms_per_step"], "steps": also_c3["steps"], "bytes_per_row_iteration": bytes_per_row_iter(7)["iteration"],
                           "kernels": k3, "loop_frac_of_peak": k3["iteration"]["gbs"] / peak, "traffic": ncu_traffic("c3"),
                           "check": {"niters": also_c3["niters"], "normr": also_c3["normr"], "x_max_err": also_c3["x_max_err"]}}
    if also_c1:
        line["also_c1"] = also_c1
        if cpu and isinstance(cpu.get("c1"), dict):
            line["also_c1"]["reference"] = cpu["c1"]
    if also_strong:
        ks = also_strong["kernels"]
        line["also_strong"] = {"workload": f"configs[4]: 27-pt global 512x512x1024 split in z over {size} GPU(s): local 512x512x{1024 // size}",
                               "scaling": "strong", "value": also_strong["value"], "unit": UNIT, "n_gpus": size,
                               "ms_per_step": also_strong["ms_per_step"], "ms_per_iteration": ks["iteration"]["ms"],
                               "exchange_ms": ks.get("exchange_ms"), "allreduce_ms": ks.get("allreduce_ms"),
                               "launches_per_iteration": ks.get("launches_per_iteration"),
                               "steps": also_strong["steps"], "comm": also_strong["comm"], "clocks": also_strong["clocks"],
                               "kernels": ks,
                               "check": {"niters": also_strong["niters"], "normr": also_strong["normr"], "x_max_err": also_strong["x_max_err"]}}
    if parity:
        line["parity"] = parity
    print(json.dumps(line), flush=True)
    if parity and not parity.get("ok", True):
        sys.exit(3)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if size != args.gpus and size > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={size}")
    if args.gpus > 1 and size == 1:
        raise SystemExit("for N > 1 launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                         "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
    args.numa = None
    if size > 1 and args.impl != "reference" and not os.environ.get("HPCCG_BENCH_NO_NUMA"):
        # one process per GPU: run on (and first-touch the page-locked host vectors from) the CPUs next to that GPU
        try:
            import hpccg_pkg
            hpccg_pkg.load()
            from hpccg_sycl_b200 import dist as hdist
            args.numa = hdist.bind_to_gpu_numa_node(local_rank)
        except Exception as e:  # noqa: BLE001
            args.numa = {"bound": False, "error": str(e)[:120]}
    w = resolve_workload(args, size)
    bpr = bytes_per_row_iter(w["stencil"], args.format, args.eager_x or args.unfused)
    config = {"workload": f"{w['name']}: {w['stencil']}-pt stencil, local {w['nx']}x{w['ny']}x{w['nz']} per GPU, "
                          f"global {w['nx']}x{w['ny']}x{w['nz'] * size}, max_iter {args.max_iter} ({args.max_iter - 1} CG iterations per step)",
              "nx": w["nx"], "ny": w["ny"], "nz_local": w["nz"], "stencil": w["stencil"], "max_iter": args.max_iter,
              "ranks": size, "matrix_format": args.format, "decomposition": "1-D in z, one rank per GPU (generate_matrix.cpp:225-229)",
              "bytes_per_row_iteration": bpr["iteration"],
              "l2": "no flush: one iteration streams %.1f GB per GPU, >> 126 MB L2" % (bpr["iteration"] * w["nx"] * w["ny"] * w["nz"] / 1e9)}
    if args.impl == "reference":
        reference_arm(args, w, config, rank, size)
        return
    if size > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        try:
            gpu_arm(args, w, config, rank, size, local_rank)
        finally:
            dist.destroy_process_group()
    else:
        gpu_arm(args, w, config, rank, size, local_rank)


if __name__ == "__main__":
    main()
