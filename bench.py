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

--impl reference times ONLY that CPU reference (rank 0 alone under torchrun) and prints the same line shape.
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


def run_cpu_reference(w: dict, iters: int, steps: int, warmup: int):
    """Times the reference's HPCCG() (OpenMP build, all host threads) on the sample. Returns (gflops list, info)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import refwrap  # the checker / baseline; never on the product path
    if refwrap.available("omp"):
        variant, kind = "omp", "reference"
    elif refwrap.available("serial"):
        variant, kind = "serial", "reference"
    else:
        variant, kind = "oracle", "port"
    nx, ny, nz = cpu_sample_dims(w)
    n = nx * ny * nz
    t0 = time.time()
    R = refwrap.RefWorld(nx, ny, nz, size=1, stencil=w["stencil"], variant=variant)
    t_gen = time.time() - t0
    cores = R.threads()
    rates, secs = [], []
    for i in range(warmup + steps):
        s = R.solve(iters + 1, 0.0, hist=False, want_x=False)
        if i >= warmup:
            rates.append(s["niters"] * FLOPS_PER_ROW_ITER * n / s["times"][0] / 1e9)
            secs.append(float(s["times"][0]))
    R.close()
    info = {"kind": kind, "cores": cores, "variant": variant,
            "sample": f"{nx}x{ny}x{nz} z-slab of the workload ({w['stencil']}-pt), {iters} CG iterations per step, "
                      f"{steps} steps after {warmup} warm-up; reference HPCCG() times[0]; matrix generation {t_gen:.1f} s not timed",
            "seconds_per_step": secs}
    return rates, info


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
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": info_out,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
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

    def sum_over_ranks(v: float) -> float:
        if size == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def measure(nx, ny, nz, stencil, steps, warmup, do_e2e, max_iter, fmt=None):
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
            "setup_s": t_setup, "ell_bytes": m.bytes(), "slots": info["slots"], "format": m.format(),
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
                          "ms_per_step": e_ms / steps, "x_max_err": max_over_ranks(float((xbufs[(steps - 1) % nbuf] - 1.0).abs().max()))}
            del xbufs
        A.destroy()
        del b, x
        torch.cuda.empty_cache()
        return res

    main = measure(w["nx"], w["ny"], w["nz"], w["stencil"], args.steps, args.warmup, not args.no_e2e, args.max_iter)
    also = None
    if size == 1 and not args.no_also and w["name"] == "weak512":
        c2 = WORKLOADS["c2"]
        also = measure(c2["nx"], c2["ny"], c2["nz"], c2["stencil"], max(args.steps, 3), max(args.warmup, 3), not args.no_e2e,
                       args.max_iter)

    also_pattern = None
    if size == 1 and not args.no_also and w["name"] == "weak512" and args.format == "sell":
        # the opt-in pattern-coded mirror on the same workload: same results bit for bit, fewer bytes
        also_pattern = measure(w["nx"], w["ny"], w["nz"], w["stencil"], args.steps, args.warmup, not args.no_e2e, args.max_iter,
                               fmt="pattern")
    H.set_matrix_format(args.format)

    cpu = None
    if rank == 0 and size == 1 and not args.no_cpu_baseline:
        try:
            rates, info = run_cpu_reference(w, args.cpu_iters, 1, 1)
            cpu = dict(info)
            cpu["value"] = statistics.mean(rates)
            cpu["unit"] = UNIT
        except Exception as e:  # noqa: BLE001 - the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}

    if size > 1:
        from hpccg_sycl_b200 import dist as hdist
        hdist.finalize()
    if rank != 0:
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
                        "check": {"niters": also["niters"], "normr": also["normr"], "x_max_err": also["x_max_err"]}}
    if also_pattern:
        kp = also_pattern["kernels"]
        line["also_pattern_format"] = {
            "what": "same workload with the opt-in pattern-coded mirror (hpccg_dev_matrix_compress): one 16-bit pattern id per "
                    "row instead of 12 B per stored entry; bit-identical SpMV; its SpMV is bound by L1 gather throughput, not HBM",
            "value": also_pattern["value"], "unit": UNIT, "ms_per_step": also_pattern["ms_per_step"], "e2e": also_pattern.get("e2e"),
            "bytes_per_row_iteration": bytes_per_row_iter(w["stencil"], "pattern", args.eager_x or args.unfused)["iteration"],
            "mirror_bytes": also_pattern["ell_bytes"], "patterns": also_pattern["format"]["patterns"],
            "kernels": kp, "loop_frac_of_peak": kp["iteration"]["gbs"] / peak,
            "check": {"niters": also_pattern["niters"], "normr": also_pattern["normr"], "x_max_err": also_pattern["x_max_err"]}}
    print(json.dumps(line), flush=True)


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
