#!/usr/bin/env python
"""bench.py -- particle.k.frame evaluations per second of the space-time correlation hot path.

    python bench.py --gpus 1 --steps K --warmup W                      (our arm, one process)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W      (our arm, N ranks)
    python bench.py --impl reference --gpus N --steps K --warmup W     (the reference's CPU code)

A *step* is one processed frame of the workload in steady state (origin ring full: every frame
captures one origin, is correlated against all B_live = nCBuffer live origins, and completes one
buffer).  Unit of work: one (particle, axial k-vector, frame) triple,
including its contribution to rho(k) and its self terms against all live origins:
``units = N * K * frames``, ``K = 3 nCSpatial``.

Default workload = BASELINE.json config 5 / the north-star target: 16,078,716 atoms (fcc 159^3),
nCSpatial 171 (K = 513), nCTime = nCBuffer = 150, sharded by particle over the ranks (strong
scaling: the total atom count is fixed).  The other configs (--workload c1..c4) are parity-test
shapes; they can be timed too but are not the headline.

Timed legs (all CUDA-event timed on the engine's stream, max over ranks):
  value : K steady-state frames whose positions are already resident in HBM (device frame pool).
  e2e   : K steady-state frames through the public host API: every step copies that step's
          positions from pinned HOST memory to the device and reads the step's reduced rho(k)
          vector back to the host (sqt_push_frame + sqt_get_rho).
  roofline : the ring kernel's own launches, event-timed inside the `value` leg.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_cells, nCSpatial, nCTime, nCBuffer, description)
    "c1": (10, 30, 150, 150, "config 1: 4,000 atoms, in.lmp2sqt.150 (nCSpatial 30, K=90), ring 150/150"),
    "c2": (20, 10, 150, 150, "config 2: 32,000 atoms, in.lmp2sqt.150.smallQ (nCSpatial 10, K=30), ring 150/150"),
    "c3": (63, 86, 150, 150, "config 3: 1,000,188 atoms, nCSpatial 86 (K=258), ring 150/150"),
    "c4": (100, 128, 150, 150, "config 4: 4,000,000 atoms, nCSpatial 128 (K=384), ring 150/150"),
    "c5": (159, 171, 150, 150, "config 5: 16,078,716 atoms, nCSpatial 171 (K=513), ring 150/150"),
}
METRIC = "particle_k_frame_evals_per_sec"
UNIT = "particle*k*frame/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_sample(wl, steps, warmup, budget_s=25.0):
    """The reference's own CPU implementation (oracle/_ref, unmodified) on a bounded sample of the workload:
    same nCSpatial / ring, fewer atoms (its cost is linear in N; its -F memory is (B+1)*N*6*nCS*16 B, which is
    1.2 TB at 1 M atoms, so same-size runs are impossible -- BASELINE.md section 3.4)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    n_cells, S, T, B, _ = WORKLOADS[wl]
    threads = host_threads()
    # probe the per-atom frame cost on a tiny system, then size the sample to ~budget_s of CPU wall
    probe = pyoracle.bench_reference(32, S, T, B, fill=B, warmup=1, steps=2, threads=threads)
    per_atom_frame = max(1e-9, statistics.mean(probe["step_times"]) / 32)
    frames_equiv = B / 2.0 + warmup + steps     # ring fill costs ~B/2 full frames
    n = int(budget_s / (per_atom_frame * frames_equiv))
    mem_cap = int(6e9 / ((B + 1) * 6 * S * 16))  # keep the reference's tables under ~6 GB
    n = max(32, min(n, mem_cap, 4 * n_cells**3))
    out = pyoracle.bench_reference(n, S, T, B, fill=B, warmup=warmup, steps=steps, threads=threads)
    per_step = statistics.mean(out["step_times"])
    return {"n_atoms": n, "threads": threads, "per_step_s": per_step, "value": n * 3 * S / per_step,
            "step_times": out["step_times"], "fill_s": out["fill_s"]}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_cells, S, T, B, desc = WORKLOADS[args.workload]
    r = reference_sample(args.workload, args.steps, args.warmup, budget_s=60.0)
    sample = ("%d atoms of the workload (same nCSpatial %d, nCTime %d, nCBuffer %d), steady-state frames after a %d-frame "
              "ring fill; unmodified reference EvalSpacetimeCorr (oracle/_ref/liblmp2sqt_ref.so), OMP_NUM_THREADS=%d"
              % (r["n_atoms"], S, T, B, B, r["threads"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["per_step_s"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + desc, "B_live": B, "K": 3 * S, "sample_atoms": r["n_atoms"]},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference", "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--atoms-cells", type=int, default=0, help="override the fcc cell count (debugging only; invalidates the line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from lmp2sqt_b200 import Deck, SpacetimeCorr, SynthSource, comm_unique_id, measure_dfma_peak, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torchrun (one rank per GPU)" % args.gpus)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n_cells, S, T, B, desc = WORKLOADS[args.workload]
    if args.atoms_cells:
        n_cells = args.atoms_cells
        desc += " [DEBUG OVERRIDE: %d cells]" % n_cells
    n_total = 4 * n_cells**3
    lo, hi = shard_bounds(n_total, rank, world)
    n_local = hi - lo
    K, W = args.steps, args.warmup
    Kvec = 3 * S

    deck = Deck(limitCorrAv=10**9, nCBuffer=B, nCSpatial=S, nCTime=T)
    eng = SpacetimeCorr(deck, "Ft", device=local_rank)
    if world > 1:
        ids = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(ids[0], rank, world)
    src = SynthSource(n_cells, lo, n_local, device=local_rank)
    L = src.L
    eng.InitSpacetimeCorr()

    # ---- setup (untimed): fill the origin ring so that every timed frame sees B live origins ----------
    t_fill = time.time()
    for _ in range(T):
        eng.push_device(src.next(), n_local, n_total, 0.0, L)
    eng.sync()
    t_fill = time.time() - t_fill

    frame_bytes = 3 * n_local * 8
    # ---- value leg: frames already resident in HBM -----------------------------------------------------
    pool_n = max(1, min(K, int(12e9 // frame_bytes)))
    pool = [torch.empty(3 * n_local, dtype=torch.float64, device="cuda") for _ in range(pool_n)]
    for _ in range(W):
        eng.push_device(src.next(), n_local, n_total, 0.0, L)
    frames_in_pool = 0

    def refill_pool(count):
        for i in range(count):
            src.next()
            src.to_device(pool[i].data_ptr())
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    total_ms = 0.0
    done = 0
    launches0 = eng.kernel_launches
    eng.profile(True)
    if rank == 0:
        sampler.start()
    while done < K:
        chunk = min(pool_n, K - done)
        refill_pool(chunk)
        barrier()
        eng.timer_start()
        for i in range(chunk):
            eng.push_device(pool[i].data_ptr(), n_local, n_total, 0.0, L)
        ms = eng.timer_stop()
        barrier()
        total_ms += max_over_ranks(ms)
        done += chunk
    clocks = sampler.stop() if rank == 0 else None
    n_ring, ring_ms = eng.profile_read()
    eng.profile(False)
    launches = eng.kernel_launches - launches0
    value = n_total * Kvec * K / (total_ms * 1e-3)
    ms_per_step = total_ms / K
    ring_ms_avg = max_over_ranks(ring_ms / max(1, n_ring))

    # ---- e2e leg: host buffers, H2D inside the timed region, result read back every step ---------------
    e2e = None
    if not args.no_e2e:
        hp_n = max(1, min(K, int(6e9 // frame_bytes)))
        hpool = [torch.empty(3 * n_local, dtype=torch.float64).pin_memory() for _ in range(hp_n)]
        rho_host = np.zeros((3, S, 2))
        e2e_ms = 0.0
        done = 0
        import ctypes as C
        while done < K:
            chunk = min(hp_n, K - done)
            for i in range(chunk):
                src.next()
                src.to_host(hpool[i].numpy().reshape(3, n_local))
            barrier()
            t0 = time.perf_counter()
            eng.timer_start()
            for i in range(chunk):
                rc = eng.lib.sqt_push_frame(eng.h, C.c_void_p(hpool[i].data_ptr()), n_local, n_total, 0.0, float(L))
                if rc < 0:
                    raise RuntimeError(eng.lib.sqt_last_error(eng.h).decode())
                eng.lib.sqt_get_rho(eng.h, rho_host.ctypes.data)   # D2H read of the step's result (synchronises)
            ms = eng.timer_stop()
            wall_ms = (time.perf_counter() - t0) * 1e3
            barrier()
            e2e_ms += max_over_ranks(max(ms, wall_ms))
            done += chunk
        e2e = {"value": n_total * Kvec * K / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(sum_over_ranks(frame_bytes)), "d2h_bytes_per_step": int(6 * S * 8 * world),
               "ms_per_step": e2e_ms / K, "api": "sqt_push_frame(pinned host SoA) + sqt_get_rho per step"}

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    hbm_peak, hbm_src = load_peaks()
    dfma_tf, dfma_mhz = measure_dfma_peak(local_rank)
    flops_per_unit = 6 + 3 * B                       # rho 2 FMA + 2 ADD, self 1 FMA + 1 ADD per live origin
    units_per_launch = n_local * Kvec
    ach_tf = units_per_launch * flops_per_unit / (ring_ms_avg * 1e-3) / 1e12
    alg_bytes = n_local * 24.0 * (1 + B)             # current positions + B origin positions per particle
    cfg_now = eng.kernel_config
    on_mma = cfg_now.startswith("self_ring_mma_kernel")
    family = "self_ring_mma_kernel" if on_mma else ("self_ring_vec_kernel" if cfg_now.startswith("self_ring_vec_kernel") else "self_ring_kernel")
    if on_mma:
        # what the tensor-core path executes instead of the recurrence: per (particle, origin, axis) an 8 x 8NT x 1 slice of
        # the blocked-Chebyshev product on DMMA (the share of the FP64 pipe those DMMAs occupy is dmma_pipe_frac)
        NT = int(cfg_now.split("NT=")[1].split(",")[0])
        dmma_flops = n_local * 3.0 * (B * 2 * 8 * 8 * NT + 2 * 2 * 8 * 8 * NT)
        pipe_note = {"dmma_flops_per_launch": dmma_flops, "dmma_pipe_frac": dmma_flops / (ring_ms_avg * 1e-3) / (dfma_tf * 1e12),
                     "fp64_pipe_frac": None}
    else:
        pipe_note = {"fp64_pipe_frac": units_per_launch * (4 + 2 * B) / (ring_ms_avg * 1e-3) / (dfma_tf * 1e12 / 2.0)}
    # "tensor" = the FP64 tensor-core (DMMA) path; it shares one pipe with vector DFMA and has the same measured peak
    # (profiles/probes/dmma_probe.cu), so one denominator serves both kernel families
    roofline = {"bound": "tensor" if on_mma else "fp64_fma",
                "bound_detail": "FP64 pipe: DMMA m8n8k4 (tensor) and vector DFMA/DADD share it and have the same measured peak; "
                                "HBM is at a few percent (roofline.hbm)",
                "kernel": family + (" + rho_mma_kernel" if "rho_mma_kernel" in cfg_now else ""),
                "achieved": ach_tf, "peak": dfma_tf, "unit": "TFLOP/s",
                "frac": ach_tf / dfma_tf, "traffic": None,
                "peak_source": "measured live: register-resident DFMA microbenchmark (sqt_measure_dfma_peak), %.0f MHz implied" % dfma_mhz,
                "flops_per_unit": flops_per_unit, "units_per_launch": units_per_launch, "avg_launch_ms": ring_ms_avg,
                **pipe_note,
                "hbm": {"achieved": alg_bytes / (ring_ms_avg * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (ring_ms_avg * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                        "algorithmic_bytes_per_launch": alg_bytes}}
    prof_json = os.path.join(ROOT, "profiles", "ring_traffic.json")
    if os.path.exists(prof_json):
        try:
            with open(prof_json) as f:
                tr = json.load(f)
            if tr.get("workload") == args.workload and tr.get("n_gpus") == world and tr.get("kernel", "").startswith(family + "<"):
                roofline["traffic"] = tr.get("dram_bytes_per_launch")
        except Exception:
            pass

    total_launches = int(sum_over_ranks(launches))
    cfg = eng.kernel_config
    eng.close()
    src.close()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = reference_sample(args.workload, steps=3, warmup=1, budget_s=20.0)
            cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                            "sample": "%d atoms of the workload (same nCSpatial %d and %d/%d ring), 3 steady-state frames after the ring "
                                      "fill, unmodified reference EvalSpacetimeCorr via oracle/_ref, OMP_NUM_THREADS=%d"
                                      % (r["n_atoms"], S, T, B, r["threads"])}
        except Exception as ex:  # the baseline is a reported number, never a reason to lose the bench line
            cpu_baseline = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "reference", "sample": "failed: %s" % ex}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + desc, "atoms_total": n_total, "atoms_per_gpu": n_local, "K": Kvec,
                       "B_live": B, "parallelism": "particle-sharded x%d, per-frame NCCL allreduce of rho(k)" % world,
                       "l2": "inputs larger than L2: the origin ring (%s per GPU) is streamed once per frame" % (cfg.split("ring=")[1].split(" ")[0] if "ring=" in cfg else "?"),
                       "kernel": cfg, "ring_fill_s": t_fill,
                       "pair_evals_per_sec": value * B},
            "clocks": clocks, "e2e": e2e, "gpu_launches": total_launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
