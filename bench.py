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


def parity_check(world, rank, local_rank, dist):
    """Driver-visible evidence that the sharded (NCCL) path computes the same numbers as the reference: before the timed
    legs a small trajectory at the headline's kernel shape (nCSpatial 171: tensor-core kernels, several particle tiles per
    rank) runs through a fresh engine on the same ranks, and rank 0's fully reduced raw sums are compared with the oracle
    restatement run on the host (oracle/: checker only -- nothing timed goes through it).  Tolerances: SURVEY.md section 8a."""
    import numpy as np
    from lmp2sqt_b200 import Deck, comm_unique_id, run_trajectories, synth
    n_cells, S, T, B, frames = 8, 171, 6, 6, 16
    traj = synth.make_trajectory(n_cells=n_cells, n_frames=frames, seed=20261021)
    ids = [comm_unique_id() if (rank == 0 and world > 1) else None]
    if world > 1:
        dist.broadcast_object_list(ids, src=0)
    deck = Deck(limitCorrAv=10**9, nCBuffer=B, nCSpatial=S, nCTime=T)
    res, eng = run_trajectories([traj], deck, "Ft", device=local_rank, rank=rank, nranks=world, unique_id=ids[0])
    got = res.to_dict()
    kernel = eng.kernel_config.split(" ")[0]
    eng.close()
    if rank != 0:
        return None
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    _, raw, _ = pyoracle.run_port(traj, S, T, B, limitCorrAv=10**9, finish=False)
    worst = 0.0
    for g, r in (("avF_qq2", "avF"), ("avF_s_qq2", "avF_s"), ("avF_d_qq2", "avF_d")):
        y = raw[r]
        worst = max(worst, float(np.max(np.abs(got[g] - y) / np.maximum(np.abs(y), 1e-3 * np.abs(y[:, :1])))))
    msd = max(float(np.max(np.abs(got[k][1:] - raw[k][1:]) / np.abs(raw[k][1:]))) for k in ("rrMSDAv", "rrMQDAv"))   # lag 0 is 0 = 0
    vh = bool(np.array_equal(got["avDrTable"], raw["avDrTable"]))
    ok = bool(worst <= 1e-6 and msd <= 1e-9 and vh and got["countCorrAv"] == raw["countCorrAv"])
    return {"n_ranks": world, "n_atoms": traj.n_atoms, "nCSpatial": S, "ring": "%d/%d" % (T, B), "frames": frames, "kernel": kernel,
            "checker": "oracle/sqt_oracle.c restatement (bit-identical to the unmodified reference, tests/test_oracle.py)",
            "max_rel_err": worst, "tolerance": 1e-6, "msd_mqd_max_rel_err": msd, "vanhove_exact": vh,
            "countCorrAv": int(got["countCorrAv"]), "ok": ok}


def time_workload(wl, K, W, world, rank, local_rank, dist, n_cells_override=0, with_e2e=True, sampler=None, n_types=0):
    """Fill the ring (untimed), then time K steady-state frames resident in HBM (`value`), the same through host buffers
    (`e2e`), and the dominant kernels' own launches (for the roofline).  Returns a dict of raw measurements."""
    import numpy as np
    import torch
    from lmp2sqt_b200 import Deck, SpacetimeCorr, SynthSource, comm_unique_id, shard_bounds

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    n_cells, S, T, B, desc = WORKLOADS[wl]
    if n_cells_override:
        n_cells = n_cells_override
        desc += " [DEBUG OVERRIDE: %d cells]" % n_cells
    n_total = 4 * n_cells**3
    lo, hi = shard_bounds(n_total, rank, world)
    n_local = hi - lo
    Kvec = 3 * S
    deck = Deck(limitCorrAv=10**9, nCBuffer=B, nCSpatial=S, nCTime=T)
    eng = SpacetimeCorr(deck, "Ft", device=local_rank, n_types=n_types)
    if world > 1:
        ids = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(ids[0], rank, world)
    if n_types > 1:      # "types {1,2} 80:20 by id"
        eng.set_types(np.where((np.arange(lo, hi) + 1) % 5 == 0, 2, 1).astype(np.int32))
        desc += "; per-type partials, %d species 80:20 by id" % n_types
    src = SynthSource(n_cells, lo, n_local, device=local_rank)
    L = src.L
    eng.InitSpacetimeCorr()
    t_fill = time.time()
    for _ in range(T):                       # setup (untimed): every timed frame sees B live origins
        eng.push_device(src.next(), n_local, n_total, 0.0, L)
    eng.sync()
    t_fill = time.time() - t_fill

    frame_bytes = 3 * n_local * 8
    pool_n = max(1, min(K, int(12e9 // frame_bytes)))
    pool = [torch.empty(3 * n_local, dtype=torch.float64, device="cuda") for _ in range(pool_n)]
    for _ in range(W):
        eng.push_device(src.next(), n_local, n_total, 0.0, L)
    total_ms, done = 0.0, 0
    launches0 = eng.kernel_launches
    eng.profile(True)
    if sampler is not None:
        sampler.start()
    while done < K:
        chunk = min(pool_n, K - done)
        for i in range(chunk):
            src.next()
            src.to_device(pool[i].data_ptr())
        barrier()
        eng.timer_start()
        for i in range(chunk):
            eng.push_device(pool[i].data_ptr(), n_local, n_total, 0.0, L)
        ms = eng.timer_stop()
        barrier()
        total_ms += reduce_ranks(ms, dist.ReduceOp.MAX if world > 1 else None)
        done += chunk
    clocks = sampler.stop() if sampler is not None else None
    n_ring, ring_ms = eng.profile_read()
    eng.profile(False)
    out = {"wl": wl, "desc": desc, "S": S, "T": T, "B": B, "Kvec": Kvec, "n_total": n_total, "n_local": n_local, "t_fill": t_fill,
           "launches": eng.kernel_launches - launches0, "total_ms": total_ms, "ms_per_step": total_ms / K,
           "value": n_total * Kvec * K / (total_ms * 1e-3), "clocks": clocks, "frame_bytes": frame_bytes,
           "ring_ms_avg": reduce_ranks(ring_ms / max(1, n_ring), dist.ReduceOp.MAX if world > 1 else None), "cfg": eng.kernel_config, "e2e": None}
    del pool

    if with_e2e:     # host buffers, H2D inside the timed region, the step's result read back every step
        hp_n = max(1, min(K, int(6e9 // frame_bytes)))
        hpool = [torch.empty(3 * n_local, dtype=torch.float64).pin_memory() for _ in range(hp_n)]
        rho_host = np.zeros((3, S, 2))
        e2e_ms, done = 0.0, 0
        import ctypes as C
        for i in range(W):   # warm-up of THIS path (untimed): host frames in, the step's result out, like the timed steps below
            src.next()
            src.to_host(hpool[i % hp_n].numpy().reshape(3, n_local))
            rc = eng.lib.sqt_push_frame(eng.h, C.c_void_p(hpool[i % hp_n].data_ptr()), n_local, n_total, 0.0, float(L))
            if rc < 0:
                raise RuntimeError(eng.lib.sqt_last_error(eng.h).decode())
            eng.lib.sqt_get_rho(eng.h, rho_host.ctypes.data)
        eng.sync()
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
                eng.lib.sqt_get_rho(eng.h, rho_host.ctypes.data)   # D2H read of the step's result (waits for that step's rho stage)
            ms = eng.timer_stop()
            wall_ms = (time.perf_counter() - t0) * 1e3
            barrier()
            e2e_ms += reduce_ranks(max(ms, wall_ms), dist.ReduceOp.MAX if world > 1 else None)
            done += chunk
        out["e2e"] = {"value": n_total * Kvec * K / (e2e_ms * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": int(reduce_ranks(frame_bytes, dist.ReduceOp.SUM if world > 1 else None)),
                      "d2h_bytes_per_step": int(6 * S * 8 * world), "ms_per_step": e2e_ms / K,
                      "api": "sqt_push_frame(pinned host SoA) + sqt_get_rho per step; the next step's host-to-device copy runs under this "
                             "step's self-term kernel (sqt_get_rho waits for the step's rho(k) stage only)"}
    out["launches"] = int(reduce_ranks(out["launches"], dist.ReduceOp.SUM if world > 1 else None))
    eng.close()
    src.close()
    return out


def roofline_of(m, dfma_tf, dfma_mhz, hbm_peak, hbm_src, world):
    """Roofline of the dominant kernels of a measured workload (DESIGN.md sections 4.2, 4.3, 7)."""
    B, Kvec, n_local = m["B"], m["Kvec"], m["n_local"]
    flops_per_unit = 6 + 3 * B                       # rho 2 FMA + 2 ADD, self 1 FMA + 1 ADD per live origin
    units_per_launch = n_local * Kvec
    ring_ms_avg = m["ring_ms_avg"]
    ach_tf = units_per_launch * flops_per_unit / (ring_ms_avg * 1e-3) / 1e12
    alg_bytes = n_local * 24.0 * (1 + B)             # current positions + B origin positions per particle
    cfg_now = m["cfg"]
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
                "frac": ach_tf / dfma_tf, "traffic": None, "traffic_source": None,
                "peak_source": "measured in THIS run by this repo's own register-resident DFMA microbenchmark (sqt_measure_dfma_peak; %.0f MHz "
                               "implied); MEASURED_PEAKS.json carries no FP64 peak; nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s "
                               "(frac_of_nominal below)" % dfma_mhz,
                "frac_of_nominal": ach_tf / 37.2,
                "achieved_source": "algorithmic flops (SURVEY.md 8d) / CUDA-event time of the rho(k) + self-term kernels' own launches, "
                                   "measured in this run on their stream",
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
            if tr.get("workload") == m["wl"] and tr.get("n_gpus") == world and tr.get("kernel", "").startswith(family + "<"):
                roofline["traffic"] = tr.get("dram_bytes_per_launch")
                roofline["traffic_source"] = ("NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one launch from the "
                                              "committed ncu --set full capture " + str(tr.get("source")))
        except Exception:
            pass
    return roofline


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
    ap.add_argument("--no-other-workloads", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from lmp2sqt_b200 import measure_dfma_peak

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

    # ---- parity of the (sharded) path against the oracle, before anything is timed ------------------------------------
    pcheck = None
    if not args.no_parity_check:
        pcheck = parity_check(world, rank, local_rank, dist)
        flag = torch.tensor([1.0 if (rank != 0 or pcheck["ok"]) else 0.0], device="cuda")
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() < 1.0:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "error": "parity_check failed: no number is reported",
                                  "parity_check": pcheck}), flush=True)
            raise SystemExit(1)

    K, W = args.steps, args.warmup
    gpu_index = local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank])
    sampler = ClockSampler(gpu_index) if rank == 0 else None
    m = time_workload(args.workload, K, W, world, rank, local_rank, dist, args.atoms_cells, with_e2e=not args.no_e2e, sampler=sampler)

    hbm_peak, hbm_src = load_peaks()
    dfma_tf, dfma_mhz = measure_dfma_peak(local_rank)
    roofline = roofline_of(m, dfma_tf, dfma_mhz, hbm_peak, hbm_src, world)

    # ---- the other BASELINE configs (parity-test shapes, not the headline): a few steady-state frames each, one GPU -----
    others = None
    if world == 1 and not args.no_other_workloads and not args.atoms_cells:
        others = {}
        for wl in sorted(WORKLOADS):
            if wl == args.workload:
                continue
            try:
                mo = time_workload(wl, 20, 3, 1, 0, local_rank, dist, with_e2e=False)
                ro = roofline_of(mo, dfma_tf, dfma_mhz, hbm_peak, hbm_src, 1)
                others[wl] = {"workload": mo["desc"], "ms_per_step": mo["ms_per_step"], "value": mo["value"], "unit": UNIT, "steps": 20,
                              "kernel_ms": mo["ring_ms_avg"], "frac": ro["frac"], "bound": ro["bound"], "hbm_frac": ro["hbm"]["frac"],
                              "gpu_launches": mo["launches"], "kernel": mo["cfg"].split(";")[0]}
            except Exception as ex:  # never lose the headline line to a side measurement
                others[wl] = {"error": str(ex)}
        try:   # BASELINE config 4 as written: two species, all partials F_ab / F_s,a beside the totals
            mo = time_workload("c4", 20, 3, 1, 0, local_rank, dist, with_e2e=False, n_types=2)
            ro = roofline_of(mo, dfma_tf, dfma_mhz, hbm_peak, hbm_src, 1)
            others["c4_types"] = {"workload": mo["desc"], "ms_per_step": mo["ms_per_step"], "value": mo["value"], "unit": UNIT, "steps": 20,
                                  "kernel_ms": mo["ring_ms_avg"], "frac": ro["frac"], "bound": ro["bound"], "gpu_launches": mo["launches"],
                                  "kernel": mo["cfg"].split(";")[0]}
        except Exception as ex:
            others["c4_types"] = {"error": str(ex)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        S, T, B = m["S"], m["T"], m["B"]
        try:
            r = reference_sample(args.workload, steps=3, warmup=1, budget_s=20.0)
            cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                            "sample": "%d atoms of the workload (same nCSpatial %d and %d/%d ring), 3 steady-state frames after the ring "
                                      "fill, unmodified reference EvalSpacetimeCorr via oracle/_ref, OMP_NUM_THREADS=%d"
                                      % (r["n_atoms"], S, T, B, r["threads"])}
        except Exception as ex:  # the baseline is a reported number, never a reason to lose the bench line
            cpu_baseline = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "reference", "sample": "failed: %s" % ex}

    if rank == 0:
        cfg = m["cfg"]
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": m["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + m["desc"], "atoms_total": m["n_total"], "atoms_per_gpu": m["n_local"], "K": m["Kvec"],
                       "B_live": m["B"], "parallelism": "particle-sharded x%d, per-frame NCCL allreduce of rho(k)" % world,
                       "l2": "inputs larger than L2: the origin ring (%s per GPU) is streamed once per frame" % (cfg.split("ring=")[1].split(" ")[0] if "ring=" in cfg else "?"),
                       "kernel": cfg, "ring_fill_s": m["t_fill"],
                       "pair_evals_per_sec": m["value"] * m["B"]},
            "clocks": m["clocks"], "e2e": m["e2e"], "gpu_launches": m["launches"], "roofline": roofline, "cpu_baseline": cpu_baseline,
            "parity_check": pcheck, "other_workloads": others,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
