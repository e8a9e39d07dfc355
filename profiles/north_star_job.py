#!/usr/bin/env python
"""The whole north-star job, not a steady-state sample: F(q,t), F_s(q,t), MSD, MQD, self van Hove of 16 078 716 atoms,
513 axial k-vectors (nCSpatial 171), 5 000 frames, ring 150/150, frames produced on the device by the Philox source.
Run under torchrun (one rank per GPU) or plainly on one GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
        profiles/north_star_job.py [--frames 5000]

Prints one JSON line on rank 0: wall-clock seconds of the frame loop (max over ranks, device synchronised), the
bookkeeping result and the full-size invariants of tests/test_gpu_fullsize.py evaluated on the final averages."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lmp2sqt_b200 import Deck, SpacetimeCorr, SynthSource, comm_unique_id, shard_bounds  # noqa: E402

N_CELLS, S, T, B, SIGMA = 159, 171, 150, 150, 0.077


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=5000)
    ap.add_argument("--dump", default="", help="write rank 0's raw sums to this .npz (to compare runs on different GPU counts)")
    ap.add_argument("--cells", type=int, default=N_CELLS, help="fcc cells per edge (159 = the north-star 16 078 716 atoms)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 4 * args.cells ** 3
    lo, hi = shard_bounds(n, rank, world)
    n_local = hi - lo
    P = args.frames
    expected = sum(max(0, (P - nb) // T) for nb in range(B))      # SURVEY.md section 8a closed form (B == T: s_nb = nb)
    deck = Deck(limitCorrAv=expected, nCBuffer=B, nCSpatial=S, nCTime=T)  # the last completed buffer triggers the print, like the reference's count pass arranges
    eng = SpacetimeCorr(deck, "Ft", device=local)
    if world > 1:
        ids = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(ids[0], rank, world)
    src = SynthSource(args.cells, lo, n_local, device=local)
    eng.InitSpacetimeCorr()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    flushed_at = None
    for p in range(P):
        rc = eng.push_device(src.next(), n_local, n, 0.0, src.L)
        if rc == 1:
            flushed_at = p
    eng.sync()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    res = eng.fetch()                                             # the latched averages of the flush
    r = res.to_dict()
    if rank == 0:
        cnt = r["countCorrAv"]
        q = (np.arange(S) + 1) * 2 * np.pi / src.L
        tau = np.arange(T)
        Fs = r["avF_s_qq2"] / (3.0 * n * cnt)
        msd = r["rrMSDAv"] / (n * cnt)
        want = np.exp(-0.5 * (q[:, None] * SIGMA) ** 2 * tau[None, :])
        line = {
            "job": "north star: %d atoms x %d k x %d frames, ring %d/%d" % (n, 3 * S, P, T, B), "n_gpus": world,
            "seconds": float(dt.item()), "ms_per_frame": float(dt.item()) / P * 1e3,
            "particle_k_frame_evals_per_sec": n * 3.0 * S * P / float(dt.item()),
            "countCorrAv": int(cnt), "countCorrAv_expected": expected, "flushed_at_frame": flushed_at,
            "max_abs_Fs_q0_minus_1": float(np.abs(Fs[:, 0] - 1).max()),
            "max_abs_Fs_minus_gaussian_walk": float(np.abs(Fs - want).max()),
            "max_rel_msd_minus_3_sigma2_tau": float(np.abs(msd[1:] / (3 * SIGMA ** 2 * tau[1:]) - 1).max()),
            "vanhove_total_per_lag_equals_N_times_buffers": bool((r["avDrTable"].sum(axis=0) == n * cnt).all()),
            "kernel": eng.kernel_config.split(";")[0],
        }
        print(json.dumps(line), flush=True)
        if args.dump:
            np.savez_compressed(args.dump, **{k: v for k, v in r.items() if isinstance(v, np.ndarray)}, countCorrAv=cnt)
    eng.close()
    src.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
