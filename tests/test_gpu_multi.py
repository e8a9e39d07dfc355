"""Multi-GPU (needs >= 2 CUDA devices; run with `gpurun --gpus 2`): one process per GPU, NCCL, particle-sharded.
The sharded result must equal the single-GPU result and the oracle (GPU-count independence)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path, S, T, B, ncell, frames, flush_limit):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from lmp2sqt_b200 import Deck, comm_unique_id, run_trajectories, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ids = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    traj = synth.make_trajectory(n_cells=ncell, n_frames=frames, seed=77, box_jitter=0.005)
    deck = Deck(limitCorrAv=flush_limit, nCBuffer=B, nCSpatial=S, nCTime=T)
    res, eng = run_trajectories([traj], deck, "Ft", device=rank, rank=rank, nranks=world, unique_id=ids[0])
    d = res.to_dict()
    np.savez(out_path % rank, **{k: v for k, v in d.items()})
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("S,T,B", [(6, 8, 4), (43, 6, 6), (100, 6, 6)])  # the last one runs the tensor-core kernels
def test_two_gpu_shards_match_single_gpu_and_oracle(tmp_path, S, T, B):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    from lmp2sqt_b200 import Deck, run_trajectories, synth
    from parity import check_raw_sums
    ncell, frames = 3, 40
    out = str(tmp_path / "rank%d.npz")
    mp.spawn(_worker, args=(2, _free_port(), out, S, T, B, ncell, frames, 10**6), nprocs=2, join=True)
    r0, r1 = dict(np.load(out % 0)), dict(np.load(out % 1))
    traj = synth.make_trajectory(n_cells=ncell, n_frames=frames, seed=77, box_jitter=0.005)
    _, raw, _ = pyoracle.run_port(traj, S, T, B, limitCorrAv=10**6, finish=False)
    for r in (r0, r1):
        r = {k: (v.item() if v.shape == () else v) for k, v in r.items()}
        check_raw_sums(r, raw, traj.n_atoms, label="2 GPUs ")
    for k in r0:                                   # every rank holds identical, fully reduced results
        assert np.array_equal(r0[k], r1[k]), k
    single, eng = run_trajectories([traj], Deck(limitCorrAv=10**6, nCBuffer=B, nCSpatial=S, nCTime=T), "Ft")
    s = single.to_dict()
    for k in ("avF_qq2", "avF_s_qq2", "rrMSDAv", "rrMQDAv"):
        np.testing.assert_allclose(r0[k], s[k], rtol=1e-10, atol=1e-9)
    assert np.array_equal(r0["avDrTable"], s["avDrTable"])


# ---- G GPUs behind one caller thread (include/sqt.h sqt_multi_*, CLI -g N): SURVEY.md section 8e frame fan-out ------------

@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("S,T,B,flags", [(6, 8, 4, "Ft"), (100, 6, 6, "Ft"), (30, 6, 6, "FtV"), (43, 6, 6, "Ft"), (30, 6, 6, "FtVMs"), (70, 6, 6, "FtMs")])
def test_in_process_fanout_matches_oracle_and_single_gpu(S, T, B, flags):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    from lmp2sqt_b200 import Deck, run_trajectories, synth
    from parity import check_raw_sums
    traj = synth.make_trajectory(n_cells=3, n_atoms=101, n_frames=40, seed=78, box_jitter=0.005)
    deck = Deck(limitCorrAv=10**6, nCBuffer=B, nCSpatial=S, nCTime=T)
    res, eng = run_trajectories([traj], deck, flags, n_gpus=2)
    assert eng.n_gpus == 2
    got = res.to_dict()
    _, raw, _ = pyoracle.run_port(traj, S, T, B, limitCorrAv=10**6, flags=flags, finish=False)
    check_raw_sums(got, raw, traj.n_atoms, label="fan-out x2 ")
    if "V" in flags:
        for k in ("avC2_v_rho", "avC2_rho_v", "avC2_v_v", "rrCvvAv", "rrCvcmvcmAv"):
            np.testing.assert_allclose(got[k], raw[k], rtol=1e-6, atol=1e-9 * np.abs(raw[k]).max(), err_msg=k)
    if "M" in flags or "s" in flags:                  # the sums the shards exchange for -M / -s (kmu_q1, sum v (x) r, per-axis MSD)
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from test_gpu_fields import _check_fields
        _check_fields(got, raw, flags, "fan-out x2 " + flags)
    single, e1 = run_trajectories([traj], deck, flags)
    s = single.to_dict()
    assert np.array_equal(got["avDrTable"], s["avDrTable"])
    for k in ("avF_qq2", "avF_s_qq2", "rrMSDAv", "rrMQDAv"):
        np.testing.assert_allclose(got[k], s[k], rtol=1e-10, atol=1e-9)
    eng.close(); e1.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("name", ["g1_midrun_print", "g3_two_files_skip", "g6_nCS70", "g7_nCS128"])
def test_cli_with_two_gpus_is_still_a_drop_in(name, tmp_path):
    """`lmp2sqt_b200.out -g 2`: one reader thread, one pinned frame, two chunk copies -- the same files as the reference's."""
    import shutil
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from golden_utils import case_dir, load_case, unpack_dumps
    from test_gpu_cli import _run_and_compare
    c, g = load_case(name), case_dir(name)
    shutil.copy(os.path.join(g, "in.lmp2sqt"), tmp_path / "in.lmp2sqt")
    dumps = unpack_dumps(name, str(tmp_path))
    _run_and_compare(c, g, dumps, tmp_path, ["-g", "2"], expect_in_stderr="2 GPUs, particle-sharded")
