"""CPU, world_size 2, gloo: the N>1 plumbing — contiguous sharding of sequences over ranks and the single final
all-gather of per-track poses — gives exactly the single-process result.  The per-rank "tracker" here is the CPU
oracle (a checker standing in for the GPU so that the test runs without one)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _work(rank, world, port, nseq, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from invcompcamtrack_b200.shard import shard_range, gather_results
    from oracle import oracle as O
    from helpers import make_case, oracle_run
    orc = O.OracleLib()
    lo, hi = shard_range(nseq, rank, world)
    poses, iters = [], []
    for s in range(lo, hi):                     # one "sequence" = one frame pair with 3 tracks
        case = make_case(seed=100 + s, w=160, h=120, lv_f=1, npts=20, ntracks=3)
        r = oracle_run(orc, case, trace_cap=0, nthreads=1)
        poses.append(r["p_out"]); iters.append(r["iters"])
    p = torch.from_numpy(np.concatenate(poses)) if poses else torch.zeros(0, 6, dtype=torch.float64)
    i = torch.from_numpy(np.concatenate(iters)) if iters else torch.zeros(0, 2, dtype=torch.int32)
    gp, gi = gather_results(p, i)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), p=gp.numpy(), i=gi.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from invcompcamtrack_b200.shard import shard_range
    for n in (0, 1, 5, 8, 256):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


@pytest.mark.parametrize("nseq", [4, 5])
def test_two_ranks_equal_one_process(tmp_path, nseq):
    port = _free_port()
    mp.spawn(_work, args=(2, port, nseq, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import oracle as O
    from helpers import make_case, oracle_run
    orc = O.OracleLib()
    ref_p, ref_i = [], []
    for s in range(nseq):
        r = oracle_run(orc, make_case(seed=100 + s, w=160, h=120, lv_f=1, npts=20, ntracks=3), trace_cap=0, nthreads=1)
        ref_p.append(r["p_out"]); ref_i.append(r["iters"])
    assert np.array_equal(got["p"], np.concatenate(ref_p))
    assert np.array_equal(got["i"], np.concatenate(ref_i))
