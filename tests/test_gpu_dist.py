"""Multi-GPU parity (needs >= 2 GPUs): reads sharded over 2 ranks, keys and items routed by prefix with NCCL all-to-all;
the ranks' sdbg pieces concatenated in rank order must equal the oracle's graph of the whole read set, bit for bit."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, k, m, seed, mode):
    import sys
    import torch
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    from gpu_common import make_reads
    from mitoflex_b200 import dist as mdist
    from mitoflex_b200 import lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        bases, starts = make_reads(seed, 60000, k, genome_len=150000, max_len=150, err=0.005)
        n = len(starts) - 1
        lo, hi = n * rank // world, n * (rank + 1) // world
        sb = bases[starts[lo]:starts[hi]]
        ss = starts[lo:hi + 1] - starts[lo]
        ctx = lib.Context(rank)
        # "p2p" takes the super-k-mer count when the library supports k (16..26); "p2p-prefix" forces the prefix-bin key exchange
        # "-nofilter": all 6 items per edge instead of the item filter across GPUs (k <= 31)
        parts_ = mode.split("-")
        runner = mdist.DistRead2Sdbg(ctx, k, m, exchange=parts_[0], skm=False if "prefix" in parts_ else None,
                                     item_filter=False if "nofilter" in parts_ else None)
        assert runner.skm == (mode.startswith("p2p") and "prefix" not in parts_ and 16 <= k <= 26)
        assert runner.filter == (parts_[0] == "p2p" and "nofilter" not in parts_ and 16 <= k <= 31)
        res = runner.run(ctx.upload_reads(sb, ss))
        g = res.sdbg.to_numpy()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), w=g["w"], last=g["last"], tip=g["tip"], mul=g["mul"],
                 tip_labels=g["tip_labels"])
        dist.barrier()
        runner.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,m,mode", [(21, 2, "p2p"), (21, 2, "p2p-prefix"), (26, 1, "p2p"), (31, 1, "p2p"), (21, 2, "nccl"), (47, 2, "p2p"),
                                      (21, 2, "p2p-nofilter"), (24, 1, "p2p-prefix-nofilter")])
def test_two_gpu_read2sdbg_matches_oracle(oracle, tmp_path, k, m, mode):
    _run_dist_case(oracle, tmp_path, k, m, mode, 2)


@pytest.mark.parametrize("k,m,mode", [(21, 2, "p2p"), (26, 1, "p2p"), (31, 2, "p2p"), (21, 1, "p2p-nofilter")])
def test_one_rank_group_read2sdbg_matches_oracle(oracle, tmp_path, k, m, mode):
    """the multi-GPU driver with a process group of ONE rank: the super-k-mer count, the item filter with its hash slices and the
    item exchange all run (every peer is the rank itself), so a single-GPU box exercises the whole of dist.py"""
    _run_dist_case(oracle, tmp_path, k, m, mode, 1)


def _run_dist_case(oracle, tmp_path, k, m, mode, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from gpu_common import make_reads
    seed = 4242 + k
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), k, m, seed, mode), nprocs=world, join=True)
    bases, starts = make_reads(seed, 60000, k, genome_len=150000, max_len=150, err=0.005)
    g = oracle.read2sdbg(oracle.Reads(bases, starts), k, m, threads=8)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for f in ("w", "last", "tip", "mul"):
        assert np.array_equal(np.concatenate([p[f] for p in parts]), getattr(g, f)), f
    assert np.array_equal(np.concatenate([p["tip_labels"] for p in parts]), g.tip_labels)
