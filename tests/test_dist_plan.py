"""Host-side logic of the multi-GPU path (mitoflex_b200/dist.py), on CPU: owner assignment, split computation, chunk
tables, and the exchange itself over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest

from mitoflex_b200 import dist as mdist


def test_assign_owners_balanced_and_monotonic():
    rng = np.random.default_rng(0)
    h = rng.integers(0, 1000, 1024)
    h[:100] *= 5     # canonical keys pile up at small prefixes
    for world in (1, 2, 3, 4, 8):
        b = mdist.assign_owners(h, world)
        assert b[0] == 0 and b[-1] == 1024 and (np.diff(b) >= 0).all()
        loads = [h[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(loads) <= h.sum() / world + h.max() + 1
    # degenerate: everything in one bin, empty histogram
    one = np.zeros(1024, np.int64)
    one[7] = 10
    b = mdist.assign_owners(one, 4)
    assert b[0] == 0 and b[-1] == 1024 and (np.diff(b) >= 0).all()
    b = mdist.assign_owners(np.zeros(1024, np.int64), 4)
    assert b[-1] == 1024 and (np.diff(b) >= 0).all()


def test_exchange_plan_accounts_for_every_record():
    rng = np.random.default_rng(1)
    H = rng.integers(0, 50, (4, 1024))
    H[:, 300:400] = 0
    plans = [mdist.exchange_plan(H, r) for r in range(4)]
    for r, p in enumerate(plans):
        assert p["send"].sum() == H[r].sum()
        assert p["n_recv"] == H[:, p["lo"]:p["hi"]].sum() == p["chunk_size"].sum()
        for s in range(4):
            assert p["recv"][s] == plans[s]["send"][r]
        # chunks tile the receive buffer in order, without gaps
        assert (p["chunk_start"] == np.concatenate([[0], np.cumsum(p["chunk_size"])[:-1]])).all()
        assert (p["chunk_seg"] >= 0).all() and (p["chunk_seg"] < p["n_segs"]).all()


def test_peer_bin_bases_match_the_chunk_table():
    """fused exchange: the address a source writes bin b to must be where the owner's chunk table expects it"""
    rng = np.random.default_rng(3)
    H = rng.integers(0, 40, (3, 1024))
    H[:, 50:80] = 0
    ptrs = [10 ** 12, 2 * 10 ** 12, 3 * 10 ** 12]
    for me in range(3):
        plan = mdist.exchange_plan(H, me)
        b = mdist.peer_bin_bases(H, plan["bounds"], me, ptrs, 8)
        for r in range(3):
            pr = mdist.exchange_plan(H, r)
            acc = sum(int(H[s, pr["lo"]:pr["hi"]].sum()) for s in range(me))
            for bb in range(pr["lo"], pr["hi"]):
                if H[me, bb] > 0:
                    assert int(b[bb]) == ptrs[r] + acc * 8
                acc += int(H[me, bb])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        n = 5000 + 700 * rank
        # skewed 44-bit keys in two 32-bit words, like canonical (k+1)-mers for k=21 (min of two uniform draws)
        a = rng.integers(0, 1 << 44, n, dtype=np.uint64)
        b = rng.integers(0, 1 << 44, n, dtype=np.uint64)
        keys = np.minimum(a, b) << np.uint64(20)
        bins = (keys >> np.uint64(54)).astype(np.int64)
        hist = np.bincount(bins, minlength=1024).astype(np.int64)
        allh = [torch.zeros(1024, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allh, torch.from_numpy(hist))
        H = np.stack([t.numpy() for t in allh])
        plan = mdist.exchange_plan(H, rank)
        order = np.argsort(bins, kind="stable")          # what the device-side partition produces: bins ascending
        rows = np.stack([(keys[order] >> np.uint64(32)).astype(np.uint32), (keys[order] & np.uint64(0xffffffff)).astype(np.uint32)],
                        axis=1).view(np.int32)
        recv = mdist.all_to_all_records(torch.from_numpy(np.ascontiguousarray(rows)), plan["send"], plan["recv"]).numpy().view(np.uint32)
        rkeys = (recv[:, 0].astype(np.uint64) << np.uint64(32)) | recv[:, 1].astype(np.uint64)
        rbins = (rkeys >> np.uint64(54)).astype(np.int64)
        assert len(rkeys) == plan["n_recv"]
        assert ((rbins >= plan["lo"]) & (rbins < plan["hi"])).all()
        for st, sz, sg in zip(plan["chunk_start"], plan["chunk_size"], plan["chunk_seg"]):
            assert (rbins[st:st + sz] == plan["lo"] + sg).all()
        np.save(os.path.join(out_dir, f"in{rank}.npy"), keys)
        np.save(os.path.join(out_dir, f"out{rank}.npy"), rkeys)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_exchange_over_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ins = np.sort(np.concatenate([np.load(tmp_path / f"in{r}.npy") for r in range(world)]))
    outs = [np.load(tmp_path / f"out{r}.npy") for r in range(world)]
    # rank order == key-range order: sorting each rank's share and concatenating gives the global sort
    glob = np.concatenate([np.sort(o) for o in outs])
    assert np.array_equal(glob, ins)


@pytest.mark.parametrize("world,nsl", [(1, 1), (2, 8), (3, 8), (8, 1024), (4, 2)])
def test_ks_exchange_plan_tiles_every_owner_buffer(world, nsl):
    """the multi-GPU item filter (dist.ks_exchange_plan): every source's bins land back to back in the owner's buffer -- inserts
    first, then queries, slice-major and source-minor -- with no gap and no overlap"""
    from mitoflex_b200 import dist as mdist
    rng = np.random.default_rng(world * 1000 + nsl)
    H = rng.integers(0, 50, size=(world, 2 * nsl)).astype(np.int64)
    H[rng.random(H.shape) < 0.2] = 0
    plans = [mdist.ks_exchange_plan(H, r) for r in range(world)]
    b = mdist.ks_slice_bounds(nsl, world)
    assert b[0] == 0 and b[-1] == nsl and (np.diff(b) >= 0).all()
    for owner in range(world):
        n = int(plans[owner]["recv_records"][owner])
        assert n == plans[owner]["n_ins"] + plans[owner]["n_qry"]
        tag = np.full((n, 3), -1, dtype=np.int64)           # (kind, slice, source) of the record stored at every position
        for src in range(world):
            p = plans[src]
            for bin_ in np.nonzero(p["owner"] == owner)[0]:
                cnt, off = int(H[src, bin_]), int(p["offset"][bin_])
                assert (tag[off:off + cnt, 0] == -1).all(), "two bins overlap"
                tag[off:off + cnt] = (bin_ // nsl, bin_ % nsl, src)
        assert (tag[:, 0] >= 0).all(), "a gap in the receive buffer"
        assert (tag[:plans[owner]["n_ins"], 0] == 0).all() and (tag[plans[owner]["n_ins"]:, 0] == 1).all()
        for kind_rows in (tag[:plans[owner]["n_ins"]], tag[plans[owner]["n_ins"]:]):
            key = kind_rows[:, 1] * world + kind_rows[:, 2]
            assert (np.diff(key) >= 0).all(), "not slice-major / source-minor"
            assert ((kind_rows[:, 1] >= b[owner]) & (kind_rows[:, 1] < b[owner + 1])).all()
        assert plans[owner]["slice_lo"] == b[owner] and plans[owner]["n_owned"] == b[owner + 1] - b[owner]
