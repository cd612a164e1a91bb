"""Multi-rank host logic on CPU (gloo, world_size 2): work partition with atomic families and the final
gather of packed knockoff rows.  The per-rank rows are produced by the CPU oracle here (no GPU in this
tier of the test-suite); the keyed generator makes every row independent of the sharding."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

import kgw_oracle as ko
from kgw_testutil import load_golden
from knockoffgwas_b200.shard import FAMILY_COST, partition_work, rows_of

REPO = Path(__file__).resolve().parents[1]


def test_partition_work_properties():
    rng = np.random.default_rng(0)
    N = 5000
    fams = []
    pool = rng.permutation(N)[:300]
    t = 0
    while t < 300:
        n = int(rng.integers(2, 6))
        fams.append(np.sort(pool[t:t + n]).astype(np.int32))
        t += n
    related = np.concatenate(fams)
    unrelated = np.setdiff1d(np.arange(N), related).astype(np.int32)
    for world in (1, 2, 3, 8):
        parts = partition_work(unrelated, fams, world)
        assert len(parts) == world
        got_u = np.concatenate([u for u, _ in parts])
        assert np.array_equal(got_u, unrelated)                       # contiguous slices, nothing lost or doubled
        got_f = sorted(f for _, fl in parts for f in fl)
        assert got_f == list(range(len(fams)))                        # every family on exactly one rank
        cost = [len(u) + FAMILY_COST * sum(len(fams[f]) for f in fl) for u, fl in parts]
        assert max(cost) - min(cost) <= FAMILY_COST * 5 + 1                   # balanced to within one family
        rows = [rows_of(u, fams, fl) for u, fl in parts]
        assert np.array_equal(np.sort(np.concatenate(rows)), np.arange(N))


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist

    sys.path[:0] = [str(REPO), str(REPO / "oracle"), str(REPO / "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from knockoffgwas_b200.shard import gather_rows

    prob, f = load_golden("toy_chr22_K10_res4")
    fams = [m for m, _ in prob.fam_clusters]
    u_r, f_r = partition_work(prob.unrelated, fams, world)[rank]
    hk = np.zeros_like(prob.H)
    res = ko.run_unrelated(prob, haps=u_r, mode="philox")
    hk[u_r] = res.hk
    rng, _ = ko.make_rng(prob, mode="philox")
    for fi in f_r:
        members, segs = prob.fam_clusters[fi]
        hk[members] = ko.run_family(prob, members, segs, rng).hk
    full = gather_rows(torch.from_numpy(hk), rows_of(u_r, fams, f_r), prob.N)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.slow
def test_two_rank_gather_equals_single_process(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "full.npy"
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    got = np.load(out)
    prob, _ = load_golden("toy_chr22_K10_res4")
    want = np.zeros_like(prob.H)
    want[prob.unrelated] = ko.run_unrelated(prob, mode="philox").hk
    rng, _ = ko.make_rng(prob, mode="philox")
    for members, segs in prob.fam_clusters:
        want[members] = ko.run_family(prob, members, segs, rng).hk
    assert np.array_equal(got, want)
