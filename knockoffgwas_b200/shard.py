"""Multi-GPU sharding of Knockoffs::generate(): haplotypes are independent given H, the references and
the HMM tables (knockoffs.cpp:654-692), IBD families are independent of each other (:440-445).  One
process per GPU takes a contiguous range of the sorted unrelated list plus whole families; there is no
collective on the data path, only a final gather of packed knockoff rows (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  SURVEY.md section 8e."""
from __future__ import annotations

import numpy as np


FAMILY_COST = 7.0   # measured on B200: a related haplotype costs about 7 unrelated ones (the CPU reference: 2-3)


def partition_work(unrelated, families, world_size: int, family_cost: float = FAMILY_COST):
    """Split the work list over `world_size` ranks.

    unrelated: sorted haplotype ids; families: list of member-id arrays (atomic units, a member costs
    `family_cost` unrelated haplotypes).
    Returns a list of (unrelated ids of the rank [contiguous slice], family indices of the rank).
    Families go to the least loaded rank in descending size (the reference's own order,
    knockoffs.cpp:364-372, made cost-aware); the unrelated list is then cut so that total costs even out."""
    from . import api
    unrelated = np.asarray(unrelated, dtype=np.int32)
    if family_cost != FAMILY_COST:
        raise ValueError("the family cost is a constant of the library (kgw_partition_work)")
    cut, dev = api.partition_work(len(unrelated), [len(f) for f in families], world_size)
    return [(unrelated[cut[r]:cut[r + 1]], [f for f in range(len(families)) if dev[f] == r]) for r in range(world_size)]


def rows_of(unrelated_r, families, fam_idx_r):
    """Haplotype rows a rank writes: its unrelated slice plus the members of its families (sorted)."""
    parts = [np.asarray(unrelated_r, dtype=np.int32)] + [np.asarray(families[f], dtype=np.int32) for f in fam_idx_r]
    return np.unique(np.concatenate(parts)) if parts else np.zeros(0, np.int32)


def gather_rows(hk_local, rows, n_total: int, group=None):
    """All-gather the packed knockoff rows every rank computed into the full [n_total][row_bytes] matrix.
    hk_local: torch uint8 tensor [n_total][row_bytes] (device tensor under NCCL, CPU tensor under gloo) in
    which this rank filled `rows`; returns the same-shaped tensor with every rank's rows filled."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rows_t = torch.as_tensor(np.asarray(rows, dtype=np.int64), device=hk_local.device)
    counts = [torch.zeros(1, dtype=torch.int64, device=hk_local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(rows)], dtype=torch.int64, device=hk_local.device), group=group)
    cap = int(max(int(c.item()) for c in counts))
    row_bytes = hk_local.shape[1]
    send_idx = torch.full((cap,), -1, dtype=torch.int64, device=hk_local.device)
    send_idx[:len(rows)] = rows_t
    send_dat = torch.zeros((cap, row_bytes), dtype=torch.uint8, device=hk_local.device)
    send_dat[:len(rows)] = hk_local[rows_t]
    all_idx = torch.empty((world * cap,), dtype=torch.int64, device=hk_local.device)
    all_dat = torch.empty((world * cap, row_bytes), dtype=torch.uint8, device=hk_local.device)
    dist.all_gather_into_tensor(all_idx, send_idx, group=group)
    dist.all_gather_into_tensor(all_dat, send_dat, group=group)
    out = hk_local.clone()
    keep = all_idx.reshape(-1) >= 0
    out[all_idx.reshape(-1)[keep]] = all_dat.reshape(-1, row_bytes)[keep]
    assert out.shape[0] == n_total
    return out
