#!/usr/bin/env python
"""bench.py -- knockoff haplotype-SNPs/s of the snpknock2 sampler hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the unmodified reference on the host CPUs

Headline workload (BASELINE.json configs[1], "C2"): synthetic chr22-shaped data, 20 000 haplotypes x 10 000 SNPs per
GPU, K = 50 references per haplotype, one partition resolution (singleton groups, the reference's most expensive
resolution).  A "step" is one Knockoffs::generate() call (snpknock2/src/knockoffs.cpp:311) for one resolution.

`value`  : CUDA events on the launching stream around the K steps; every input resident in HBM; each step = the
           sampler kernel + the copy of the packed knockoff rows into pinned host memory ("D2H / gather
           of Hk included"); `value_kernel_only` is the same without the copy.
`e2e`    : the one-shot C-ABI call kgw_generate() with pinned HOST buffers (plan create, H2D, tables, kernel, D2H,
           destroy inside the timed region).
N > 1    : (torchrun, one rank per GPU, weak scaling) ONE population of N x 20 000 haplotypes; every rank holds the
           whole H and the references, takes the shard kgw_partition_work gives it (no collective on the data path)
           and sends its packed rows to rank 0 over NCCL (gather, on a side stream overlapped with the next step's
           kernel, the last one inside the timed region); after the timed region rank 0 re-computes sampled rows of
           every other rank's shard on its own GPU and requires them to equal the gathered rows.
Beside the headline the N = 1 line carries one leg per north-star configuration (BASELINE.json configs[2..4]):
`c3_shard` (one GPU's eighth of 700k x 60k, K=50), `c4_resolutions` (six partition levels through
kgw_plan_set_groups on one resident plan) and `c5_shard` (one GPU's eighth of configs[4]: K=100, 60k SNPs, 10 % related haplotypes in the same
call), each with its own roofline dict and clocks, and the rows either side of the path.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "knockoff_haplotype_snps_per_sec_K50"
UNIT = "hap-SNPs/s"
RHO, LAMBDA = 1.0, 1e-3  # module_2_knockoffs.sh values


def b_alg(K: int) -> float:
    """Algorithmic bytes per haplotype-SNP (SURVEY.md 8d contract figure): one fp64 write + one
    fp64 read of the K backward messages, K reference bits + 1 target bit read, 1 knockoff bit written."""
    return 16.0 * K + (K + 2) / 8.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--haps", type=int, default=20000, help="haplotypes per GPU")
    ap.add_argument("--snps", type=int, default=10000)
    ap.add_argument("--K", type=int, default=50)
    ap.add_argument("--mean-group", type=float, default=1.0, help="mean variant-group size (1 = singletons)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--block-loci", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--related-pairs", type=int, default=2368, help="synthetic IBD pairs of the related-branch leg (0 = skip)")
    ap.add_argument("--legs", default="c3,c4,c5,related,neighbours,cpu",
                    help="comma list of the extra legs of the N = 1 line (c3, c4, c5, related, neighbours, cpu); '' = headline only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--ref-sample", type=int, default=0,
                    help="--impl reference: haplotypes of the workload the reference is timed on (0 = all of them)")
    ap.add_argument("--skip-sweeps", type=int, default=0, help="development only: timing experiments, results invalid")
    return ap.parse_args()


def workload_config(a, world):
    return {
        "workload": f"synthetic chr22-shaped: {a.haps} haplotypes x {a.snps} SNPs per GPU, K={a.K}, "
                    f"one partition resolution (mean group size {a.mean_group:g}), unrelated haplotypes",
        "haplotypes_per_gpu": a.haps, "snps": a.snps, "K": a.K, "mean_group_size": a.mean_group,
        "haplotypes_total": a.haps * world, "hmm_rho": RHO, "hmm_lambda": LAMBDA,
        "sharding": ("one population of %d haplotypes, H and references on every rank, work list cut by kgw_partition_work "
                     "(contiguous slices of the sorted unrelated list), no data-path collective; packed Hk rows of step i sent "
                     "to rank 0 (NCCL gather) on a side stream, overlapped with the kernel of step i+1, the last one inside "
                     "the timed region" % (a.haps * world)) if world > 1 else "single GPU",
        "l2": "256 MiB buffer rewritten between timed steps (L2 flush)",
    }


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and clock-event reasons during the timed region (NVML, 20 ms period)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(device_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.ok:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}



# --------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(a, steps: int, warmup: int, threads: int, n_sample: int):
    """Times the reference's own CPU implementation of the path (all SNPs; with n_sample < a.haps on a haplotype
    sample whose references are re-drawn inside it -- cost is linear in haplotypes).
    Uses the unmodified reference in oracle/_ref/ when present (kind "reference": snpknock2_timer = the reference's
    objects with a std::chrono timer around Knockoffs::generate and nothing else), else the C restatement in
    oracle/ (kind "port").  Returns (per-step seconds list, sample haps, kind, threads)."""
    sys.path.insert(0, str(REPO / "oracle"))
    from knockoffgwas_b200.synth import make_problem_fast
    threads = threads or len(os.sched_getaffinity(0))
    n_s = max(2 * (a.K + 2), min(n_sample, a.haps))
    n_s -= n_s % 2
    threads = min(threads, n_s)
    prob = make_problem_fast(n_s, a.snps, a.K, seed=2020, mean_group=a.mean_group, rho=RHO, lam=LAMBDA, block=min(5000, n_s))
    import refio
    if refio.available():
        import tempfile
        # one process, (warmup + steps) identical partition columns => that many timed generate() calls
        with tempfile.TemporaryDirectory(prefix="kgw_ref_") as tmp:
            tmp = Path(tmp)
            files = refio.write_inputs(tmp / "in", prob.H, prob.p, prob.ref, prob.groups, prob.cm, prob.bp)
            ids = [f"rs{j}" for j in range(prob.p)]
            ncol = warmup + steps
            with open(files["part"], "w") as f:
                f.write("SNP " + " ".join(f"res_{c}" for c in range(ncol)) + "\n")
                f.writelines(f"{ids[j]} " + " ".join([str(int(prob.groups[j]))] * ncol) + "\n" for j in range(prob.p))
            secs = refio.run_reference_all_resolutions(files, tmp / "out", rho=RHO, lam=LAMBDA, n_threads=threads)
        assert len(secs) == ncol, (len(secs), ncol)
        return secs[warmup:], n_s, "reference", threads
    # port: oracle restatement, haplotype ranges on Python threads (ctypes releases the GIL)
    import kgw_oracle as ko
    op = ko.Problem(H=prob.H, p=prob.p, K=prob.K, ref_local=prob.ref[:, None, :], ref_global=prob.ref, b=prob.b,
                    mu=prob.mu, groups=prob.groups, seed=prob.seed)
    bounds = np.linspace(0, n_s, threads + 1).astype(int)
    secs = []
    for _ in range(warmup + steps):
        t0 = time.perf_counter()
        thr = [threading.Thread(target=ko.run_unrelated, args=(op, np.arange(bounds[w], bounds[w + 1], dtype=np.int32)))
               for w in range(threads)]
        [t.start() for t in thr]
        [t.join() for t in thr]
        secs.append(time.perf_counter() - t0)
    return secs[warmup:], n_s, "port", threads


def cpu_baseline_dict(a, steps, warmup, threads, n_sample):
    secs, n_s, kind, thr = cpu_reference_run(a, steps, warmup, threads, n_sample)
    v = n_s * a.snps * len(secs) / sum(secs)
    what = "all" if n_s >= a.haps else f"{n_s} of"
    return {"value": v, "unit": UNIT, "cores": thr, "kind": kind,
            "sample": f"{what} {a.haps} haplotypes x all {a.snps} SNPs, K={a.K}"
                      + ("" if n_s >= a.haps else ", references re-drawn inside the sample")
                      + f"; {len(secs)} timed Knockoffs::generate() calls after {warmup} warm-up, --n_threads {thr}"}, secs, n_s


def run_reference_arm(a, rank, world):
    """The reference arm: the whole C2 workload (all 20 000 haplotypes) on every host thread; about 16 s per step on
    16 threads, so the driver's K + W steps end within a few minutes (--ref-sample bounds it further)."""
    if rank != 0:
        return
    d, secs, n_s = cpu_baseline_dict(a, a.steps, a.warmup, a.cpu_threads, a.ref_sample or a.haps)
    line = {"impl": "reference", "metric": METRIC, "value": d["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a, 1),
            "cpu_baseline": d,
            "e2e": {"value": d["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- GPU arm
class _DevBuf:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


def hbm_peak():
    peaks_file = REPO / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        return float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def roofline_dict(units, K, kernel_ms, kernel, note, traffic=None):
    """Contract roofline of one launch: algorithmic bytes (SURVEY.md 8d per-unit figure x units) over the live
    CUDA-event duration of the kernel, against the measured HBM peak."""
    peak, peak_src = hbm_peak()
    achieved = units * b_alg(K) / (kernel_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "kernel": kernel, "kernel_ms": kernel_ms, "algorithmic_bytes_per_hap_snp": b_alg(K), "hap_snps_per_launch": units,
            "peak_source": peak_src, "note": note}


SAMPLER_NOTE = ("contract bytes assume the reference's materialised fp64 backward table; this kernel stores one message row "
                "every block_loci loci plus four partial sums per lane and locus and rebuilds single entries on demand, so "
                "its DRAM traffic is below the algorithmic bytes (see DESIGN.md)")


def timed_runs(torch, plan, stream, dev, steps, warmup, flush, after_run=None):
    """`steps` timed plan.run() calls on `stream` after `warmup` untimed ones, L2 flushed before each; returns
    (total ms of the timed region, per-step kernel ms, clocks).  after_run(i) is enqueued after every run (inside the
    timed region)."""
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    with torch.cuda.stream(stream):
        for _ in range(max(warmup, 0)):
            flush.zero_()
            plan.run()
            if after_run:
                after_run(-1)
        torch.cuda.synchronize(dev)
        sampler = ClockSampler(dev.index)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            flush.zero_()
            ev[i][0].record(stream)
            plan.run()
            ev[i][1].record(stream)
            if after_run:
                after_run(i)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        clocks = sampler.stop()
    return float(e0.elapsed_time(e1)), [ev[i][0].elapsed_time(ev[i][1]) for i in range(steps)], clocks


def leg_c3_shard(torch, kg, synth, dev, stream, flush, a):
    """One GPU's eighth of BASELINE.json configs[2] (700k haplotypes x 60k SNPs, K=50): the configuration the
    north-star target (>= 50 % of the HBM roofline per GPU) is quoted on."""
    N, p, K = 87500, 60000, 50
    prob = synth.make_problem_fast(N, p, K, seed=2021, rho=RHO, lam=LAMBDA)
    opts = kg.Options(device=dev.index, stream=stream.cuda_stream, lanes_per_hap=a.lanes, block_loci=a.block_loci)
    plan = kg.Plan(prob.H, p, K, prob.ref[:, None, :], prob.b, prob.mu, prob.groups, seed=prob.seed, options=opts)
    info = plan.info()
    total_ms, kern, clocks = timed_runs(torch, plan, stream, dev, 3, 1, flush)
    # the row after the path at this size: .bed body of (H, Hk), 2 x 656 MB in, 1.3 GB out
    bed = None
    try:
        body, _ = plan.encode_bed()
        body, bed_ms = plan.encode_bed(out=body)
        rb = (p + 7) // 8
        bed = {"kernel_ms": bed_ms, "input_bytes": int(2 * N * rb), "output_bytes": int(body.nbytes),
               "kernel_GBps": (body.nbytes + 2 * N * rb) / (bed_ms * 1e-3) / 1e9, "frac_of_hbm_peak": (body.nbytes + 2 * N * rb) / (bed_ms * 1e-3) / 1e9 / hbm_peak()[0]}
        del body
    except Exception as e:
        bed = {"error": repr(e)[:200]}
    plan.close()
    kg.release_workspace()
    k_ms = float(np.mean(kern))
    return {"workload": f"{N} haplotypes x {p} SNPs, K={K}, singleton groups, unrelated (1/8 of 700k x 60k)", "steps": 3, "warmup": 1, "bed_writer": bed,
            "ms_per_step": total_ms / 3, "value": N * p * 3 / (total_ms * 1e-3), "unit": UNIT, "clocks": clocks,
            "roofline": roofline_dict(N * p, K, k_ms, "kgw::sampler_kernel", SAMPLER_NOTE),
            "kernel": {k: info[k] for k in ("grid", "block", "lanes_per_hap", "states_per_lane", "scratch_bytes", "block_loci", "segment_loci")}}


def leg_c4_resolutions(torch, kg, synth, plan, prob, dev, stream, flush):
    """BASELINE.json configs[3]: six partition levels (single SNP ... ~50 SNPs per group) on ONE resident plan: H, the
    references and the scratch stay in HBM, kgw_plan_set_groups replaces the per-locus knockoff tables."""
    grng = np.random.default_rng(4)
    levels = []
    for mean_group in (1.0, 2.0, 5.0, 10.0, 25.0, 50.0):
        groups = synth.make_groups(prob.p, mean_group, grng)
        t0 = time.perf_counter()
        plan.set_groups(groups)
        set_ms = (time.perf_counter() - t0) * 1e3
        total_ms, kern, clocks = timed_runs(torch, plan, stream, dev, 3, 1, flush)
        k_ms = float(np.mean(kern))
        levels.append({"mean_group_size": mean_group, "groups": int(groups[-1]) + 1, "set_groups_host_ms": set_ms, "kernel_ms": k_ms,
                       "hap_snps_per_s": prob.N * prob.p / (k_ms * 1e-3), "clocks": clocks,
                       "roofline": roofline_dict(prob.N * prob.p, prob.K, k_ms, "kgw::sampler_kernel", SAMPLER_NOTE)})
    plan.set_groups(prob.groups)
    return {"workload": f"{prob.N} haplotypes x {prob.p} SNPs, K={prob.K}: six resolutions through kgw_plan_set_groups on the "
                        f"headline's resident plan (per GPU share of a 200k-haplotype chr22)", "levels": levels}


def leg_c5_shard(torch, kg, synth, dev, stream, flush, a):
    """One GPU's eighth of BASELINE.json configs[4] (350k samples = 700k haplotypes, K=100, 10 % related haplotypes):
    78 750 unrelated haplotypes and 4 375 IBD pairs in ONE generate() call, 60k SNPs."""
    N, p, K, pairs = 87500, 60000, 100, 4375
    prob = synth.make_problem_fast(N, p, K, seed=2022, rho=RHO, lam=LAMBDA)
    synth.add_ibd_pairs(prob, pairs)
    related = np.concatenate([m for m, _ in prob.families])
    unrelated = np.setdiff1d(np.arange(N, dtype=np.int32), related).astype(np.int32)
    opts = kg.Options(device=dev.index, stream=stream.cuda_stream)
    out = {"workload": f"{len(unrelated)} unrelated + {len(related)} related haplotypes ({len(prob.families)} IBD pairs) x {p} SNPs, "
                       f"K={K}, singleton groups, one generate() call (one GPU's eighth of 700k haplotypes)"}
    plan = kg.Plan(prob.H, p, K, prob.ref[:, None, :], prob.b, prob.mu, prob.groups, unrelated=unrelated, ref_global=prob.ref,
                   seed=prob.seed, families=prob.families, options=opts)
    total_ms, kern, clocks = timed_runs(torch, plan, stream, dev, 2, 1, flush)
    plan.close()
    units = N * p
    out.update({"steps": 2, "warmup": 1, "ms_per_step": total_ms / 2, "value": units * 2 / (total_ms * 1e-3), "unit": UNIT, "clocks": clocks,
                "roofline": roofline_dict(units, K, float(np.mean(kern)), "kgw::sampler_kernel + kgw::family_kernel",
                                          "both kernels of the call; contract bytes of a related haplotype-SNP taken equal to an "
                                          "unrelated one (one fp64 K-vector written and read per locus), see related_branch")})
    # the two branches on their own, same population
    for name, kw in (("unrelated_only", dict(unrelated=unrelated)),
                     ("related_only", dict(unrelated=np.zeros(0, np.int32), ref_global=prob.ref, families=prob.families))):
        pl = kg.Plan(prob.H, p, K, prob.ref[:, None, :], prob.b, prob.mu, prob.groups, seed=prob.seed, options=opts, **kw)
        t_ms, kn, _ = timed_runs(torch, pl, stream, dev, 1, 1, flush)
        pl.close()
        n = len(unrelated) if name == "unrelated_only" else len(related)
        out[name] = {"ms": t_ms, "hap_snps_per_s": n * p / (t_ms * 1e-3)}
    kg.release_workspace()
    return out


def run_b200_arm(a, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import knockoffgwas_b200 as kg
    from knockoffgwas_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback)")
    # stdout carries the JSON line only: whatever the libraries print while the job runs (NCCL writes its version line
    # to stdout at communicator creation) is sent to stderr; the real stdout comes back for the final print
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    legs = set(x for x in a.legs.split(",") if x) if (world == 1 and not a.no_e2e) else set()
    if a.no_cpu_baseline:
        legs.discard("cpu")

    # ONE population for the whole job (every rank generates the same one from the same seeds)
    N_tot = a.haps * world
    prob = synth.make_problem_fast(N_tot, a.snps, a.K, seed=2020, mean_group=a.mean_group, rho=RHO, lam=LAMBDA)
    p, K = prob.p, prob.K
    cut, _ = kg.partition_work(N_tot, [], world)                     # contiguous slices of the sorted unrelated list
    mine = np.arange(cut[rank], cut[rank + 1], dtype=np.int32)
    units = len(mine) * p                                            # haplotype-SNPs per step on this rank
    rb = (p + 7) // 8

    stream = torch.cuda.Stream(device=dev)
    opts = kg.Options(device=local_rank, block_loci=a.block_loci, lanes_per_hap=a.lanes, ctas_per_sm=a.ctas_per_sm,
                      stream=stream.cuda_stream, skip_sweeps=a.skip_sweeps)
    plan = kg.Plan(prob.H, p, K, prob.ref[:, None, :], prob.b, prob.mu, prob.groups, seed=prob.seed, unrelated=mine, options=opts)
    info = plan.info()
    stride = info["row_stride_bytes"]
    hk_dev = torch.as_tensor(_DevBuf(plan.device_hk_ptr(), N_tot * stride), device=dev).view(N_tot, stride)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    host_out = torch.zeros((N_tot, rb), dtype=torch.uint8).pin_memory()    # the caller's Hk matrix
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    rows_max = int(max(cut[r + 1] - cut[r] for r in range(world)))
    staged = torch.zeros((rows_max, stride), dtype=torch.uint8, device=dev) if world > 1 else None
    gathered = [torch.zeros((rows_max, stride), dtype=torch.uint8, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    gather_done = [None]

    def after_run(i):
        if world == 1:
            # N = 1: the packed rows go to the caller's pinned host matrix (blocking call: pack kernel + one D2H)
            plan.download(out=host_out.numpy())
            return
        # N > 1: this rank's rows -> rank 0 over NCCL, from a staged copy, on a side stream (overlaps the next kernel)
        if gather_done[0] is not None:
            stream.wait_event(gather_done[0])          # the previous gather has read `staged`
        staged[:len(mine)].copy_(hk_dev[cut[rank]:cut[rank + 1]], non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(stream)
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            dist.gather(staged, gathered, dst=0)
            gather_done[0] = torch.cuda.Event()
            gather_done[0].record(comm)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- the timed region: W warm-up steps, barrier, K steps, max over ranks
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(a.steps)]
    with torch.cuda.stream(stream):
        for _ in range(max(a.warmup, 0)):
            flush.zero_()
            plan.run()
            after_run(-1)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        t_wall0 = time.perf_counter()
        ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_start.record(stream)
        for i in range(a.steps):
            flush.zero_()
            ev[i][0].record(stream)
            plan.run()
            ev[i][1].record(stream)
            after_run(i)
        if world > 1:
            stream.wait_event(gather_done[0])          # the last gather is part of the timed region
        ev_end.record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop()
    kern_ms = [ev[i][0].elapsed_time(ev[i][1]) for i in range(a.steps)]
    total_ms = float(ev_start.elapsed_time(ev_end))   # K steps incl. the L2 flushes and the D2H (N = 1) / every gather (N > 1)
    tt = torch.tensor([total_ms, float(np.sum(kern_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms_max, kern_sum_max = float(tt[0].item()), float(tt[1].item())
    value = a.haps * world * p * a.steps / (total_ms_max * 1e-3)
    value_kernel_only = a.haps * world * p * a.steps / (kern_sum_max * 1e-3)

    # ---- N > 1: the gathered matrix on rank 0 must be what one GPU computes (keyed generator => exact)
    check = None
    if world > 1:
        if rank == 0:
            full = torch.zeros((N_tot, stride), dtype=torch.uint8, device=dev)
            for r in range(world):
                full[cut[r]:cut[r + 1]] = gathered[r][:cut[r + 1] - cut[r]]
            full_h = full[:, :rb].cpu().numpy()
            rng = np.random.default_rng(1)
            pick = np.unique(np.concatenate([rng.choice(np.arange(cut[r], cut[r + 1]), size=64, replace=False) for r in range(1, world)]
                                            + [np.array([cut[r] for r in range(1, world)]), np.array([cut[r + 1] - 1 for r in range(world)])])).astype(np.int32)
            solo = kg.Plan(prob.H, p, K, prob.ref[:, None, :], prob.b, prob.mu, prob.groups, seed=prob.seed, unrelated=pick,
                           options=kg.Options(device=local_rank))
            solo.run()
            ref_rows, _ = solo.download()
            solo.close()
            same = bool(np.array_equal(ref_rows[pick], full_h[pick]))
            filled = bool(full_h.any(axis=1).all())
            check = {"rows_checked": int(len(pick)), "equal_to_single_gpu": same, "every_row_filled": filled}
            if not (same and filled):
                raise SystemExit("bench.py: the gathered knockoff matrix differs from a single-GPU run")
        dist.barrier()
    hk_plan, _ = plan.download()

    # ---- end to end through the one-shot C-ABI call with pinned host buffers
    e2e = None
    if not a.no_e2e and a.e2e_steps > 0:
        pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
        tH, tref, tb, tmu, tg = pin(prob.H), pin(prob.ref.astype(np.int32)), pin(prob.b), pin(prob.mu), pin(prob.groups)
        t_out = torch.zeros_like(tH).pin_memory()
        o1 = kg.Options(device=local_rank, block_loci=a.block_loci, lanes_per_hap=a.lanes, ctas_per_sm=a.ctas_per_sm)
        call = lambda: kg.generate_into(tH.numpy(), p, K, tref.numpy(), tb.numpy(), tmu.numpy(), tg.numpy(), out=t_out.numpy(),
                                        unrelated=mine, seed=prob.seed, options=o1)
        for _ in range(3):
            call()  # warm-up: first touch of pinned pages, module load, device buffer pool primed (W >= 3)
        barrier()
        per_call = []
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            tc = time.perf_counter()
            call()
            per_call.append(round((time.perf_counter() - tc) * 1e3, 2))
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        if not np.array_equal(t_out.numpy()[mine], hk_plan[mine]):
            raise SystemExit("bench.py: one-shot and resident-plan knockoffs differ")
        tables = p * (16 + 128 + 128)  # KgwLocusHMM + KgwLocusKO + KgwEmit rows
        e2e = {"value": a.haps * world * p * a.e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(prob.H.nbytes + prob.ref.nbytes + tables + 4 * len(mine)),
               "d2h_bytes_per_step": int(len(mine) * rb), "steps": a.e2e_steps, "ms_per_call": per_call,
               "call": "kgw_generate (one-shot: plan create + H2D + kernel + D2H + destroy), pinned host buffers"}

    # ---- roofline of the dominant (only) kernel, live CUDA-event durations
    k_ms = float(np.mean(kern_ms))
    traffic = None
    tf = REPO / "profiles" / "ncu_traffic.json"
    traffic_src = None
    if tf.exists():
        try:
            t = json.loads(tf.read_text())
            ent = t.get(f"{a.haps}x{a.snps}xK{a.K}xg{a.mean_group:g}", {})
            traffic, traffic_src = ent.get("dram_bytes_per_launch"), ent.get("source")
        except Exception:
            traffic = None
    roofline = roofline_dict(units, K, k_ms, "kgw::sampler_kernel", SAMPLER_NOTE, traffic)
    roofline["traffic_source"] = traffic_src or "profiles/ncu_traffic.json (ncu --set full capture of this command, dram__bytes_read.sum + dram__bytes_write.sum; not re-measured by this run)"
    # secondary bounds BASELINE.md section 3 asks for next to the contract roofline
    flops = (53.0 if a.mean_group <= 1.5 else 32.0) * K              # fp64 flops per hap-SNP (singleton / coarse groups)
    roofline["secondary"] = {
        "compulsory_bytes_per_hap_snp": (K + 2) / 8.0,
        "measured_dram_bytes_per_hap_snp": (traffic / units) if traffic else None,
        "fp64_flops_per_hap_snp": flops,
        "fp64_TFLOPs_achieved": units * flops / (k_ms * 1e-3) / 1e12,
        "fp64_TFLOPs_nominal": 37.0,
        "statement": "the kernel moves fewer bytes than the contract figure (no materialised backward table), so the HBM "
                     "fraction above is a rate of useful work, not of bus use; it is bound by per-warp dependency chains "
                     "(fp64 issue + shuffle latency), see DESIGN.md section 4.1; nominal fp64 peak from the B200 data sheet, not measured",
    }

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a, world), "clocks": clocks,
            "e2e": e2e, "gpu_launches": a.steps * (info["launches_per_run"] + (1 if world == 1 else 0)), "roofline": roofline,
            "value_kernel_only": value_kernel_only,
            "value_includes": "sampler kernel + pack kernel + D2H of the packed knockoff rows into pinned host memory, every step"
                              if world == 1 else "sampler kernel + staged copy + NCCL gather of the packed rows to rank 0, every step",
            "kernel": {k: info[k] for k in ("grid", "block", "lanes_per_hap", "states_per_lane", "scratch_bytes", "block_loci", "segment_loci")},
            "wall_s_timed_region": t_wall}
    if check is not None:
        line["sharded_result_check"] = check

    def guarded(name, fn):
        try:
            line[name] = fn()
        except Exception as e:
            line[name] = {"error": repr(e)[:300]}

    # ---- north-star configurations beside the headline (N = 1 only)
    if "c4" in legs and rank == 0:
        guarded("c4_resolutions", lambda: leg_c4_resolutions(torch, kg, synth, plan, prob, dev, stream, flush))

    # ---- next row of the scope table (outside the timed step, reported beside it): device-side .bed body of
    # the resident H / Hk (kgw_plan_encode_bed replaces writeBED, plink_interface.cpp:67-103)
    N = N_tot
    if "neighbours" in legs and rank == 0:
        def bed_leg():
            plan.run(); plan.sync()
            body, _ = plan.encode_bed()                      # warm-up
            t0 = time.perf_counter()
            body, bed_ms = plan.encode_bed(out=body)
            bed_wall = time.perf_counter() - t0
            return {"kernel_ms": bed_ms, "call_ms_with_d2h": bed_wall * 1e3, "output_bytes": int(body.nbytes),
                    "input_bytes": int(2 * N * rb),
                    "kernel_GBps": (body.nbytes + 2 * N * rb) / (bed_ms * 1e-3) / 1e9 if bed_ms > 0 else None,
                    "note": "2-bit SNP-major columns of H and Hk with the reference's swap pattern; not part of `value`"}
        guarded("bed_writer", bed_leg)

        # the row before the path: K nearest haplotypes inside kinship clusters of 5000 consecutive haplotypes on the
        # thinned chromosome (kgw_assign_references replaces findNeighbors, kinship_computer.cpp:174-233)
        def refs_leg():
            csize = 5000
            clusters = [np.arange(c0, min(c0 + csize, N), dtype=np.int32) for c0 in range(0, N, csize)]
            kg.assign_references(prob.H, p, K, clusters)                  # warm-up
            nn, ref_ms = kg.assign_references(prob.H, p, K, clusters)
            pt = (p + 9) // 10
            pair_words = float(sum(len(c) ** 2 for c in clusters)) * ((pt + 31) // 32)
            # the clustering half: one k-means iteration (distances to two centroids + new centroids) over all haplotypes
            # on the thinned matrix (every 10th SNP), kgw_kinship_* (kinship_computer.cpp:314-405)
            bits = np.unpackbits(prob.H, axis=1, bitorder="little")[:, :p][:, ::10]
            Ht = np.packbits(bits, axis=1, bitorder="little")
            km = kg.KinshipMatrix(Ht, bits.shape[1])
            members = np.arange(N, dtype=np.int32)
            mu0, mu1 = km.centroids(members, (members % 2).astype(np.int32))
            km.distances(members, mu0, mu1)                                # warm-up
            t0 = time.perf_counter()
            d0, d1 = km.distances(members, mu0, mu1)
            km.centroids(members, (d0 < d1).astype(np.int32) ^ 1)
            kmeans_ms = (time.perf_counter() - t0) * 1e3
            km.close()
            return {"device_ms": ref_ms, "pairs": float(sum(len(c) ** 2 for c in clusters)),
                    "kmeans_iteration_call_ms": kmeans_ms,
                    "thinned_snps": pt, "Gpair_words_per_s": pair_words / (ref_ms * 1e-3) / 1e9,
                    "note": "thin + all-pairs Hamming (xor/popc) + K-smallest with ties in cluster order; "
                            "device time incl. kernel launches, excl. H2D of H; not part of `value`"}
        guarded("reference_selection", refs_leg)

        # the producer of H: phased BGEN image of the benchmark's own population -> packed haplotype matrix
        # (kgw_bgen_read replaces the loop of BgenReader::read_worker, bgen_reader.cpp:139-261)
        def bgen_leg():
            image = synth.write_bgen_image(prob.H, p, compression=0)
            sidx, vcol = np.arange(N // 2, dtype=np.int32), np.arange(p, dtype=np.int32)
            kg.read_bgen_image(image, sidx, vcol)                          # warm-up
            t0 = time.perf_counter()
            Hb, bg_ms = kg.read_bgen_image(image, sidx, vcol, return_ms=True)
            bg_wall = time.perf_counter() - t0
            blk_bytes = p * (10 + N // 2 + N)
            return {"kernel_ms": bg_ms, "call_ms": bg_wall * 1e3, "identical_to_H": bool(np.array_equal(Hb, prob.H)),
                    "block_bytes": int(blk_bytes), "output_bytes": int(Hb.nbytes),
                    "kernel_GBps": (blk_bytes + Hb.nbytes) / (bg_ms * 1e-3) / 1e9 if bg_ms > 0 else None,
                    "note": "uncompressed layout-2 blocks, 8-bit probabilities; call = index + staging + H2D + kernel + D2H; "
                            "not part of `value`"}
        guarded("bgen_ingest", bgen_leg)

        # the same file zlib-compressed (what qctool writes by default): the blocks go to the GPU compressed and
        # bgen_inflate_kernel inflates them there (one warp per variant); beside it libz on the host threads
        def bgen_zlib_leg():
            import os
            pz = min(p, 4000)
            Hz = np.ascontiguousarray(prob.H[:, :(pz + 7) // 8]).copy()
            if pz % 8:
                Hz[:, -1] &= (1 << (pz % 8)) - 1
            image = synth.write_bgen_image(Hz, pz, compression=1)
            sidx, vcol = np.arange(N // 2, dtype=np.int32), np.arange(pz, dtype=np.int32)
            out = {"variants": int(pz), "compressed_bytes": int(image.nbytes), "block_bytes": int(pz * (10 + N // 2 + N))}
            prev = os.environ.get("KGW_BGEN_HOST_INFLATE")
            try:
                for mode in ("device", "host"):
                    os.environ["KGW_BGEN_HOST_INFLATE"] = "1" if mode == "host" else "0"
                    kg.read_bgen_image(image, sidx, vcol)                  # warm-up
                    t0 = time.perf_counter()
                    Hb, ms = kg.read_bgen_image(image, sidx, vcol, return_ms=True)
                    out[mode + "_inflate"] = {"call_ms": (time.perf_counter() - t0) * 1e3, "kernel_ms": ms,
                                              "identical_to_H": bool(np.array_equal(Hb, Hz))}
            finally:
                if prev is None:
                    os.environ.pop("KGW_BGEN_HOST_INFLATE", None)
                else:
                    os.environ["KGW_BGEN_HOST_INFLATE"] = prev
            out["note"] = "first 4000 variants of the population, zlib level 1; device: kernel_ms = inflate + header + unpack kernels; not part of `value`"
            return out
        guarded("bgen_zlib_ingest", bgen_zlib_leg)

        # the same for the text format (kgw_haps_read replaces the loop of HapsReader::read, haps_reader.cpp:93-124)
        def haps_leg():
            bits = np.unpackbits(prob.H, axis=1, bitorder="little")[:, :p]
            txt = np.full((p, 2 * N), ord(" "), dtype=np.uint8)
            txt[:, 0::2] = bits.T + ord("0")
            txt[:, -1] = ord("\n")
            txt = txt.reshape(-1)
            sidx, rows = np.arange(N // 2, dtype=np.int32), np.arange(p, dtype=np.int32)
            kg.read_haps_text(txt, sidx, rows)                             # warm-up
            t0 = time.perf_counter()
            Ht, hp_ms = kg.read_haps_text(txt, sidx, rows, return_ms=True)
            hp_wall = time.perf_counter() - t0
            return {"kernel_ms": hp_ms, "call_ms": hp_wall * 1e3, "identical_to_H": bool(np.array_equal(Ht, prob.H)),
                    "text_bytes": int(txt.nbytes), "output_bytes": int(Ht.nbytes),
                    "kernel_GBps": (txt.nbytes + Ht.nbytes) / (hp_ms * 1e-3) / 1e9 if hp_ms > 0 else None,
                    "note": "tokenizer + gather/transpose kernels; call = row index + H2D of the text + kernels + D2H; "
                            "not part of `value`"}
        guarded("haps_ingest", haps_leg)

    # ---- the IBD-conditioned branch of the same generate() call (generate_related, knockoffs.cpp:329-571) on
    # synthetic related pairs (SURVEY.md 8d C5 recipe) drawn inside the same population; reported beside `value`
    if "related" in legs and rank == 0 and a.related_pairs > 0:
        def related_leg():
            import copy
            rp = copy.copy(prob)
            rp.H, rp.ref = prob.H.copy(), prob.ref.copy()
            synth.add_ibd_pairs(rp, min(a.related_pairs, N // 4))
            fplan = kg.Plan(rp.H, p, K, rp.ref[:, None, :], rp.b, rp.mu, rp.groups, unrelated=np.zeros(0, np.int32),
                            ref_global=rp.ref, seed=rp.seed, families=rp.families, options=kg.Options(device=local_rank, stream=stream.cuda_stream))
            total_ms, kern, fclk = timed_runs(torch, fplan, stream, dev, 3, 1, flush)
            fplan.close()
            n_rel = 2 * len(rp.families)
            fam_ms = float(np.mean(kern))
            rf = roofline_dict(n_rel * p, K, fam_ms, "kgw::family_kernel",
                               "contract bytes of a related haplotype-SNP taken equal to an unrelated one: one fp64 K-vector written "
                               "and one read per locus (16 K + (K+2)/8 B), the least an implementation that materialises one message "
                               "table moves; the reference's loopy BP materialises two (messages_left / messages_right, "
                               "family_bp.cpp:421-427) and sweeps them up to 20 times per junction draw")
            return {"kernel_ms": fam_ms, "pairs": len(rp.families), "haplotypes": n_rel,
                    "hap_snps_per_s": n_rel * p / (fam_ms * 1e-3), "clocks": fclk, "roofline": rf,
                    "segments_per_pair": float(np.mean([len(sl) for _, sl in rp.families])),
                    "note": "kgw::family_kernel (loopy BP + junction draws + family knockoffs), related haplotypes only; not part of `value`"}
        guarded("related_branch", related_leg)

    plan.close()
    kg.release_workspace()
    if "c3" in legs and rank == 0:
        guarded("c3_shard", lambda: leg_c3_shard(torch, kg, synth, dev, stream, flush, a))
    if "c5" in legs and rank == 0:
        guarded("c5_shard", lambda: leg_c5_shard(torch, kg, synth, dev, stream, flush, a))

    if "cpu" in legs and rank == 0:
        try:
            # bounded sample: about 2.5 s per generate() call at ~1e6 hap-SNPs/s/thread
            thr = a.cpu_threads or len(os.sched_getaffinity(0))
            n_s = min(5000, a.haps, int(2.5e6 * thr / a.snps))
            line["cpu_baseline"], _, _ = cpu_baseline_dict(a, 3, 1, a.cpu_threads, n_s)
        except Exception as e:  # the GPU numbers stand on their own
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "error": repr(e)[:300]}
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world)
    else:
        if world != a.gpus and world == 1 and a.gpus > 1:
            raise SystemExit("bench.py --gpus N > 1 must be launched with torchrun (one rank per GPU)")
        run_b200_arm(a, rank, local_rank, world)


if __name__ == "__main__":
    main()
