#!/usr/bin/env python
"""bench.py -- knockoff haplotype-SNPs/s of the snpknock2 sampler hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the unmodified reference on the host CPUs

Workload (BASELINE.json configs[1]): synthetic chr22-shaped data, 20 000 haplotypes x 10 000 SNPs,
K = 50 references per haplotype, one partition resolution (singleton groups, the reference's most
expensive resolution), per GPU.  With N > 1 (torchrun, one rank per GPU) every rank processes its
own shard of 20 000 haplotypes of one population (weak scaling, no collective on the data path)
and the packed knockoff rows are all-gathered over NCCL inside the timed step.

A "step" is one Knockoffs::generate() call (snpknock2/src/knockoffs.cpp:311) for one resolution.
`value` is timed with CUDA events on the launching stream with every input resident in HBM;
`e2e` goes through the one-shot C-ABI call kgw_generate() with pinned HOST buffers (H2D, table
build, kernel, D2H inside the timed region).  Prints ONE JSON line (rank 0).
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--skip-sweeps", type=int, default=0, help="development only: timing experiments, results invalid")
    return ap.parse_args()


def workload_config(a, world):
    return {
        "workload": f"synthetic chr22-shaped: {a.haps} haplotypes x {a.snps} SNPs per GPU, K={a.K}, "
                    f"one partition resolution (mean group size {a.mean_group:g}), unrelated haplotypes",
        "haplotypes_per_gpu": a.haps, "snps": a.snps, "K": a.K, "mean_group_size": a.mean_group,
        "haplotypes_total": a.haps * world, "hmm_rho": RHO, "hmm_lambda": LAMBDA,
        "sharding": "by haplotype, no data-path collective; NCCL all-gather of the packed Hk rows of step i on a side stream, overlapped with the kernel of step i+1, the last one inside the timed region" if world > 1 else "single GPU",
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
def cpu_reference_run(a, steps: int, warmup: int, threads: int):
    """Times the reference's own CPU implementation of the path on a bounded haplotype sample of the
    workload (all SNPs, references re-drawn inside the sample -- cost is linear in haplotypes).
    Uses the unmodified reference binary in oracle/_ref/ when present (kind "reference"), else the
    C restatement in oracle/ (kind "port").  Returns (per-step seconds list, sample haps, kind, threads)."""
    sys.path.insert(0, str(REPO / "oracle"))
    from knockoffgwas_b200.synth import make_problem
    threads = threads or len(os.sched_getaffinity(0))
    # about 2.5 s per step at ~1e6 hap-SNPs/s/thread
    n_s = int(2.5e6 * threads / a.snps)
    n_s = max(2 * (a.K + 2), min(n_s, 5000, a.haps))
    n_s -= n_s % 2
    threads = min(threads, n_s)
    prob = make_problem(n_s, a.snps, a.K, seed=2020, mean_group=a.mean_group, rho=RHO, lam=LAMBDA, block=n_s)
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


def cpu_baseline_dict(a, steps, warmup, threads):
    secs, n_s, kind, thr = cpu_reference_run(a, steps, warmup, threads)
    v = n_s * a.snps * len(secs) / sum(secs)
    return {"value": v, "unit": UNIT, "cores": thr, "kind": kind,
            "sample": f"{n_s} of {a.haps} haplotypes x all {a.snps} SNPs, K={a.K}, references re-drawn inside the sample; "
                      f"{len(secs)} timed Knockoffs::generate() calls after {warmup} warm-up, --n_threads {thr}"}, secs, n_s


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    d, secs, n_s = cpu_baseline_dict(a, a.steps, a.warmup, a.cpu_threads)
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


def run_b200_arm(a, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import knockoffgwas_b200 as kg
    from knockoffgwas_b200 import synth
    from knockoffgwas_b200.synth import make_problem

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

    prob = make_problem(a.haps, a.snps, a.K, seed=2020, mean_group=a.mean_group, rho=RHO, lam=LAMBDA, hap_seed=rank)
    N, p, K = prob.N, prob.p, prob.K
    units = N * p  # haplotype-SNPs per step on this rank

    stream = torch.cuda.Stream(device=dev)
    opts = kg.Options(device=local_rank, block_loci=a.block_loci, lanes_per_hap=a.lanes, ctas_per_sm=a.ctas_per_sm,
                      stream=stream.cuda_stream, skip_sweeps=a.skip_sweeps)
    plan = kg.Plan(prob.H, p, K, prob.ref[:, None, :], prob.b, prob.mu, prob.groups, seed=prob.seed, options=opts)
    info = plan.info()
    hk_bytes = N * info["row_stride_bytes"]
    hk_dev = torch.as_tensor(_DevBuf(plan.device_hk_ptr(), hk_bytes), device=dev)
    gathered = torch.empty(world * hk_bytes, dtype=torch.uint8, device=dev) if world > 1 else None
    staged = torch.empty(hk_bytes, dtype=torch.uint8, device=dev) if world > 1 else None
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gather_done = [None]

    def step():
        # The gather of the packed knockoff rows (the only collective, NCCL over NVLink) runs on a side stream
        # from a staged copy, overlapped with the next step's kernel; it is inside the timed region (the end
        # event is recorded after the last gather).
        plan.run()
        if world > 1:
            if gather_done[0] is not None:
                stream.wait_event(gather_done[0])      # the previous gather has read `staged`
            staged.copy_(hk_dev, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                dist.all_gather_into_tensor(gathered, staged)
                gather_done[0] = torch.cuda.Event()
                gather_done[0].record(comm)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
    with torch.cuda.stream(stream):
        for _ in range(max(a.warmup, 0)):
            flush.zero_()
            step()
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
            if world > 1:
                if gather_done[0] is not None:
                    stream.wait_event(gather_done[0])
                staged.copy_(hk_dev, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(stream)
                with torch.cuda.stream(comm):
                    comm.wait_event(ready)
                    dist.all_gather_into_tensor(gathered, staged)
                    gather_done[0] = torch.cuda.Event()
                    gather_done[0].record(comm)
            ev[i][2].record(stream)
        if world > 1:
            stream.wait_event(gather_done[0])          # the last gather is part of the timed region
        ev_end.record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop()
    step_ms = [ev[i][0].elapsed_time(ev[i][2]) for i in range(a.steps)]
    kern_ms = [ev[i][0].elapsed_time(ev[i][1]) for i in range(a.steps)]
    total_ms = float(ev_start.elapsed_time(ev_end))   # K steps incl. the L2 flushes and, at N > 1, every gather
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms_max = float(tt.item())
    value = units * world * a.steps / (total_ms_max * 1e-3)

    hk_plan, _ = plan.download()

    # ---- end to end through the one-shot C-ABI call with pinned host buffers
    e2e = None
    if not a.no_e2e and a.e2e_steps > 0:
        pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
        tH, tref, tb, tmu, tg = pin(prob.H), pin(prob.ref.astype(np.int32)), pin(prob.b), pin(prob.mu), pin(prob.groups)
        t_out = torch.empty_like(tH).pin_memory()
        o1 = kg.Options(device=local_rank, block_loci=a.block_loci, lanes_per_hap=a.lanes, ctas_per_sm=a.ctas_per_sm)
        call = lambda: kg.generate_into(tH.numpy(), p, K, tref.numpy(), tb.numpy(), tmu.numpy(), tg.numpy(), out=t_out.numpy(),
                                        seed=prob.seed, options=o1)
        for _ in range(3):
            call()  # warm-up: first touch of pinned pages, module load, device buffer pool primed (W >= 3)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            call()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        if not np.array_equal(t_out.numpy(), hk_plan):
            raise SystemExit("bench.py: one-shot and resident-plan knockoffs differ")
        tables = p * (16 + 128 + 128)  # KgwLocusHMM + KgwLocusKO + KgwEmit rows
        e2e = {"value": units * world * a.e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(prob.H.nbytes + prob.ref.nbytes + tables + 4 * N),
               "d2h_bytes_per_step": int(N * ((p + 7) // 8)), "steps": a.e2e_steps,
               "call": "kgw_generate (one-shot: plan create + H2D + kernel + D2H + destroy), pinned host buffers"}

    # ---- roofline of the dominant (only) kernel, live CUDA-event durations
    peaks_file = REPO / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    k_ms = float(np.mean(kern_ms))
    achieved = units * b_alg(K) / (k_ms * 1e-3) / 1e9
    traffic = None
    tf = REPO / "profiles" / "ncu_traffic.json"
    if tf.exists():
        try:
            t = json.loads(tf.read_text())
            key = f"{a.haps}x{a.snps}xK{a.K}xg{a.mean_group:g}"
            traffic = t.get(key, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "kgw::sampler_kernel", "kernel_ms": k_ms,
                "algorithmic_bytes_per_hap_snp": b_alg(K), "peak_source": peak_src,
                "note": "contract bytes assume the reference's materialised fp64 backward table; this kernel stores one "
                        "message row every block_loci loci plus four partial sums per lane and locus and rebuilds single "
                        "entries on demand, so its DRAM traffic is below the algorithmic bytes (see DESIGN.md)"}

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
            "e2e": e2e, "gpu_launches": a.steps * info["launches_per_run"], "roofline": roofline,
            "kernel": {k: info[k] for k in ("grid", "block", "lanes_per_hap", "states_per_lane", "scratch_bytes", "block_loci", "segment_loci")},
            "wall_s_timed_region": t_wall}

    # ---- next row of the scope table (outside the timed step, reported beside it): device-side .bed body of
    # the resident H / Hk (kgw_plan_encode_bed replaces writeBED, plink_interface.cpp:67-103)
    if rank == 0 and not a.no_e2e:
        try:
            body, _ = plan.encode_bed()                      # warm-up
            t0 = time.perf_counter()
            body, bed_ms = plan.encode_bed(out=body)
            bed_wall = time.perf_counter() - t0
            line["bed_writer"] = {"kernel_ms": bed_ms, "call_ms_with_d2h": bed_wall * 1e3, "output_bytes": int(body.nbytes),
                                  "input_bytes": int(2 * N * ((p + 7) // 8)),
                                  "kernel_GBps": (body.nbytes + 2 * N * ((p + 7) // 8)) / (bed_ms * 1e-3) / 1e9 if bed_ms > 0 else None,
                                  "note": "2-bit SNP-major columns of H and Hk with the reference's swap pattern; not part of `value`"}
        except Exception as e:
            line["bed_writer"] = {"error": repr(e)[:200]}

    # ---- the row before the path (outside the timed step): K nearest haplotypes inside kinship clusters of 5000
    # consecutive haplotypes on the thinned chromosome (kgw_assign_references replaces findNeighbors,
    # kinship_computer.cpp:174-233)
    if rank == 0 and not a.no_e2e:
        try:
            csize = 5000
            clusters = [np.arange(c0, min(c0 + csize, N), dtype=np.int32) for c0 in range(0, N, csize)]
            kg.assign_references(prob.H, p, K, clusters)                  # warm-up
            nn, ref_ms = kg.assign_references(prob.H, p, K, clusters)
            pt = (p + 9) // 10
            pair_words = float(sum(len(c) ** 2 for c in clusters)) * ((pt + 31) // 32)
            line["reference_selection"] = {"device_ms": ref_ms, "pairs": float(sum(len(c) ** 2 for c in clusters)),
                                           "thinned_snps": pt, "Gpair_words_per_s": pair_words / (ref_ms * 1e-3) / 1e9,
                                           "note": "thin + all-pairs Hamming (xor/popc) + K-smallest with ties in cluster order; "
                                                   "device time incl. kernel launches, excl. H2D of H; not part of `value`"}
        except Exception as e:
            line["reference_selection"] = {"error": repr(e)[:200]}

    # ---- the producer of H (outside the timed step): phased BGEN image of the benchmark's own population ->
    # packed haplotype matrix (kgw_bgen_read replaces the loop of BgenReader::read_worker, bgen_reader.cpp:139-261)
    if rank == 0 and world == 1 and not a.no_e2e:
        try:
            image = synth.write_bgen_image(prob.H, p, compression=0)
            sidx, vcol = np.arange(N // 2, dtype=np.int32), np.arange(p, dtype=np.int32)
            kg.read_bgen_image(image, sidx, vcol)                          # warm-up
            t0 = time.perf_counter()
            Hb, bg_ms = kg.read_bgen_image(image, sidx, vcol, return_ms=True)
            bg_wall = time.perf_counter() - t0
            blk_bytes = p * (10 + N // 2 + N)
            line["bgen_ingest"] = {"kernel_ms": bg_ms, "call_ms": bg_wall * 1e3, "identical_to_H": bool(np.array_equal(Hb, prob.H)),
                                   "block_bytes": int(blk_bytes), "output_bytes": int(Hb.nbytes),
                                   "kernel_GBps": (blk_bytes + Hb.nbytes) / (bg_ms * 1e-3) / 1e9 if bg_ms > 0 else None,
                                   "note": "uncompressed layout-2 blocks, 8-bit probabilities; call = index + staging + H2D + kernel + D2H; "
                                           "not part of `value`"}
        except Exception as e:
            line["bgen_ingest"] = {"error": repr(e)[:200]}

    # ---- the same for the text format (kgw_haps_read replaces the loop of HapsReader::read, haps_reader.cpp:93-124)
    if rank == 0 and world == 1 and not a.no_e2e:
        try:
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
            line["haps_ingest"] = {"kernel_ms": hp_ms, "call_ms": hp_wall * 1e3, "identical_to_H": bool(np.array_equal(Ht, prob.H)),
                                   "text_bytes": int(txt.nbytes), "output_bytes": int(Ht.nbytes),
                                   "kernel_GBps": (txt.nbytes + Ht.nbytes) / (hp_ms * 1e-3) / 1e9 if hp_ms > 0 else None,
                                   "note": "tokenizer + gather/transpose kernels; call = row index + H2D of the text + kernels + D2H; "
                                           "not part of `value`"}
            del txt, bits
        except Exception as e:
            line["haps_ingest"] = {"error": repr(e)[:200]}

    # ---- the IBD-conditioned branch of the same generate() call (generate_related, knockoffs.cpp:329-571) on
    # synthetic related pairs (SURVEY.md 8d C5 recipe) drawn inside the same population; reported beside `value`
    if rank == 0 and world == 1 and not a.no_e2e and a.related_pairs > 0:
        try:
            import copy
            rp = copy.copy(prob)
            rp.H, rp.ref = prob.H.copy(), prob.ref.copy()
            synth.add_ibd_pairs(rp, min(a.related_pairs, N // 4))
            fplan = kg.Plan(rp.H, p, K, rp.ref[:, None, :], rp.b, rp.mu, rp.groups, unrelated=np.zeros(0, np.int32),
                            ref_global=rp.ref, seed=rp.seed, families=rp.families)
            fplan.run(); fplan.sync()
            fam_ms = fplan.run_timed(2)
            fplan.close()
            n_rel = 2 * len(rp.families)
            line["related_branch"] = {"kernel_ms": fam_ms, "pairs": len(rp.families), "haplotypes": n_rel,
                                      "hap_snps_per_s": n_rel * p / (fam_ms * 1e-3),
                                      "segments_per_pair": float(np.mean([len(sl) for _, sl in rp.families])),
                                      "note": "kgw::family_kernel (loopy BP + junction draws + family knockoffs), one warp per "
                                              "family, related haplotypes only; not part of `value`"}
        except Exception as e:
            line["related_branch"] = {"error": repr(e)[:200]}

    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            line["cpu_baseline"], _, _ = cpu_baseline_dict(a, 3, 1, a.cpu_threads)
        except Exception as e:  # the GPU numbers stand on their own
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "error": repr(e)[:300]}
    plan.close()
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
