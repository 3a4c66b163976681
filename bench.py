#!/usr/bin/env python
"""bench.py - Gpairs/s of 256-bit Hamming brute-force matching / ms per loop check (BASELINE.json metric).

One "step" = one loop check: a 2000-descriptor query frame against the whole keyframe DB
(BASELINE config 4: 100k keyframes x 2000 rows = 200 M rows, 6.4 GB), i.e. 400 G descriptor pairs.
With --gpus N (launched by torchrun, one rank per GPU) the DB is sharded keyframe k -> rank k % N
(strong scaling: total work fixed), and every step ends with the path's real exchange: ONE NCCL all_gather of
every rank's [per-keyframe records | packed per-query bests] and one merge kernel (unsigned min of the bests,
records re-interleaved into global keyframe order).

  value   : whole-job Gpairs/s with the query already resident in HBM (lc_loop_check_batch_device on
            torch's current stream + the collectives), CUDA events, max over ranks.
  e2e     : the same through the host-buffer API (lc_loop_check at N=1, ShardedLoopDB.loop_check at N>1):
            query H2D from pinned memory and records + best D2H inside the timed region.
  roofline: tensor bound (the default path is a +/-1 FP4 GEMM on tcgen05: 512 flop per pair) against the measured
            cuBLAS bf16 peak x 4; the integer-popc roofline of SURVEY.md 8d (148 SM x 16 POPC/clk x f_SM / 8) is kept
            beside it.  The kernel time is liblc_b200's own CUDA-event bracket around the matching kernel.
  cpu_baseline: cv2.BFMatcher(NORM_HAMMING).match per keyframe + numpy filter (the OpenCV call the
            reference's matcher_ makes), all host threads, on a bounded sample of the same workload.

--impl reference runs only that CPU arm (the reference is CPU-only; its own loop_closing.cpp is absent and
its OpenCV C++ dependency is not in the image, so the arm is OpenCV's matcher through cv2 + the restated
README rule in oracle/cv2_ref.py).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NQ = 2000
ROWS_PER_KF = 2000
QUERY_SEED, DB_SEED = 1234, 5678
CHUNK_KF = 2500  # keyframes generated per device chunk (160 MB)
ARM_PAUSE_S = 1.0  # idle time before each arm's warm-up (see main)
POPC_PER_CLK_PER_SM = 16  # measured: profiles/r01_pipe_rates_b200.json (15.1 sustained in a pure-POPC loop)
POPC_PER_PAIR = 8
# ncu --set full, lc::fx::mxf4_match_kernel on 4000 keyframes x 2000 rows (profiles/r02_ncu_mxf4_match.txt: 256.4 MB read,
# 16.2 MB written):
# dram read + write bytes per keyframe (64,000 B of rows are the algorithmic read; 4,000 B of u16 distances the write)
TRAFFIC_BYTES_PER_KF = 68_165

WORKLOADS = {
    "c4": dict(keyframes=100_000, name="C4: 2000-descriptor query vs 100k-keyframe DB (200M rows x 256 bit, 6.4 GB)"),
    "c3": dict(keyframes=1_000, name="C3: 2000-descriptor query vs 1k-keyframe DB (2M rows, 64 MB)"),
    # BASELINE config 5 (throughput mode): one step = ONE batched check of 64 query frames (25.6 Tpairs)
    "c5": dict(keyframes=100_000, frames=64, name="C5: batched loop check, 64 query frames x 2000 descriptors vs 100k-keyframe DB (200M rows, 6.4 GB)"),
}


def workload_config(args, world):
    """The `config` object of the JSON line - the same for both arms (the reference arm times a bounded sample OF this
    workload and says so in cpu_baseline.sample)."""
    wl = WORKLOADS[args.workload]
    n_kf = args.keyframes or wl["keyframes"]
    local_rows = ((n_kf + world - 1) // world) * ROWS_PER_KF  # rank 0's shard, the largest
    frames = wl.get("frames", 1)
    return {"workload": wl["name"] if not args.keyframes else f"{frames} query frame(s) vs {n_kf}-keyframe DB x {ROWS_PER_KF} rows", "keyframes": n_kf,
            "rows_per_keyframe": ROWS_PER_KF, "nq": NQ, "query_frames": frames, "pairs_per_step": float(frames) * NQ * ROWS_PER_KF * n_kf,
            "sharding": f"keyframe k -> rank k % {world}; one exchange of [records | packed bests] per step: a single kernel over peer "
                        f"memory (CUDA IPC, NVLink) or, as the fallback, one NCCL all_gather + a merge kernel - see `exchange`"
                        if world > 1 else "single GPU",
            "kernel": args.kernel,
            "timing": f"each arm: {ARM_PAUSE_S:.0f} s idle, warm-up steps, then the timed steps (CUDA events, max over ranks)",
            "l2": f"inputs larger than L2: each step streams the {local_rows * 32 / 1e9:.2f} GB DB shard from HBM"}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def query_rows():
    return np.random.default_rng(QUERY_SEED).integers(0, 256, (NQ, 32), np.uint8)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi-equivalent clock / throttle-reason samples (NVML) taken DURING the timed region."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self._th = threading.Thread(target=self._run, daemon=True)
            self._th.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._th:
            self._th.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_loop_check_sample(q, n_kf, threads, seed=DB_SEED):
    """The reference's access pattern on the host: one BFMatcher.match(q, kf) per keyframe, then the
    2 x min filter / count in numpy.  Returns (seconds, pairs, kind)."""
    from oracle import cv2_ref
    from oracle.params import Params
    import datagen
    kfs = [datagen.keyframe(k, ROWS_PER_KF, seed) for k in range(n_kf)]
    fids = list(range(n_kf))
    if cv2_ref.available():
        import cv2
        cv2.setNumThreads(threads)
        t0 = time.perf_counter()
        cv2_ref.loop_check(q, kfs, fids, n_kf + 1000, Params(loop_threshold=0.15), want_best=True)
        dt = time.perf_counter() - t0
        kind = "cv2.BFMatcher(NORM_HAMMING).match per keyframe + numpy 2xmin filter (OpenCV %s)" % cv2.__version__
    else:  # the plain-C port
        from oracle import c_oracle
        t0 = time.perf_counter()
        c_oracle.loop_check(q, np.concatenate(kfs), [len(k) for k in kfs], fids, n_kf + 1000, Params(), nthreads=threads)
        dt = time.perf_counter() - t0
        kind = "oracle/lc_oracle.c (plain-C port, pthreads)"
    return dt, float(len(q)) * ROWS_PER_KF * n_kf, kind


def cpu_baseline(q, budget_s=12.0):
    threads = os.cpu_count() or 1
    dt, pairs, kind = cpu_loop_check_sample(q, 8, threads)  # probe (also warms cv2's thread pool)
    n = int(max(16, min(4000, budget_s / (dt / 8))))
    dt, pairs, kind = cpu_loop_check_sample(q, n, threads)
    n1 = int(max(8, min(200, n // (4 * max(threads // 2, 1)))))  # the same at cv::setNumThreads(1), ~3 s
    dt1, pairs1, _ = cpu_loop_check_sample(q, n1, 1)
    return {"value": pairs / dt / 1e9, "unit": "Gpairs/s", "cores": threads, "kind": "port",
            "single_thread": {"value": pairs1 / dt1 / 1e9, "unit": "Gpairs/s", "cores": 1, "sample": f"{n1} keyframes, cv2.setNumThreads(1)"},
            "sample": f"{n} keyframes x {ROWS_PER_KF} rows vs the {NQ}-row query ({pairs / 1e9:.1f} Gpairs, {dt:.1f} s): {kind}; "
                      f"the reference's own loop_closing.cpp is absent, this is the OpenCV call its matcher_ makes",
            "ms_per_keyframe": dt / n * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    q = query_rows()
    threads = os.cpu_count() or 1
    dt, _, _ = cpu_loop_check_sample(q, 8, threads)
    n = int(max(8, min(2000, 4.0 / (dt / 8))))  # ~4 s per step
    times = []
    for s in range(args.warmup + args.steps):
        dt, pairs, kind = cpu_loop_check_sample(q, n, threads)
        if s >= args.warmup:
            times.append(dt)
    tot = float(np.sum(times))
    val = pairs * len(times) / tot / 1e9
    sample = (f"each step = a bounded sample of the workload: {n} of its keyframes x {ROWS_PER_KF} rows vs the {NQ}-row query "
              f"({pairs / 1e9:.1f} Gpairs); {kind}")
    print(json.dumps({
        "impl": "reference", "metric": "Gpairs/s 256-bit Hamming BF match (loop check vs keyframe DB)", "value": val,
        "unit": "Gpairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot / len(times) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": val, "unit": "Gpairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def build_db(db, torch, n_kf, rank, world, device, plant_at, q_dev):
    """Synthetic DB generated on the device, chunk by chunk; chunk bytes depend only on (DB_SEED, chunk) so a
    keyframe's bytes never depend on the shard layout.  Keyframe `plant_at` is a row-permuted exact copy of
    the query (known answer for the full-size check)."""
    perm = torch.from_numpy(np.random.default_rng(QUERY_SEED + 1).permutation(NQ)).to(device)
    ctx = db.shard.ctx
    for c0 in range(0, n_kf, CHUNK_KF):
        c1 = min(n_kf, c0 + CHUNK_KF)
        g = torch.Generator(device=device)
        g.manual_seed(DB_SEED * 1_000_003 + c0)
        chunk = torch.randint(0, 256, (c1 - c0, ROWS_PER_KF, 32), dtype=torch.uint8, device=device, generator=g)
        if c0 <= plant_at < c1:
            chunk[plant_at - c0] = q_dev[perm]
        first = (rank - c0) % world  # first owned keyframe inside the chunk
        mine = chunk[first::world].contiguous()
        ctx.db_append_many_device(np.arange(c0, c1, dtype=np.int32), np.full(c1 - c0, ROWS_PER_KF, np.int32), mine.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream or 1)  # copies ordered behind the generator kernels
        ctx.synchronize()
        db.n_keyframes += c1 - c0
        del chunk, mine
    return perm.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--keyframes", type=int, default=0, help="override the workload's keyframe count")
    ap.add_argument("--kernel", default="auto", choices=["auto", "popc", "csa", "tensor", "mxf4"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import slam_loop_closing_b200 as lc
    from slam_loop_closing_b200.sharded import ShardedLoopDB

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: liblc_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wl = WORKLOADS[args.workload]
    n_kf = args.keyframes or wl["keyframes"]
    kernel = {"auto": lc.KERNEL_AUTO, "popc": lc.KERNEL_POPC, "csa": lc.KERNEL_CSA, "tensor": lc.KERNEL_TENSOR, "mxf4": lc.KERNEL_MXF4}[args.kernel]
    cur_id = n_kf + 29  # every keyframe is at least min_loop_gap back
    q = query_rows()
    F = wl.get("frames", 1)
    import datagen
    qs = [q] + [datagen.uniform(NQ, 7000 + i) for i in range(F - 1)]  # frame 0 is the query of the single-frame workloads
    db = ShardedLoopDB(device=local, loop_threshold=0.15, kernel=kernel, timing=1)  # every check event-timed (match_ms feeds the roofline)
    ctx = db.shard.ctx
    q_dev_all = torch.from_numpy(np.concatenate(qs)).to(device)
    q_dev = q_dev_all[:NQ]
    torch.cuda.synchronize()
    plant_at = n_kf // 3
    perm = build_db(db, torch, n_kf, rank, world, device, plant_at, q_dev)
    sizes = ctx.db_size()
    pairs_per_step = float(F) * NQ * sizes["rows"]
    nq_arr, cur_arr = np.full(F, NQ, np.int32), np.full(F, cur_id, np.int32)

    # ---- correctness gate at full size (size-independent properties; outside the timed region)
    rec, best = db.loop_check(qs, [cur_id] * F)
    problems = []
    r = rec[0]
    if F > 1 and not all((rec[f]["n_good"] >= 0).all() and rec[f][plant_at]["min_dist"] > 0 for f in range(1, F)):
        problems.append("records of the other query frames malformed")
    if not (len(r) == n_kf and (r["n_good"] >= 0).all() and (r["frame_id"] == np.arange(n_kf)).all()):
        problems.append("records malformed")
    if not (r[plant_at]["min_dist"] == 0 and r[plant_at]["n_good"] == NQ):
        problems.append(f"planted keyframe record {r[plant_at]}")
    want_best = (np.uint64(plant_at) * np.uint64(ROWS_PER_KF) + np.argsort(perm).astype(np.uint64))  # dist 0 in the high word
    if not np.array_equal(best[0], want_best):
        problems.append("merged per-query best != planted rows")
    if rank == 0:  # sampled keyframes against the CPU oracle (checker only)
        from oracle import c_oracle
        from oracle.params import Params
        for k in [kk for kk in (0, 1, n_kf // 2, n_kf - 1) if kk % world == rank]:
            _, rows = ctx.db_read(k)
            w = c_oracle.match_features(q, rows, Params(), nthreads=os.cpu_count() or 1)
            if (int(r[k]["min_dist"]), int(r[k]["n_good"])) != (w["min_dist"], w["n_good"]):
                problems.append(f"keyframe {k}: gpu {r[k]} vs oracle {(w['min_dist'], w['n_good'])}")
    if problems:
        raise SystemExit("PARITY FAILURE at full size: " + "; ".join(problems))

    # ---- device-resident arm ("value")
    def step_device():
        return db.merged_device(q_dev_all, nq_arr, cur_arr)

    # Both arms start from the same state: a second of idle (a B200 that has just run a second of FP4 MMAs is 1-2 % slower
    # under its power cap than one that starts from idle - tools/arm_order_probe.py - and the arms run one after the other),
    # then the warm-up steps, then the timed steps.
    barrier()
    time.sleep(ARM_PAUSE_S)
    for _ in range(args.warmup):
        step_device()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            out = step_device()
        ev1.record()
        barrier()
        dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
        gpu_launches = ctx.launch_count() - launches0

        # ---- end-to-end arm: host query in, host records + best out, every step
        def step_e2e():
            if world == 1:
                return ctx.loop_check(q, cur_id, want_best=True) if F == 1 else ctx.loop_check_batch(qs, [cur_id] * F)
            return db.loop_check(qs, [cur_id] * F, copy=False)  # results read where the download left them (pinned host memory)

        barrier()
        time.sleep(ARM_PAUSE_S)
        for _ in range(min(args.warmup, 3)):
            step_e2e()
        barrier()
        kernel_ms, match_ms, match_launches = [], [], 1
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(args.steps):
            step_e2e()
            if world == 1:
                tm = ctx.last_timing()
                kernel_ms.append(tm.kernel_ms)
                match_ms.append(tm.match_ms)
                match_launches = max(1, tm.match_launches)
        ev1.record()
        barrier()
        e2e_wall_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e_ms = max(e2e_wall_ms, max_over_ranks(ev0.elapsed_time(ev1)))
    clocks = clk.summary()

    # ---- dominant-kernel time for the roofline: CUDA events recorded by the library on its own stream around the
    #      launches of the matching kernel (lc_timing.match_ms), live, inside the e2e timed region
    if world > 1:
        for _ in range(min(3, args.steps)):
            ctx.loop_check(q, cur_id, want_best=True) if F == 1 else ctx.loop_check_batch(qs, [cur_id] * F)
            tm = ctx.last_timing()
            kernel_ms.append(tm.kernel_ms)
            match_ms.append(tm.match_ms)
            match_launches = max(1, tm.match_launches)
    # the FP4 tensor pipe's own peak on this GPU, now: a pure tcgen05.mma kind::mxf4 issue loop (lc_probe_mxf4_peak)
    pipe_macs, pipe_mhz = ctx.probe_mxf4_peak(200)
    k_ms = max_over_ranks(float(np.mean(kernel_ms)))
    m_ms = max_over_ranks(float(np.mean(match_ms)))
    timing = ctx.last_timing()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    pk = peaks()
    sm_max = float(pk.get("sm_max_mhz", clocks.get("sm_max_mhz") or 1965.0))
    n_sm = torch.cuda.get_device_properties(device).multi_processor_count
    pairs_per_gpu = pairs_per_step / world
    kernel_names = {1: "lc::kf_scan_kernel", 2: "lc::split_kernel", 3: "lc::tc::tensor_match_kernel", 4: "lc::fx::mxf4_match_kernel"}
    hbm_peak = float(pk.get("hbm_gbs", 6650.0))
    algo_bytes = sizes["local_rows"] * 32 + F * (NQ * 32 + sizes["local_keyframes"] * (16 + 2 * NQ) + NQ * 8)
    hbm = {"achieved": algo_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
           "frac": algo_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": algo_bytes,
           "peak_source": "measured" if pk else "fallback"}
    popc_peak = n_sm * POPC_PER_CLK_PER_SM * sm_max * 1e6 / POPC_PER_PAIR / 1e9  # Gpairs/s per GPU
    pairs_s = pairs_per_gpu / (m_ms * 1e-3)
    if timing.path in (3, 4):
        # tensor bound.  Algorithmic work: one pair = a 256-term +/-1 dot product = 256 MAC = 512 flop.
        # Peak: the driver-measured cuBLAS bf16 figure scaled by the pipe's rate ratio (kind::mxf4 issues 4x, kind::i8
        # 2x the bf16 MACs per clock: 16384 / 8192 / 4096 MAC/clk/SM) - there is no measured FP4 or int8 entry.
        ratio, kind = (4.0, "fp4 (kind::mxf4)") if timing.path == 4 else (2.0, "int8 (kind::i8)")
        bf16 = float(pk.get("bf16_tflops", 1590.0))
        bf16_sus = float(pk.get("bf16_tflops_sustained", 1400.0))
        ach = pairs_s * 512 / 1e12
        mac_clk = 16384 if timing.path == 4 else 8192
        roof = {
            "bound": "tensor", "achieved": ach, "peak": ratio * bf16, "unit": "TFLOP/s", "frac": ach / (ratio * bf16),
            "traffic": TRAFFIC_BYTES_PER_KF * sizes["local_keyframes"] / match_launches if F == 1 else None,  # measured for the single-frame check only
            "kernel": kernel_names[timing.path], "path": "tcgen05 mxf4 GEMM + max, best row resolved afterwards" if timing.path == 4 else "tcgen05 int8 GEMM + arg-max",
            "ms_per_launch": m_ms / match_launches, "launches_per_step": match_launches,
            "algorithmic_flop_per_launch": pairs_per_gpu * 512 / match_launches,
            "peak_source": f"{'measured' if pk else 'fallback'} cuBLAS bf16 burst {bf16:.1f} TFLOP/s x {ratio:.0f} ({kind} MAC rate / bf16 MAC rate); "
                           f"of {'measured' if pk else 'fallback'}",
            "frac_of_sustained": ach / (ratio * bf16_sus),
            "nominal": {"peak": 2250.0 * ratio, "frac": ach / (2250.0 * ratio)},
            "pipe_peak_measured": {"peak": pipe_macs * 2 / 1e12, "unit": "TFLOP/s", "frac": ach / (pipe_macs * 2 / 1e12), "sm_mhz_held": pipe_mhz,
                                   "how": "lc_probe_mxf4_peak in this run, after the timed region: 200 ms of nothing but tcgen05.mma kind::mxf4 "
                                          "(M128 N256 K64, operands in shared memory) on every SM; the clock is cycle counter / elapsed time"}
                                  if timing.path == 4 else None,
            "pipe": {"pairs_per_clk_per_sm": pairs_s / n_sm / ((clocks.get("sm_mhz") or sm_max) * 1e6),
                     "peak_pairs_per_clk_per_sm": mac_clk / 256, "at_sm_mhz": clocks.get("sm_mhz") or sm_max},
            "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per keyframe of one --set full capture "
                              "(profiles/), scaled to this launch's keyframes",
            "note": "per GPU; CUDA events on the library's stream around the matching kernel's launches inside the e2e timed "
                    "region; a fraction above 1 means the FP4 pipe holds its clock under the power cap better than cuBLAS bf16",
            "what_is_left": "tensor pipe 93 % active (ncu); the rest is the accumulator read-out - 4 B leave TMEM per 256 MACs because "
                            "K = 256 is the descriptor length - beside the MMAs' own accumulator traffic, then 3 % tile padding "
                            "(2000 rows = 15.6 query tiles, 8.9 train tiles): profiles/r02_ab_drain_schedule.txt",
            "hbm": hbm, "int_popc": {"peak": popc_peak, "unit": "Gpairs/s", "x": pairs_s / 1e9 / popc_peak},
        }
    else:
        ach = pairs_s / 1e9
        roof = {
            "bound": "int_popc", "achieved": ach, "peak": popc_peak, "unit": "Gpairs/s", "frac": ach / popc_peak, "traffic": None,
            "kernel": kernel_names.get(timing.path, "?"), "path": {1: "whole-keyframe scan", 2: "split + finalize"}.get(timing.path, "?"),
            "ms_per_launch": m_ms / match_launches,
            "note": f"per GPU; integer kernels (neither HBM nor tensor bounds them): peak = {n_sm} SM x {POPC_PER_CLK_PER_SM} POPC/clk x "
                    f"{sm_max:.0f} MHz / {POPC_PER_PAIR} POPC per pair; the carry-save kernel needs 4 POPC per pair, so frac may exceed 1",
            "hbm": hbm,
        }
    line = {
        "metric": "Gpairs/s 256-bit Hamming BF match (loop check vs keyframe DB)",
        "value": pairs_per_step * args.steps / (dev_ms * 1e-3) / 1e9, "unit": "Gpairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": {4: "fp4", 3: "s8"}.get(timing.path, "u32"), "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": {"value": pairs_per_step * args.steps / (e2e_ms * 1e-3) / 1e9, "unit": "Gpairs/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": F * NQ * 32,
                "d2h_bytes_per_step": F * (int(n_kf * 16 + NQ * 8) if world == 1 else int((n_kf + world) * 16 + NQ * 8)),
                "api": ("lc_loop_check (C ABI, host buffers)" if F == 1 else "lc_loop_check_batch (C ABI, host buffers)") if world == 1
                       else "ShardedLoopDB.loop_check (host buffers, NCCL merge)"},
        "gpu_launches": int(gpu_launches), "clocks": clocks, "roofline": roof,
        "exchange": None if world == 1 else ("p2p: lc_merge_shards_peer_device, one kernel signals, waits and merges out of the peers' slots"
                                             if db._xchg_bytes else "nccl: all_gather_into_tensor + lc_merge_shards_device"),
        "parity": "full-size gate passed: planted-copy record, merged per-query best, sampled keyframes vs CPU oracle",
    }
    if world == 1 and not args.no_extra and F == 1:
        line["extra"] = extras(lc, torch, q, kernel)
        if n_kf >= 1000:  # BASELINE config 5 in shape, on this GPU's resident DB: 64 query frames in one batched check
            frames = [q] + [datagen.uniform(NQ, 7000 + i) for i in range(63)]
            ctx.loop_check_batch(frames, [cur_id] * 64)  # first call of this shape: result buffers (102 MB pinned) grow
            t0 = time.perf_counter()
            res, _ = ctx.loop_check_batch(frames, [cur_id] * 64)
            dt = time.perf_counter() - t0
            tm = ctx.last_timing()
            ok = bool(np.array_equal(res[0], r))  # frame 0 is the single check's query: same records expected
            line["extra"]["c5_batch_64_frames"] = {"e2e_ms": dt * 1e3, "e2e_gpairs_s": 64 * pairs_per_step / dt / 1e9,
                                                   "match_kernel_ms": tm.match_ms, "match_launches": tm.match_launches,
                                                   "frame0_records_equal_single_check": ok,
                                                   "api": "lc_loop_check_batch (C ABI, 64 host query frames in, records + bests out)"}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(q)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def c1_sequence(lc, kernel, n_frames=97, cpu_frames=8):
    """BASELINE config 1 in shape (the reference's own run: ~97 processed frames, ORB 2000, min_loop_gap 30, threshold 0.15)
    on synthetic descriptors: per frame matchFeatures(prev, cur) + store append + detectLoops(cur), i.e. the calls
    processFrame makes (include/loop_closing.hpp:34), through the C ABI with host buffers.  ORB detection and the pose
    stages stay on the host and are not part of it.  The same calls through cv2 on the last `cpu_frames` frames beside it."""
    import datagen
    from oracle import cv2_ref
    from oracle.params import Params
    frames = [datagen.uniform(2000, 900 + i) for i in range(n_frames)]
    p = Params(loop_threshold=0.15, min_loop_gap=30)

    def run(c, fused=False):
        per = []
        for i, d in enumerate(frames):
            t0 = time.perf_counter()
            if fused:  # lc_process_frame: one upload, one GPU submission, one wait per frame
                c.process_frame(i, d)
            else:
                if i:
                    c.match_pair(frames[i - 1], d)
                c.db_append(i, d)
                c.loop_check(d, i, want_best=False)
            per.append(time.perf_counter() - t0)
        return per

    with lc.Context(kernel=kernel, loop_threshold=p.loop_threshold, min_loop_gap=p.min_loop_gap) as c:
        run(c)  # warm-up pass (buffer growth, first launches)
    with lc.Context(kernel=kernel, loop_threshold=p.loop_threshold, min_loop_gap=p.min_loop_gap) as c:
        per = run(c)
    with lc.Context(kernel=kernel, loop_threshold=p.loop_threshold, min_loop_gap=p.min_loop_gap) as c:
        run(c, True)
    with lc.Context(kernel=kernel, loop_threshold=p.loop_threshold, min_loop_gap=p.min_loop_gap) as c:
        per_fused = run(c, True)
    pair_matches = (n_frames - 1) + sum(max(0, i - 29) for i in range(n_frames))  # consecutive + eligible loop matches
    out = {"gpu_ms_total": float(np.sum(per) * 1e3), "gpu_us_per_frame_median": float(np.median(per) * 1e6),
           "gpu_us_last_frame": float(per[-1] * 1e6), "pair_matches": pair_matches,
           "gpu_gpairs_s": pair_matches * 4e6 / float(np.sum(per)) / 1e9,
           "api": "lc_match_pair + lc_db_append + lc_loop_check per frame (ctypes, host buffers)",
           "fused": {"gpu_ms_total": float(np.sum(per_fused) * 1e3), "gpu_us_per_frame_median": float(np.median(per_fused) * 1e6),
                     "gpu_us_last_frame": float(per_fused[-1] * 1e6), "gpu_gpairs_s": pair_matches * 4e6 / float(np.sum(per_fused)) / 1e9,
                     "api": "lc_process_frame per frame: pair match vs the resident previous frame + append + loop check in one submission"}}
    if cv2_ref.available():
        import cv2
        cv2.setNumThreads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        n = 0
        for i in range(n_frames - cpu_frames, n_frames):
            cv2_ref.match_features(frames[i - 1], frames[i], p)
            cv2_ref.loop_check(frames[i], frames[:i + 1], list(range(i + 1)), i, p, want_best=False)
            n += 1 + max(0, i - 29)
        dt = time.perf_counter() - t0
        out.update({"cpu_ms_per_frame_last_frames": dt / cpu_frames * 1e3, "cpu_gpairs_s": n * 4e6 / dt / 1e9, "cpu_frames_timed": cpu_frames,
                    "cpu_cores": os.cpu_count() or 1, "cpu_kind": "cv2.BFMatcher(NORM_HAMMING).match per frame pair (OpenCV %s)" % cv2.__version__})
    return out


def extras(lc, torch, q, kernel):
    """The other single-GPU BASELINE configs, for context (not the headline): C2 pair latency, C3 loop check."""
    import datagen
    out = {}
    t = datagen.uniform(2000, DB_SEED)
    with lc.Context(kernel=kernel, loop_threshold=0.15) as c:
        for _ in range(5):
            c.match_pair(q, t)
        ts = []
        for _ in range(200):  # default context: pair calls are not event-timed (each record costs ~3 us of a ~60 us call)
            t0 = time.perf_counter()
            c.match_pair(q, t)
            ts.append(time.perf_counter() - t0)
        with lc.Context(kernel=kernel, timing=1) as ct:  # the same call with the library's CUDA events on, for the kernel time
            ks = []
            for _ in range(55):
                ct.match_pair(q, t)
                ks.append(ct.last_timing().kernel_ms)
            ks = ks[5:]
        out["c2_pair_2000x2000"] = {"e2e_us_median": float(np.median(ts) * 1e6), "e2e_gpairs_s": 4e6 / float(np.median(ts)) / 1e9,
                                    "kernel_us_median": float(np.median(ks) * 1e3),
                                    "kernel_gpairs_s": 4e6 / (float(np.median(ks)) * 1e-3) / 1e9, "launches_per_call": 1,
                                    "api": "lc_match_pair (C ABI via ctypes, host buffers in and out)"}
        out["c1_sequence_97_frames"] = c1_sequence(lc, kernel)
        g = torch.Generator(device="cuda")
        g.manual_seed(DB_SEED)
        rows = torch.randint(0, 256, (1000 * 2000, 32), dtype=torch.uint8, device="cuda", generator=g)
        c.db_append_many_device(np.arange(1000, dtype=np.int32), np.full(1000, 2000, np.int32), rows.data_ptr(),
                                torch.cuda.current_stream().cuda_stream or 1)
        c.synchronize()
        for _ in range(3):
            c.loop_check(q, 1029)
        ts, ks = [], []
        for _ in range(10):
            t0 = time.perf_counter()
            c.loop_check(q, 1029)
            ts.append(time.perf_counter() - t0)
            ks.append(c.last_timing().kernel_ms)
        out["c3_1k_keyframes"] = {"e2e_ms_median": float(np.median(ts) * 1e3), "kernel_ms_median": float(np.median(ks)),
                                  "kernel_gpairs_s": 4e9 / (float(np.median(ks)) * 1e-3) / 1e9, "path": c.last_timing().path}
        # the shipped pipeline's matcher (src/main.cpp:509-534): 4000 SIFT descriptors vs 300 views x 4000, knn k=2 + ratio
        rng = np.random.default_rng(DB_SEED)
        base = rng.integers(0, 200, (4000, 128)).astype(np.uint8)
        for k in range(300):
            c.l2_db_append(k, np.roll(base, k, axis=0))
        sq = rng.integers(0, 200, (4000, 128)).astype(np.uint8)
        for _ in range(2):
            c.l2_loop_check(sq, 400, 1)
        ts, ks = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            c.l2_loop_check(sq, 400, 1)
            ts.append(time.perf_counter() - t0)
            ks.append(c.last_timing().match_ms)
        out["sift_l2_knn2_300_views_4000x4000"] = {"e2e_ms_median": float(np.median(ts) * 1e3), "kernel_ms_median": float(np.median(ks)),
                                                  "kernel_gpairs_s_128d": 300 * 16e6 / (float(np.median(ks)) * 1e-3) / 1e9}
    return out


if __name__ == "__main__":
    sys.exit(main())
