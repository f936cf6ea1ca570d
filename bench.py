#!/usr/bin/env python
"""bench.py -- tree-likelihood evals/sec of the BAMSE hot path on B200 (contract in the task
statement, section 4 of the tier framing).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl ours|reference]

Default workload: BASELINE.json configs[2] (c3: 20 samples, 2000 mutations, -nc 8 -n 100), the configuration
the baseline quotes for 1/2/4/8 B200 and the largest one whose every leg (resident, enumeration-only, early-exit,
end to end, CPU sample) fits a few minutes; the same workload at every N, so the per-N values are one strong-scaling
curve.  `--config c2` is the round-1 default (a 4 ms step since the pair path: launch-bound beyond 2 GPUs);
lines of every configuration from one build are in profiles/scale/.

One "step" = one pass of the hot path over the configuration's whole work list: every labeled
rooted tree (K^(K-1)) of every candidate clustering (x every (e,s) pair for c5), thresholded
and merged into the global top-t.  With N > 1 (torchrun, one rank per GPU) the tree-index
range of every clustering is split N ways (no data-path collective); the only exchange is one
NCCL all-gather of each rank's top-t records per step, followed by the same merge everywhere.

Prints ONE JSON line (rank 0).  `value` = evals/s with inputs resident in HBM (CUDA-event
timed); `e2e` = the same pass through the host-buffer C ABI (bamse_upload_problem +
bamse_score_topk: H2D of the read-count tables and constants, D2H of the records inside the
timed region); `roofline` = the enumeration kernel against its binding SM pipe;
`cpu_baseline` = the reference's own get_tree_score on this box's host cores (bounded sample).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tree-likelihood evals/sec (trees x clusterings)"
UNIT = "evals/s"


# ------------------------------------------------------------------------------------------
# algorithmic operation counts
def mean_leaves(K):
    return K * (1.0 - 1.0 / K) ** (K - 1) if K > 1 else 1.0


def algorithmic_ops_per_eval(K, M, L=512):
    """(conv terms, transcendental ops, fp32 flops) averaged over all labeled rooted trees."""
    lm1 = mean_leaves(K) - 1.0
    n_c = M * lm1 * (L * (L + 1) / 2.0)
    n_s = M * K * L
    xu = n_c + M * lm1 * L + 2 * n_s + M * L + M
    flops = 4 * n_c + 4 * n_s + 2 * M * K * L
    return n_c, xu, flops


# ------------------------------------------------------------------------------------------
class ClockSampler(object):
    """samples nvidia-smi clocks / throttle reasons; start() early (nvidia-smi needs ~0.3 s to
    produce its first row), then mark the timed regions with begin()/end(): only rows whose
    timestamp falls inside a marked region are summarised."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.windows = []
        self._t0 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def begin(self):
        self._t0 = time.time()

    def end(self):
        if self._t0 is not None:
            self.windows.append((self._t0, time.time()))
            self._t0 = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)                     # let the last rows of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power, sm_all = [], [], set(), [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t_arrival, r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                # rows are stamped by nvidia-smi; parse its local-time stamp, fall back to arrival time
                try:
                    ts = time.mktime(time.strptime(f[0][:19], "%Y/%m/%d %H:%M:%S")) + float("0" + f[0][19:])
                except Exception:
                    ts = t_arrival
                c, m, pw = float(f[1]), float(f[2]), float(f[3])
            except ValueError:
                continue
            sm_all.append(c)
            if not any(lo - 0.03 <= ts <= hi + 0.03 for lo, hi in self.windows):
                continue
            sm.append(c)
            mx.append(m)
            power.append(pw)
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "samples_total": len(sm_all),
                "reasons": sorted(reasons), "regions": "resident + e2e timed regions"}


# ------------------------------------------------------------------------------------------
def hbm_evidence(algorithmic_bytes, traffic, kernel_s):
    """Why the roofline is not an HBM one: DRAM bytes of the kernel (committed ncu capture) over its measured
    duration, against the measured copy bandwidth of this pool's B200s (MEASURED_PEAKS.json, else the
    profiling guide's fallback)."""
    out = {"algorithmic_bytes_per_launch": algorithmic_bytes,
           "note": "inputs are KBs per step; the DRAM traffic is the derived tables (written once, read by the next level "
                   "and the enumeration), a few percent of the HBM peak: the XU pipe binds"}
    peak, src = 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        pass
    if traffic and kernel_s > 0:
        gbs = traffic / kernel_s / 1e9
        out.update({"dram_gbs": gbs, "hbm_peak_gbs": peak, "hbm_peak_source": src, "hbm_frac": gbs / peak})
    return out


def pipe_peaks():
    """measured SM-pipe peaks of this pool's B200 (profiles/pipe_peaks.json, produced by
    profiles/pipe_peaks.cu under gpurun); nominal fallback = 148 SMs x lanes x 1.965 GHz."""
    path = os.path.join(ROOT, "profiles", "pipe_peaks.json")
    nominal = {"fp32_tflops": 148 * 128 * 2 * 1.965e9 / 1e12, "xu_tops": 148 * 16 * 1.965e9 / 1e12,
               "fp64_tflops": 148 * 64 * 2 * 1.965e9 / 1e12, "source": "nominal (148 SM x lanes x 1.965 GHz)"}
    try:
        with open(path) as f:
            d = json.load(f)
        d.setdefault("source", "measured (profiles/pipe_peaks.json)")
        for k, v in nominal.items():
            d.setdefault(k, v)
        return d
    except Exception:
        return nominal


# ------------------------------------------------------------------------------------------
def reference_eval_worker(args):
    """scores a few work items with the reference's own clustering.get_tree_score (or the
    numpy port when oracle/_ref is absent).  Runs in a worker process."""
    kind, As_list, Bs_list, err, items = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    t0 = time.perf_counter()
    if kind == "reference":
        import ref_shim
        cl = ref_shim.load().clustering
        for ci, tree in items:
            As, Bs = np.array(As_list[ci]), np.array(Bs_list[ci])
            K, M = As.shape
            D = np.array([[cl.distrib(As[k, m], Bs[k, m], err) for m in range(M)] for k in range(K)])
            P = np.array([[cl.prior_fractions(K) for m in range(M)] for k in range(K)])
            float(np.sum(cl.tree_integrate_multisample(np.array(tree), D, P)))
    else:
        import bamse_oracle as orc
        for ci, tree in items:
            As, Bs = np.array(As_list[ci]), np.array(Bs_list[ci])
            K, M = As.shape
            D = np.array([[orc.distrib(As[k, m], Bs[k, m], err) for m in range(M)] for k in range(K)])
            float(orc.tree_score_samples(tree, D, orc.prior_const(K)).sum())
    return len(items), time.perf_counter() - t0


def cpu_reference_rate(work, n_items, cores, seed=0):
    """evals/s of the reference CPU path on `cores` worker processes over a seeded random
    sample of `n_items` (clustering, tree) work items of this workload."""
    import multiprocessing as mp
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bamse_oracle as orc
    import ref_shim
    kind = "reference" if ref_shim.available() else "port"
    rng = np.random.default_rng(seed)
    clusts = work["clusts"]
    weights = np.array([float(int(c.As.shape[0]) ** (int(c.As.shape[0]) - 1)) for c in clusts])
    picks = rng.choice(len(clusts), size=n_items, p=weights / weights.sum())
    items = []
    for ci in picks:
        K = int(clusts[ci].As.shape[0])
        items.append((int(ci), orc.decode_tree(K, int(rng.integers(0, K ** (K - 1))))))
    As_list = [c.As.tolist() for c in clusts]
    Bs_list = [c.Bs.tolist() for c in clusts]
    chunks = [items[i::cores] for i in range(cores)]
    chunks = [c for c in chunks if c]
    t0 = time.perf_counter()
    if cores == 1:
        res = [reference_eval_worker((kind, As_list, Bs_list, work["errs"][0], chunks[0]))]
    else:
        with mp.get_context("spawn").Pool(len(chunks)) as pool:
            res = pool.map(reference_eval_worker, [(kind, As_list, Bs_list, work["errs"][0], c) for c in chunks])
    wall = time.perf_counter() - t0
    busy = max(r[1] for r in res)   # slowest worker, excludes process start-up
    n = sum(r[0] for r in res)
    per_core = statistics.mean(r[0] / r[1] for r in res)     # the reference is single-threaded: its rate on one core
    return dict(value=n / busy, unit=UNIT, cores=len(chunks), kind=kind, wall_s=wall, per_core=per_core,
                single_process_evals_per_s=per_core, cpu_model=cpu_model(),
                sample="%d seeded random (clustering, tree) items of %s, %d worker processes" % (n, work["name"], len(chunks)))


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation on this box's host cores."""
    if rank != 0:
        return
    from bamse_b200 import workloads
    work = workloads.build(args.config)
    cores = os.cpu_count() or 1
    per_step = args.ref_items if args.ref_items > 0 else cores
    rates = []
    for _ in range(args.warmup):
        cpu_reference_rate(work, max(1, cores), cores, seed=99)
    t0 = time.perf_counter()
    n_done = 0
    for s in range(args.steps):
        r = cpu_reference_rate(work, per_step, cores, seed=s)
        rates.append(r)
        n_done += per_step
    wall = time.perf_counter() - t0
    value = statistics.mean(r["value"] for r in rates)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, work, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": rates[0]["cores"], "kind": rates[0]["kind"],
                             "sample": "%d steps x %s" % (args.steps, rates[0]["sample"])},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_other_config(name, args):
    """A second configuration through this same script in a child process (one GPU, same build, same box, seconds
    later), summarised under `also_measured`: with c3 as the default workload the line still carries c2, the default
    of round 1.  Never fatal: a failing child is reported as an error string."""
    try:
        cmd = [sys.executable, os.path.abspath(__file__), "--gpus", "1", "--config", name, "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--precision", args.precision, "--topk", str(args.topk),
               "--no-cpu-baseline", "--also", ""]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if out.returncode != 0 or not lines:
            return {"config": name, "error": ("rc %d: " % out.returncode) + out.stderr.strip()[-300:]}
        d = json.loads(lines[-1])
        ex = (d.get("roofline") or {}).get("executed") or {}
        return {"config": d["config"]["workload"], "evals_per_step": d["config"]["evals_per_step"], "value": d["value"],
                "unit": d["unit"], "ms_per_step": d["ms_per_step"], "steps": d["steps"], "warmup": d["warmup"],
                "e2e": {k: d["e2e"].get(k) for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step",
                                                     "cold_single_shot_ms")},
                "gpu_launches": d.get("gpu_launches"), "table_build_ms": ex.get("table_build_ms"),
                "enumeration_ms": ex.get("enumeration_ms"), "xu_pipe_frac": ex.get("xu_pipe_frac"),
                "clocks": d.get("clocks"), "note": "not the headline: `python bench.py --config %s` in a child process" % name}
    except Exception as e:                                   # noqa: BLE001 -- an extra must never cost the headline
        return {"config": name, "error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def config_dict(args, work, world):
    import collections
    ks = collections.Counter(int(c.As.shape[0]) for c in work["clusts"])
    return {"workload": "%s: simdata.py synthetic, %s; all K^(K-1) labeled rooted trees of every candidate clustering"
                        % (work["name"], work["desc"]),
            "clusterings": len(work["clusts"]), "K_histogram": {str(k): v for k, v in sorted(ks.items())},
            "param_sets": len(work["errs"]), "samples": work["M"],
            "evals_per_step": work["n_evals_per_param"] * len(work["errs"]),
            "parallelism": "x%d: whole (parameter set, clustering) pairs per rank by estimated cost, the most expensive "
                           "ones split over all ranks (by pack node sets on the pair path, by interleaved tree index "
                           "otherwise); one all-gather of the top-k records" % world,
            "l2": "256 MiB write between steps inside the timed region; every step rebuilds its tables (GBs) from the counts"}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--topk", type=int, default=5)
    ap.add_argument("--ref-items", type=int, default=0,
                    help="work items per step of the reference arm (default: one per host core, so a step lasts one item -- "
                         "about 3 s on c2 and 14 s on c3 -- whatever the core count)")
    ap.add_argument("--cpu-items", type=int, default=0,
                    help="sample size of the cpu_baseline leg; default: 200 items (SURVEY 8d) where an item costs a few "
                         "core-seconds (c1, c2, c5), 96 on c3 (~14 core-seconds per item: 20 samples, 7 clusters), 32 on c4 "
                         "-- about a minute of wall time on 16 cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--also", default=None, choices=["", "c1", "c2", "c3", "c5"],
                    help="one more configuration, measured by a child process after the headline and summarised under "
                         "`also_measured` (one GPU only); default: c2 when the workload is c3, none otherwise; '' = none")
    ap.add_argument("--search-only", action="store_true",
                    help="not a bench line: run ONE pass of the early-exit search (for configurations whose full\n"
                         "scoring pass is too long, e.g. c4 = 3e9 trees) and print its own JSON")
    ap.add_argument("--lean", action="store_true",
                    help="long configurations (c4): skip the extra legs (enumeration-only, early-exit search) and time "
                         "ONE end-to-end step without its own warm-up")
    ap.add_argument("--emulate-shard-of", type=int, default=0,
                    help="debug: one GPU scores only shard 0 of W (what each rank of a W-GPU run does)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.also is None:
        args.also = "c2" if args.config == "c3" else ""
    if args.cpu_items <= 0:
        args.cpu_items = {"c3": 96, "c4": 32}.get(args.config, 200)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from bamse_b200 import cuda_shim, workloads
    from bamse_b200.clustering import _pack_items, _pack_problem
    from bamse_b200.tree_tools import canonizeF

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    prec = cuda_shim.FP32 if args.precision == "fp32" else cuda_shim.FP64

    ctx = cuda_shim.Context(local_rank)
    work = workloads.build(args.config, ctx=ctx)     # identical on every rank (seeded)
    # largest tree spaces first, so that the cheap clusterings fill the tail of the launch
    work["clusts"].sort(key=lambda c: -int(c.As.shape[0]))
    clusts, errs = work["clusts"], work["errs"]
    P, C = len(errs), len(clusts)

    # ---- untimed preparation = the callers of the path (bamse.py:95-97): reorder, greedy
    #      candidate, threshold -- batched GPU list-scoring calls, fp64 check mode.
    Ks, var, ref, Kmax = _pack_problem(clusts)
    ctx.upload_problem([errs[0]], Ks, var, ref)
    items, spans = [], []
    for ci, c in enumerate(clusts):
        pi = c._pair_items()
        spans.append((len(items), len(pi)))
        items += [(ci, t, n) for t, n in pi]
    if items:
        s = ctx.score_list(*_pack_items(Kmax, items), cuda_shim.FP64)[0]
        for c, (o, n) in zip(clusts, spans):
            c._apply_order(s[o:o + n] if n else [])
    Ks, var, ref, Kmax = _pack_problem(clusts)
    # ---- shard BEFORE the upload: the library plans which rank owns which (parameter set, clustering) -- whole
    #      pairs where it can, the most expensive ones split (pack node sets / interleaved tree indices) -- and builds tables only
    #      for what this rank scores (bamse_set_shard + BAMSE_OPT_SHARD_PLAN)
    ctx.set_shard(rank, world)
    if args.emulate_shard_of > 1 and world == 1:
        ctx.set_shard(0, args.emulate_shard_of)
    ctx.upload_problem(errs, Ks, var, ref)            # all P parameter sets resident from here on
    cand = [[[-1] for _ in clusts] for _ in range(P)]
    for n in range(1, Kmax):
        items, owners = [], []
        for ci, c in enumerate(clusts):
            if c.As.shape[0] > n:
                # candidates depend on e: score each p's own partial tree (P x n items per clustering)
                for p in range(P):
                    owners.append((p, ci, len(items)))
                    items += [(ci, cand[p][ci] + [x], list(range(n + 1))) for x in range(n)]
        sc = ctx.score_list(*_pack_items(Kmax, items), cuda_shim.FP64)
        for p, ci, o in owners:
            cand[p][ci] = cand[p][ci] + [int(np.argmax(sc[p, o:o + n]))]
    items = [(ci, cand[p][ci], list(range(int(Ks[ci])))) for p in range(P) for ci in range(C)]
    sc = ctx.score_list(*_pack_items(Kmax, items), cuda_shim.FP64)
    thr = np.array([[sc[p, p * C + ci] + math.log(canonizeF(cand[p][ci])['scre']) for ci in range(C)] for p in range(P)])
    const = np.array([[c.cluster_prior + c.maxlikelihood for c in clusts] for _ in range(P)])

    tb = te = None
    total_evals = work["n_evals_per_param"] * P
    topk = args.topk

    # ---- cold single shot: a FRESH context (what one command-line run pays): data-independent pair tables /
    #      recipes of every cluster count present + upload + tables + enumeration + fp64 re-score, timed once
    cold_ms = None
    if not args.lean:
        torch.cuda.synchronize()
        t_cold = time.perf_counter()
        cold = cuda_shim.Context(local_rank)
        cold.set_shard(rank, world)
        cold.upload_problem(errs, Ks, var, ref)
        cold.score_topk(const, thr, None, None, topk=topk, precision=prec)
        cold_ms = 1e3 * (time.perf_counter() - t_cold)     # context creation -> ranked records in host memory
        cold.close()                                        # (freeing GBs of tables is not part of a run's latency)

    # ---- device-resident state for the `value` measurement
    # a real (non-default) torch stream: handle 0 would mean "the ctx's own stream" to the ABI
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    d_const = torch.from_numpy(const).to(dev)
    d_thr = torch.from_numpy(thr).to(dev)
    rec_words = topk * P * 4                         # a bamse_record = 4 x 8 bytes
    d_local = torch.zeros(rec_words, dtype=torch.int64, device=dev)
    d_all = torch.zeros(world * rec_words, dtype=torch.int64, device=dev)
    d_final = torch.zeros(rec_words, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # fp32 selection keeps trees within `slack` nats below the fp64 threshold (the host API re-scores in fp64)
    slack = 0.0 if prec == cuda_shim.FP64 else 1e-3

    def step_resident(tables=True):
        flush.fill_(1)
        if tables:
            # every derived table of the problem from the read counts resident in HBM: part of the work per
            # problem (the sub-forest / context tables ARE the memoised convolutions), so inside the timed region
            ctx.rebuild_tables()
        ctx.enumerate_topk_dev(d_const.data_ptr(), d_thr.data_ptr(), None, None, topk, prec, slack,
                               d_local.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_local)
            ctx.merge_topk_dev(d_all.data_ptr(), world, P, topk, d_final.data_ptr())
        else:
            d_final.copy_(d_local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.search_only:
        ctx.set_early_exit(True)
        ctx.samples_evaluated(reset=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_resident()
        e1.record()
        barrier()
        ts = torch.tensor([e0.elapsed_time(e1), float(ctx.samples_evaluated(reset=True))], dtype=torch.float64, device=dev)
        if world > 1:
            tm = ts.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            ts[0] = tm[0]
        final = d_final.cpu().numpy().view(cuda_shim.RECORD_DTYPE).reshape(P, topk)
        if rank == 0:
            print(json.dumps({"not_a_bench_line": "early-exit search, one pass", "config": config_dict(args, work, world),
                              "n_gpus": world, "seconds": float(ts[0]) / 1e3,
                              "trees_resolved_per_s": total_evals / (float(ts[0]) / 1e3),
                              "tree_samples_evaluated_frac": float(ts[1]) / (total_evals * work["M"]),
                              "top": [{"total": float(final["total"][0, j]), "clust": int(final["clust"][0, j]),
                                       "tree": int(final["tree"][0, j])} for j in range(topk)]}), flush=True)
        ctx.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    ctx.kernel_time(reset=True)
    ctx.table_time(reset=True)
    ctx.kernel_stats(reset=True)
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    sampler.end()
    ms = ev0.elapsed_time(ev1)
    kern_ms, kern_n = ctx.kernel_time(reset=True)
    tab_ms, tab_n = ctx.table_time(reset=True)
    ctx.kernel_stats(reset=True)
    variant = ctx.fast_variant
    launches = ctx.launch_count - launches0
    t = torch.tensor([ms, kern_ms / max(1, kern_n), tab_ms / max(1, tab_n)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, kern_ms_max, tab_ms_max = float(t[0]), float(t[1]), float(t[2])
    final = d_final.cpu().numpy().view(cuda_shim.RECORD_DTYPE).reshape(P, topk)

    # ---- executed-work counters of one (untimed) step: table builders + enumeration kernels of this rank
    ctx.set_option(cuda_shim.OPT_STATS, 1)
    step_resident()
    barrier()
    stats = ctx.kernel_stats(reset=True)
    # ... and of the enumeration alone (one more untimed step over the tables just built): the difference is the
    #     table builders' share, so each kernel group gets its own executed XU-pipe fraction
    step_resident(tables=False)
    barrier()
    stats_enum = ctx.kernel_stats(reset=True)
    ctx.set_option(cuda_shim.OPT_STATS, 0)
    ctx.kernel_time(reset=True)
    ctx.table_time(reset=True)

    # ---- extra (not the headline): enumeration alone over tables that stay resident between steps
    enum_only = None
    if not args.lean:
        for _ in range(2):
            step_resident(tables=False)
        barrier()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        for _ in range(args.steps):
            step_resident(tables=False)
        ee1.record()
        barrier()
        t_en = torch.tensor([ee0.elapsed_time(ee1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_en, op=dist.ReduceOp.MAX)
        enum_only = {"value": total_evals * args.steps / (float(t_en[0]) / 1e3), "unit": UNIT,
                     "ms_per_step": float(t_en[0]) / args.steps,
                     "note": "tables kept from the previous step (NOT the headline: a new problem always pays its tables)"}
        ctx.kernel_time(reset=True)

    # ---- extra (not the headline): the same resident pass with the exact sample-level early exit
    #      (BAMSE_OPT_EARLY_EXIT) -- trees are resolved, not all of their samples are scored
    search = None
    if prec == cuda_shim.FP32 and not args.lean:
        ctx.set_early_exit(True)
        step_resident()
        barrier()
        ctx.samples_evaluated(reset=True)
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es0.record()
        for _ in range(args.steps):
            step_resident()
        es1.record()
        barrier()
        n_samp = ctx.samples_evaluated(reset=True)
        ctx.set_early_exit(False)
        ts = torch.tensor([es0.elapsed_time(es1), float(n_samp)], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = ts.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            ts[0] = tmax[0]
        final_s = d_final.cpu().numpy().view(cuda_shim.RECORD_DTYPE).reshape(P, topk)
        search = {"value": total_evals * args.steps / (float(ts[0]) / 1e3), "unit": "trees resolved/s",
                  "tree_samples_evaluated_frac": float(ts[1]) / (total_evals * work["M"] * args.steps),
                  "same_best_tree": bool(int(final_s["tree"][0, 0]) == int(final["tree"][0, 0])
                                         and int(final_s["clust"][0, 0]) == int(final["clust"][0, 0])),
                  "note": "optional exact early exit over samples (per-sample scores are <= 0); NOT the headline: "
                          "`value` scores every sample of every tree"}
        ctx.kernel_time(reset=True)                          # keep the kernel timer clean for later legs

    # ---- e2e: the host-buffer C ABI, H2D + D2H inside the timed region
    ctx.set_stream(None)
    h2d = var.nbytes + ref.nbytes + Ks.nbytes + 8 * P + const.nbytes + thr.nbytes
    d2h = topk * P * 32 if prec == cuda_shim.FP64 else min(64, topk + 8) * P * 32 * 2
    var_pin = torch.from_numpy(var).pin_memory().numpy()
    ref_pin = torch.from_numpy(ref).pin_memory().numpy()

    def step_e2e():
        ctx.upload_problem(errs, Ks, var_pin, ref_pin)
        rec, par, cnt = ctx.score_topk(const, thr, tb, te, topk=topk, precision=prec)
        if world > 1:
            loc = torch.from_numpy(rec.view(np.int64).reshape(-1).copy()).to(dev)
            dist.all_gather_into_tensor(d_all, loc)
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            ctx.merge_topk_dev(d_all.data_ptr(), world, P, topk, d_final.data_ptr())
            out = d_final.cpu()
            ctx.set_stream(None)
            return out.numpy().view(cuda_shim.RECORD_DTYPE).reshape(P, topk)
        return rec

    for _ in range(0 if args.lean else min(args.warmup, 2)):
        step_e2e()
    barrier()
    e2e_steps = 1 if args.lean else args.steps
    sampler.begin()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_rec = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    te2e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te2e, op=dist.ReduceOp.MAX)
    e2e_s = float(te2e[0])

    if rank == 0:
        # sanity: the resident path and the host path agree on the winning trees
        same = (int(final["tree"][0, 0]), int(final["clust"][0, 0])) == (int(e2e_rec["tree"][0, 0]), int(e2e_rec["clust"][0, 0]))
        value = total_evals * args.steps / (ms_max / 1e3)
        e2e_value = total_evals * e2e_steps / e2e_s
        # roofline of the enumeration kernel (per launch, this rank's share of the trees)
        peaks = pipe_peaks()
        terms = xu = flops = 0.0
        for k in Ks:
            n_c, x, f = algorithmic_ops_per_eval(int(k), work["M"])
            share = (int(k) ** (int(k) - 1)) / world * P
            terms += n_c * share
            xu += x * share
            flops += f * share
        # the step's own kernels: table builders (k_tables*, k_forest_level, k_ctx_level) + enumeration
        # (k_enumerate_pairs, k_enumerate_fast); both groups are event-timed inside the library
        kernel_s = (kern_ms_max + tab_ms_max) / 1e3
        if args.precision == "fp64":
            kernel_s = kern_ms_max / 1e3
            bound, achieved, peak, unit = "fp64-pipe", 40 * terms / kernel_s / 1e12, peaks["fp64_tflops"], "TFLOP/s"
            kname, executed = "k_enumerate_logdomain<double>", None
        else:
            bound, achieved, peak, unit = "xu (MUFU ex2/lg2)", xu / kernel_s / 1e12, peaks["xu_tops"], "Tops/s"
            kname = "table builders (k_forest_level, k_ctx_level) + k_enumerate_pairs / k_enumerate_fast"
            ex = stats["mufu_ops"] / kernel_s / 1e12     # this rank's executed MUFU ops of one step per second
            mufu_enum = min(stats_enum["mufu_ops"], stats["mufu_ops"])
            mufu_tab = stats["mufu_ops"] - mufu_enum
            groups = {
                "table_builders": {"kernels": "k_tables_fast, k_forest_level, k_ctx_level", "ms_per_step": tab_ms_max,
                                   "mufu_ops_per_step": mufu_tab, "share_of_kernel_time": tab_ms_max / (1e3 * kernel_s),
                                   "xu_pipe_frac": (mufu_tab / (tab_ms_max / 1e3) / 1e12 / peaks["xu_tops"]) if tab_ms_max > 0 else None},
                "enumeration": {"kernels": "k_enumerate_pairs, k_enumerate_fast", "ms_per_step": kern_ms_max,
                                "mufu_ops_per_step": mufu_enum, "share_of_kernel_time": kern_ms_max / (1e3 * kernel_s),
                                "xu_pipe_frac": (mufu_enum / (kern_ms_max / 1e3) / 1e12 / peaks["xu_tops"]) if kern_ms_max > 0 else None,
                                "bound": "L1 wavefronts of the windowed table reads (profiles/r02_k_enumerate_pairs_*.txt), not the XU pipe"}}
            executed = {"xu_tops": ex, "xu_pipe_frac": ex / peaks["xu_tops"], "by_kernel_group": groups,
                        "mufu_ops_per_step": stats["mufu_ops"], "convolutions_per_step": stats["convs"],
                        "conv_terms_executed_per_step": stats["conv_terms"],
                        "conv_terms_algorithmic_per_step": terms,
                        "table_build_ms": tab_ms_max, "enumeration_ms": kern_ms_max,
                        "note": "achieved/frac use the ALGORITHMIC op count (SURVEY 8d: every term of every "
                                "convolution of every tree); sub-forest / context tables memoise the convolutions "
                                "across trees and only terms within 2^-26 of an output's maximum are exponentiated, "
                                "so frac exceeds 1 by orders of magnitude; xu_pipe_frac is the XU pipe utilisation "
                                "of what actually executed (table builders + enumeration kernels, event-timed)"}
        # DRAM traffic of the step's table-building kernels (1 launch of each per step), from the committed
        # ncu --set full capture of this configuration (profiles/r02_traffic.json); null for other configurations
        traffic, traffic_detail, traffic_s = None, None, kernel_s
        for tname in ("r02_traffic_%s.json" % args.config, "r02_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", tname)) as f:
                    tj = json.load(f)
            except Exception:
                continue
            if args.config == tj.get("config") and world == tj.get("n_gpus", 1) and args.precision == "fp32":
                traffic = int(tj["dram_bytes_read"] + tj["dram_bytes_write"])
                traffic_detail = {"dominant_kernel": tj["dominant_kernel"], "per_kernel": tj["kernels"], "source": tj["source"],
                                  "scope": tj.get("scope", "every kernel of one step")}
                if tj.get("kernel_time") == "table_build" and tab_ms_max > 0:
                    traffic_s = tab_ms_max / 1e3       # the capture covers the table builders only
                break
        roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peaks["source"], "kernel": kname,
                    "kernel_ms_per_launch": 1e3 * kernel_s, "executed": executed,
                    "hbm": hbm_evidence(int(h2d), traffic, traffic_s), "traffic_detail": traffic_detail}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32" if prec == cuda_shim.FP32 else "f64",
                "data": "synthetic", "config": config_dict(args, work, world), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": 1e3 * e2e_s / e2e_steps, "agrees_with_resident_path": bool(same),
                        "cold_single_shot_ms": cold_ms},
                "gpu_launches": int(launches), "roofline": roofline, "search_mode_extra": search,
                "enumeration_only_extra": enum_only,
                "best": {"total": float(final["total"][0, 0]), "clust": int(final["clust"][0, 0]),
                         "tree": int(final["tree"][0, 0])}}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_rate(work, args.cpu_items, os.cpu_count() or 1)
    ctx.close()
    if rank == 0:
        if world == 1 and args.also and args.also != args.config and not args.lean:
            line["also_measured"] = run_other_config(args.also, args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
