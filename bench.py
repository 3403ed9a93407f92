#!/usr/bin/env python3
"""bench.py — index-build throughput of arroy's distance / split hot path on B200.

One "step" = one complete forest build (Writer::build of the reference, src/writer.rs:487-629)
over one synthetic item matrix. Contract: `python bench.py --gpus N --steps K --warmup W`
prints ONE JSON line (rank 0). `--impl reference` times the reference's CPU path instead (the
Rust crate cannot be compiled in this image, so it is the C++ restatement in oracle/, kind
"port": same AVX+FMA kernels, one task per tree on all host cores — see BASELINE.md §3).

Workloads (BASELINE.json configs; SURVEY.md §8d synthetic data: element (i,j) = gen::<f32>()
number i*d+j of StdRng::from_seed([42;32]) minus 0.5; build rng = fresh StdRng([42;32])):
  c2 (default)  1 000 000 x 768  Cosine      n_trees = 50    <- BASELINE.json configs[1]
  c3            10 000 000 x 768 DotProduct  n_trees = 100
  c4            10 000 000 x 1536 Cosine     n_trees = 100   (meant for 8 GPUs)
  c1            10 000 x 64      Euclidean   n_trees = 10    (raw [0,1) data)
  small         100 000 x 768    Cosine      n_trees = 16    (quick check)
  c5            4096 queries x 100 000 shared candidates, d = 768, Cosine, top-100: the batched
                re-rank (reader.rs:381-399) as a tcgen05 TF32 pre-filter + exact re-score; its own
                metric (queries/s) and a "tensor" roofline for the score contraction

Keys beyond the base contract: "roofline" (dominant kernel = the side()/margin scan inside
work_kernel), "cpu_baseline", "e2e" (through Writer.builder(rng).build() with the items as host
leaf values: decode + H2D + device build + D2H + NodeCodec encoding + metadata, all timed),
"clocks", "gpu_launches".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = bytes([42] * 32)
WORKLOADS = {
    "c2": dict(n=1_000_000, d=768, metric="cosine", n_trees=50, centre=0.5, name="C2 1Mx768 Cosine n_trees=50"),
    "c3": dict(n=10_000_000, d=768, metric="dot-product", n_trees=100, centre=0.5, name="C3 10Mx768 DotProduct n_trees=100"),
    "c4": dict(n=10_000_000, d=1536, metric="cosine", n_trees=100, centre=0.5, name="C4 10Mx1536 Cosine n_trees=100"),
    "c1": dict(n=10_000, d=64, metric="euclidean", n_trees=10, centre=0.0, name="C1 10kx64 Euclidean n_trees=10"),
    "small": dict(n=100_000, d=768, metric="cosine", n_trees=16, centre=0.5, name="small 100kx768 Cosine n_trees=16"),
    "c5": dict(n=100_000, d=768, metric="cosine", n_trees=0, centre=0.5, nq=4096, k=100, name="C5 4096 queries x 100k candidates re-rank, d=768 Cosine top-100"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def derive_seeds(ab, n_trees):
    """seed chain of Writer::build (src/writer.rs:575, :795) with the product's own StdRng."""
    import numpy as np
    user = ab.StdRng.from_seed(SEED)
    s1 = bytes(np.array([user.next_u32() & 0xff for _ in range(32)], dtype=np.uint8))
    r1 = ab.StdRng.from_seed(s1)
    return [bytes(np.array([r1.next_u32() & 0xff for _ in range(32)], dtype=np.uint8)) for _ in range(n_trees)]


def run_reference(args, wl):
    """The reference's CPU path (oracle port) on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import oracle
    oracle.build_lib()
    cores = os.cpu_count() or 1
    n, d, T = wl["n"], wl["d"], wl["n_trees"]
    # bounded sample: the same item matrix, as many trees as keep one step within ~tens of seconds
    t_sample = T if args.ref_trees is None else args.ref_trees
    if args.ref_trees is None and n * d >= 5e8:
        t_sample = min(T, max(8, cores // 2))
    data = oracle.synth_rows(SEED, d, 0, n, wl["centre"], threads=min(cores, 64))
    ids = np.arange(n, dtype=np.uint32)
    times = []
    scanned = 0
    for step in range(args.warmup + args.steps):
        db = oracle.Db(wl["metric"], d)
        db.set_items(ids, data)
        rng = oracle.StdRng(SEED)
        t0 = time.perf_counter()
        ref_threads = 1 if args.workload == "c1" else min(cores, t_sample)   # configs[0]: single-thread CPU reference
        db.build(rng, n_trees=t_sample, threads=ref_threads)
        dt = time.perf_counter() - t0
        scanned = db.scanned_rows
        if step >= args.warmup:
            times.append(dt)
        del db
    sec = sum(times) / len(times)
    # vectors/s of the FULL forest, extrapolated linearly in the number of trees (trees are independent)
    value = n / (sec * T / t_sample)
    line = {
        "impl": "reference", "metric": "index-build vectors/sec", "value": value, "unit": "vectors/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": wl["name"], "n": n, "d": d, "distance": wl["metric"], "n_trees": T},
        "cpu_baseline": {"value": value, "unit": "vectors/s", "cores": ref_threads, "kind": "port",
                         "sample": "%d of %d trees over the full %dx%d matrix, %d thread(s) (one tree per thread), extrapolated x%.2f" % (t_sample, T, n, d, ref_threads, T / t_sample),
                         "scan_GBps": scanned * d * 4 / sec / 1e9},
        "e2e": {"value": value, "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



def tensor_peak():
    """TF32 dense peak: half the measured dense bf16 rate (tcgen05 kind::tf32 runs at half the kind::f16 rate)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if "bf16_tflops_sustained" in d:
            return d["bf16_tflops_sustained"] / 2.0, "measured bf16 sustained %.1f TFLOP/s / 2 (TF32 rate; MEASURED_PEAKS.json)" % d["bf16_tflops_sustained"]
    return 1100.0, "fallback: nominal dense TF32 (B200_PROFILING.md)"


def c5_cpu(wl, data_host, qh, n_queries, threads):
    """The reference's re-rank loop (oracle port) for a few queries, one query per thread."""
    import numpy as np
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    n, nq, k = wl["n"], wl["nq"], wl["k"]
    rows = np.arange(n, dtype=np.uint32)
    hdr = qh
    m = oracle.METRICS[wl["metric"]]

    def one(i):
        return oracle.rerank(m, data_host[n + i], (float(hdr[n + i]), 0.0), data_host, hdr, None, rows, k)
    one(0)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(one, range(n_queries)))
    return n_queries / (time.perf_counter() - t0), res


def run_c5(args, wl):
    """BASELINE.json configs[4]: batched 4096-query x 100k-candidate re-rank, d = 768, 1 GPU (replicas for N > 1)."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n, d, nq, k, metric = wl["n"], wl["d"], wl["nq"], wl["k"], wl["metric"]
    cores = os.cpu_count() or 1
    if args.impl == "reference":
        if rank != 0:
            return
        import oracle
        oracle.build_lib()
        data = oracle.synth_rows(SEED, d, 0, n + nq, wl["centre"], threads=min(cores, 64))
        hdr = np.sqrt((data.astype(np.float64) ** 2).sum(1)).astype(np.float32)   # only the timing matters here
        sample = min(nq, max(2 * cores, 64))
        qps = []
        for step in range(args.warmup + args.steps):
            v, _ = c5_cpu(wl, data, hdr, sample, cores)
            if step >= args.warmup:
                qps.append(v)
        value = sum(qps) / len(qps)
        print(json.dumps({"impl": "reference", "metric": "batched re-rank queries/sec", "value": value, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": sample / value * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": wl["name"], "n_candidates": n, "n_queries": nq, "d": d, "distance": metric, "k": k},
                          "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": "%d of %d queries per step, one query per thread" % (sample, nq)},
                          "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import torch
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import arroy_b200 as ab
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = ab.Context(local_rank)
    items = torch.empty((n + nq, d), dtype=torch.float32, device=dev)
    ctx.synth_device(SEED, d, 0, n + nq, wl["centre"], items.data_ptr())
    ctx.stage_items_device(metric, np.arange(n + nq, dtype=np.uint32), d, items.data_ptr())
    h0, _ = ctx.item_headers()
    q_host = items[n:].cpu().numpy()          # queries = the rows that continue the stream after the candidates
    qh = np.ascontiguousarray(h0[n:])
    rows = np.arange(n, dtype=np.uint32)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    out = None
    for _ in range(max(args.warmup, 1)):
        out = ctx.rerank_shared(q_host, qh, rows, k)
    c0 = ctx.counters()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ctx.timer_start()
    t0 = time.perf_counter()
    gemm_ms = []
    for _ in range(args.steps):
        out = ctx.rerank_shared(q_host, qh, rows, k)
        gemm_ms.append(ctx.rerank_breakdown())
    dev_ms = ctx.timer_stop()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    c1 = ctx.counters()
    t_ms = torch.tensor([max(dev_ms, 0.0)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    value = world * nq / (ms_per_step * 1e-3)
    bd = {kk: sum(b[kk] for b in gemm_ms) / len(gemm_ms) for kk in gemm_ms[0]}
    stats = ctx.rerank_stats()
    peak, peak_src = tensor_peak()
    flop = 2.0 * nq * n * d
    line = {
        "metric": "batched re-rank queries/sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (tf32 tensor-core pre-filter, exact f32 re-score)", "data": "synthetic",
        "config": {"workload": wl["name"], "n_candidates": n, "n_queries": nq, "d": d, "distance": metric, "k": k, "parallelism": "replicas only (re-rank is single-GPU)",
                   "l2": "score matrix (1.6 GB) and candidates (307 MB) larger than L2; no flush needed",
                   "timing": "CUDA events on the library stream around the C-ABI call (host query / result buffers, copies included), max over ranks"},
        "gpu_launches": int(c1["launches"] - c0["launches"]),
        "rerank": {"breakdown_ms": bd, "survivors_per_query": stats["survivors"] / max(stats["queries"], 1), "fallback_chunks": stats["fallback_chunks"],
                   "exact_pairs_per_s": nq * n / (ms_per_step * 1e-3)},
        "roofline": {"bound": "tensor", "kernel": "tcgemm_tf32_kernel (tcgen05.mma kind::tf32 + TMA + TMEM, fused distance-estimate epilogue)",
                     "achieved": flop / (bd["score_gemm_ms"] * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": flop / (bd["score_gemm_ms"] * 1e-3) / 1e12 / peak,
                     "peak_source": peak_src, "traffic": 1.919e9,
                     "traffic_note": "dram read 0.330 GB + write 1.589 GB per launch (profiles/r01_c5_prefilter_ncu_raw.csv); operand traffic L2->SM is 14.7 GB per launch, the actual limiter"},
        "e2e": {"value": world * nq / (wall / args.steps), "unit": "queries/s", "h2d_bytes_per_step": int((c1["h2d_bytes"] - c0["h2d_bytes"]) / args.steps),
                "d2h_bytes_per_step": int((c1["d2h_bytes"] - c0["d2h_bytes"]) / args.steps), "note": "wall clock around arroy_b200_rerank_shared with pageable host buffers"},
        "clocks": clocks,
    }
    if rank == 0:
        if not args.no_cpu_baseline:
            import oracle
            oracle.build_lib()
            data_host = items.cpu().numpy()
            sample = min(nq, max(cores, 32))
            v, res = c5_cpu(wl, data_host, h0, sample, cores)
            same = all(out[0][i, :out[2][i]].tolist() == res[i][0].tolist() and out[1][i, :out[2][i]].tobytes() == res[i][1].tobytes() for i in range(sample))
            line["cpu_baseline"] = {"value": v, "unit": "queries/s", "cores": cores, "kind": "port", "sample": "%d of %d queries, one query per thread" % (sample, nq),
                                    "results_identical_on_sample": bool(same)}
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("ARROY_BENCH_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--ref-trees", type=int, default=None, help="trees built per step by --impl reference / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-writer-e2e", action="store_true")
    ap.add_argument("--no-query", action="store_true")
    ap.add_argument("--queries", type=int, default=1000)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.workload == "c5":
        run_c5(args, wl)
        return
    if args.impl == "reference":
        run_reference(args, wl)
        return

    import numpy as np
    import torch
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank == 0:
        ge.build()
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import arroy_b200 as ab
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = ab.Context(local_rank)
    n, d, T, metric = wl["n"], wl["d"], wl["n_trees"], wl["metric"]
    ids = np.arange(n, dtype=np.uint32)

    # synthetic items, generated on the device of rank 0 (counter-based ChaCha12 stream)
    items = torch.empty((n, d), dtype=torch.float32, device=dev)
    if rank == 0:
        ctx.synth_device(SEED, d, 0, n, wl["centre"], items.data_ptr())
    torch.cuda.synchronize()
    seeds = derive_seeds(ab, T)
    my_trees = list(range(rank, T, world))  # trees are independent units: tree t -> rank t mod world

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    from arroy_b200 import parallel

    def one_step():
        # multi-GPU: ONE NCCL broadcast of the item buffer over NVLink, then no further data exchange
        if dist is not None:
            parallel.broadcast_items(dist, items, src=0)
            torch.cuda.synchronize()
        ctx.stage_items_device(metric, ids, d, items.data_ptr())
        if metric == "dot-product":
            ctx.dot_preprocess()
        counts = ctx.build_trees_begin([seeds[t] for t in my_trees])
        # tiny all-gather of node counts so every rank can number its nodes like a single-GPU build
        parallel.gather_counts(dist, counts, T, rank, world, device=dev if dist is not None else None)
        return counts

    for _ in range(args.warmup):
        one_step()
    c0 = ctx.counters()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ctx.timer_start()
    t0 = time.perf_counter()
    scanned = 0
    for _ in range(args.steps):
        one_step()
        scanned += ctx.build_stats()["scanned_rows"]
    dev_ms = ctx.timer_stop()   # CUDA events on the library's stream (the launching stream)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    c1 = ctx.counters()
    t_ms = torch.tensor([max(dev_ms, 0.0)], dtype=torch.float64, device=dev)
    sc = torch.tensor([float(scanned)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(sc, op=dist.ReduceOp.SUM)
    ms_per_step = float(t_ms.item()) / args.steps
    value = n / (ms_per_step * 1e-3)
    launches = c1["launches"] - c0["launches"]
    last_stats = ctx.build_stats()
    breakdown = ctx.build_breakdown()

    line = {
        "metric": "index-build vectors/sec", "value": value, "unit": "vectors/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "n": n, "d": d, "distance": metric, "n_trees": T, "parallelism": "trees sharded t mod %d" % world,
                   "l2": "inputs (%.1f GB) larger than L2; no flush needed" % (n * d * 4 / 1e9), "timing": "CUDA events on the library stream, max over ranks",
                   "wall_ms_per_step": wall * 1e3 / args.steps},
        "gpu_launches": int(launches),
        "build": {"scanned_rows_per_step": float(sc.item()) / args.steps, "device_steps": last_stats["steps"], "create_split_calls": last_stats["create_split_calls"],
                  "random_splits": last_stats["random_splits"], "algorithmic_GB_per_step": float(sc.item()) / args.steps * d * 4 / 1e9,
                  "whole_build_GBps": float(sc.item()) / args.steps * d * 4 / 1e9 / (ms_per_step * 1e-3),
                  "schedule": "lockstep" if os.environ.get("ARROY_B200_LOCKSTEP") else "async per-tree graph branches", "breakdown_ms_last_step": breakdown},
    }
    if rank == 0:
        line["clocks"] = clocks
        hbm, which = peaks()
        # roofline of the dominant kernel: work_kernel (side()/margin scan). One extra, untimed build
        # with CUDA events around every work_kernel launch on its launching stream.
        os.environ["ARROY_B200_PROFILE"] = "1"
        ctx.build_trees_begin([seeds[t] for t in my_trees])   # rank-local: no collective in here
        os.environ.pop("ARROY_B200_PROFILE")
        st = ctx.build_stats()
        scan_ms, steps_dev = st["scan_ms"], st["steps"]
        alg_bytes = st["scanned_rows"] * d * 4
        achieved = alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
        # DRAM traffic per launch: ratio dram/algorithmic of the committed ncu --set full capture of this
        # kernel, applied to this run's average algorithmic bytes per launch
        traffic, traffic_note = None, "no ncu capture committed"
        tp = os.path.join(ROOT, "profiles", "r01_work_kernel_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            ratio = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["algorithmic_bytes"]
            traffic = ratio * alg_bytes / max(steps_dev, 1)
            traffic_note = "avg algorithmic bytes/launch x %.4f (dram/algorithmic of %s)" % (ratio, tj["source"])
        r = np.random.default_rng(0)
        normal = (r.standard_normal(d) / np.sqrt(d)).astype(np.float32)
        root_ms, _ = ctx.time_scan(normal, (0.0, 0.0), n, iters=5, flush_l2=True)
        line["roofline"] = {
            "bound": "hbm", "kernel": "work_kernel (side()/margin scan + id partition)", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
            "peak_source": which, "traffic": traffic, "traffic_note": traffic_note,
            "per_launch": {"launches": steps_dev, "avg_ms": scan_ms / max(steps_dev, 1), "avg_algorithmic_GB": alg_bytes / max(steps_dev, 1) / 1e9},
            "root_scan": {"rows": n, "ms": root_ms, "GBps": n * d * 4 / (root_ms * 1e-3) / 1e9, "frac": n * d * 4 / (root_ms * 1e-3) / 1e9 / hbm},
            "share_of_step": scan_ms / (st["build_ms"] if st["build_ms"] else 1.0),
        }
    # ---- e2e: through the C ABI with HOST buffers (what a fork of the Rust crate would call) -----------
    # items = raw stored Leaf values [0x00][header][d x f32] at unaligned host addresses, exactly
    # what LMDB hands to ImmutableLeafs::new; nodes come back through the thread-safe arena sink
    # (the TmpNodes stand-in). Timed: decode + H2D + device build + D2H + NodeCodec encoding.
    if rank == 0 and not args.no_e2e and world == 1:
        host = items.cpu().numpy()
        ctx.stage_items_device(metric, ids, d, items.data_ptr())
        if metric == "dot-product":
            ctx.dot_preprocess()
        h0, h1 = ctx.item_headers()          # D::new_header / preprocess result, as stored by the writer
        hf = 2 if metric == "dot-product" else 1
        stride = 1 + 4 * hf + 4 * d           # odd => every value is byte aligned only
        blob = np.zeros(n * stride, dtype=np.uint8)
        b2 = blob.reshape(n, stride)
        b2[:, 1:5] = h0.view(np.uint8).reshape(n, 4)
        if hf == 2:
            b2[:, 5:9] = h1.view(np.uint8).reshape(n, 4)
        b2[:, 1 + 4 * hf:] = host.view(np.uint8).reshape(n, 4 * d)
        ptrs = (blob.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(stride)).astype(np.uint64)
        arena = ab.Arena()
        e_times, h2d, d2h, bd = [], 0, 0, None
        n_warm = max(1, min(args.warmup, 2))
        for step in range(n_warm + args.steps):
            arena.clear()
            cc0 = ctx.counters()
            t0 = time.perf_counter()
            ctx.stage_items_ptrs(metric, d, ids, ptrs)
            ctx.build_trees_into_arena(arena, seeds, list(range(T)), T)
            dt = time.perf_counter() - t0
            cc1 = ctx.counters()
            if step >= n_warm:
                e_times.append(dt)
                h2d, d2h = cc1["h2d_bytes"] - cc0["h2d_bytes"], cc1["d2h_bytes"] - cc0["d2h_bytes"]
                bd = ctx.build_breakdown()
        e_sec = sum(e_times) / len(e_times)
        n_nodes, node_bytes = arena.stats()
        line["e2e"] = {"value": n / e_sec, "unit": "vectors/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e_sec * 1e3,
                       "api": "arroy_b200_stage_items(leaf value pointers) + arroy_b200_build_trees(arena sink)", "nodes": int(n_nodes), "node_bytes": int(node_bytes),
                       "build_breakdown_ms": bd}
        del arena, blob, b2, ptrs
        # the same through the host mirror's Writer (adds the in-memory key/value table standing in for LMDB)
        if not args.no_writer_e2e:
            env = ab.Env(local_rank)
            env._ctx = ctx
            w = ab.Writer(env, 0, d, metric)
            w_times = []
            for step in range(1 + max(1, args.steps - 1)):
                w.clear()                # a FIRST build every step (re-adding items to a built index would take the incremental path)
                w.add_items(ids, host)   # Writer::add_item x n: the key/value puts are the caller's side of the API, not timed
                t0 = time.perf_counter()
                w.builder(ab.StdRng.from_seed(SEED)).n_trees(T).build()
                if step >= 1:
                    w_times.append(time.perf_counter() - t0)
            ws = sum(w_times) / len(w_times)
            line["e2e_writer"] = {"value": n / ws, "unit": "vectors/s", "ms_per_step": ws * 1e3, "api": "Writer.builder(rng).n_trees(T).build()", "breakdown_ms": w.build_timings()}
            if not args.no_query:
                # QPS @ recall: by_item queries for items 0..Q-1, top-100, default search_k
                Q, k = args.queries, 100
                reader = ab.Reader.open(env, 0, metric)
                qitems = np.arange(Q, dtype=np.uint32)
                reader.nns_batch_by_item(qitems, k)  # warm-up (stages the items, sizes the scratch buffers)
                q_times = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    out_ids, out_dist, out_len, qms = reader.nns_batch_by_item(qitems, k)
                    q_times.append(time.perf_counter() - t0)
                q_sec = min(q_times)
                t0 = time.perf_counter()
                single = [reader.nns(k).by_item(int(i)) for i in qitems[:50]]
                one_sec = (time.perf_counter() - t0) / 50
                assert all([s[0] for s in single[i]] == out_ids[i, :out_len[i]].tolist() for i in range(50))
                # exact ground truth: the same distance over ALL rows (brute force on the device)
                allrows = np.arange(n, dtype=np.uint32)
                hits = 0
                QG = min(Q, 100)
                h0q, _ = ctx.item_headers()
                offs = (np.arange(QG + 1, dtype=np.uint64) * np.uint64(n))
                g_rows, _, g_len = ctx.rerank_batch(host[:QG], h0q[:QG], np.tile(allrows, QG), offs, k)
                for i in range(QG):
                    hits += len(set(g_rows[i, :g_len[i]].tolist()) & set(out_ids[i, :out_len[i]].tolist()))
                line["query"] = {"qps_batched": Q / q_sec, "queries": Q, "k": k, "search_k": k * T, "recall_at_100": hits / (QG * k), "recall_queries": QG,
                                 "tree_walk_ms": qms["tree_walk_ms"], "rerank_ms": qms["rerank_ms"], "qps_one_at_a_time": 1.0 / one_sec,
                                 "api": "Reader.nns_batch_by_item: device tree walk (one warp per query) + fused bf16 pre-filter / exact re-score / top-k kernel; one-at-a-time = Reader.nns(100).by_item"}
                del reader
            env._ctx = None
            del w, env
        del host
    elif world > 1 and not args.no_e2e and d % 32 == 0 and metric != "dot-product":
        # ---- e2e at N > 1: rank 0 decodes + uploads the host leaf values, ONE NCCL broadcast straight out of the
        # library's item buffer, every rank builds and encodes its trees into its own host arena (sharded_build) ------
        class _DevView:   # zero-copy torch view of the staged item matrix of this context
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (int(ptr), False), "version": 2}
        if rank == 0:
            host = items.cpu().numpy()
            ctx.stage_items_device(metric, ids, d, items.data_ptr())
            h0, _ = ctx.item_headers()
            stride = 1 + 4 + 4 * d
            blob = np.zeros(n * stride, dtype=np.uint8)
            b2 = blob.reshape(n, stride)
            b2[:, 1:5] = h0.view(np.uint8).reshape(n, 4)
            b2[:, 5:] = host.view(np.uint8).reshape(n, 4 * d)
            ptrs = (blob.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(stride)).astype(np.uint64)
            del host
        arena = ab.Arena()
        roots = list(range(T))
        e_ms = []
        cc0 = None
        n_warm = 1
        for step in range(n_warm + args.steps):
            arena.clear()
            barrier()
            if step == n_warm:
                cc0 = ctx.counters()
            t0 = time.perf_counter()
            if rank == 0:
                ctx.stage_items_ptrs(metric, d, ids, ptrs)
                (p_items, _, _), ld_items = ctx.device_ptrs()
                src = torch.as_tensor(_DevView(p_items, (n, ld_items)), device=dev)
                parallel.broadcast_items(dist, src, src=0)
            else:
                parallel.broadcast_items(dist, items, src=0)
                torch.cuda.synchronize()
                ctx.stage_items_device(metric, ids, d, items.data_ptr())
            parallel.sharded_build(ctx, dist, rank, world, seeds, roots, T, arena=arena, device=dev)
            barrier()
            if step >= n_warm:
                e_ms.append((time.perf_counter() - t0) * 1e3)
        cc1 = ctx.counters()
        t_e = torch.tensor([sum(e_ms) / len(e_ms), float(cc1["h2d_bytes"] - cc0["h2d_bytes"]) / args.steps, float(cc1["d2h_bytes"] - cc0["d2h_bytes"]) / args.steps],
                           dtype=torch.float64, device=dev)
        t_max = t_e.clone()
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_e, op=dist.ReduceOp.SUM)
        if rank == 0:
            e_sec = float(t_max[0].item()) * 1e-3
            line["e2e"] = {"value": n / e_sec, "unit": "vectors/s", "h2d_bytes_per_step": int(t_e[1].item()), "d2h_bytes_per_step": int(t_e[2].item()), "ms_per_step": e_sec * 1e3,
                           "api": "rank 0: arroy_b200_stage_items(leaf value pointers); NCCL broadcast of the item buffer; every rank: arroy_b200_build_trees_begin / _emit (arena sink) for its trees",
                           "timing": "wall clock between barriers, max over ranks"}
    elif rank == 0:
        line["e2e"] = None
    # ---- CPU baseline (oracle port) on a bounded sample, rank 0, N=1 only --------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        cores = os.cpu_count() or 1
        t_sample = args.ref_trees or (T if n * d < 5e8 else min(T, max(8, cores // 2)))
        data = oracle.synth_rows(SEED, d, 0, n, wl["centre"], threads=min(cores, 64))
        db = oracle.Db(metric, d)
        db.set_items(ids, data)
        t0 = time.perf_counter()
        cpu_threads = 1 if args.workload == "c1" else min(cores, t_sample)   # configs[0] names the single-thread CPU reference
        db.build(oracle.StdRng(SEED), n_trees=t_sample, threads=cpu_threads)
        sec = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / (sec * T / t_sample), "unit": "vectors/s", "cores": cpu_threads, "kind": "port",
                                "sample": "%d of %d trees over the full %dx%d matrix in %.3f s on %d thread(s), extrapolated x%.2f" % (t_sample, T, n, d, sec, cpu_threads, T / t_sample)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
