#!/usr/bin/env python
"""bench.py — bitmap-scan rows/s and HBM GB/s of the CUBIT scan hot path on B200.

Workload (BASELINE.json configs[1]), per GPU:
    synthetic table of --rows rows (default 10^9), one cardinality-100 CUBIT index per
    selectivity point s ∈ {1e-4, 1e-3, 1e-2, 0.1, 0.25, 0.5} (values 10..19 share mass s),
    int64 payload column = global row id.
One STEP = the whole selectivity sweep: six queries
    OR over the 10 value bitvectors 10..19 → sorted int64 row IDs → probe payload at those
    rows (values materialised) → COUNT, SUM(payload)
each as two sm_100a kernels: the single-pass fused merge+decode kernel, then the bit-driven probe
kernel (gather + SUM).
    value  = table rows covered per second, inputs resident in HBM, device-timed (CUDA events)
    e2e    = the same sweep through the synchronous C-ABI call a DuckDB table function makes
             (host predicate structs in, aggregate row + first 2048-row DataChunk out)
Multi-GPU (torchrun, one rank per GPU): weak scaling, every rank owns --rows rows of a
row-range-sharded table (row_base = rank * rows); the only collective is one NCCL
all-reduce of the six (COUNT, SUM) aggregates per step.

--impl reference times the CPU restatement of the same path (oracle/, multi-threaded C) on
this box's host cores: the mounted reference has no CUBIT source to run.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from fractions import Fraction

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SELECTIVITIES = ["1e-4", "1e-3", "1e-2", "0.1", "0.25", "0.5"]
SEED = 0xC0B17
CARD, HOT_LO, HOT_N = 100, 10, 10
METRIC = "bitmap_scan_rows_per_s"
COL_PAYLOAD, COL_VALUE = 0, 1


def threshold(sel):
    return int(Fraction(sel) * (1 << 64))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[0]))
                    for nm, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU arm
def cpu_sweep(rows, steps, warmup, threads):
    """the oracle's multi-threaded scan on the same sweep, `rows` rows per selectivity point"""
    import numpy as np
    import oracle
    data = []
    payload = oracle.synth_column(0, rows)
    for s in SELECTIVITIES:
        bv = oracle.synth_bitvectors(rows, 0, SEED, threshold(s), CARD, HOT_LO, HOT_N, HOT_LO, HOT_N, threads)
        data.append([bv[i] for i in range(HOT_N)])
    n_words = (rows + 63) // 64
    bufs = (np.empty(n_words, dtype=np.uint64), np.empty(rows, dtype=np.int64), np.empty(rows, dtype=np.int64))
    times, checks = [], None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        res = []
        for grp in data:
            cnt, _ids, _vals, tot = oracle.scan_mt([grp], payload=payload, n_threads=threads, bufs=bufs)
            res.append((cnt, tot))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        checks = res
    times.sort()
    med = times[len(times) // 2]
    return len(SELECTIVITIES) * rows / med, med, checks


def duckdb_cpu_sweep(rows, checks):
    """the REFERENCE's own CPU path (unmodified DuckDB built from /root/reference: seq_scan + pushed-down filter) on
    the same sample, when the bundle of tools/build_ref_bundle.py travelled with the repo (baseline/_ref/); its
    answers must equal the oracle's on that sample.  None when the bundle is absent."""
    import subprocess
    exe = os.path.join(ROOT, "baseline", "_ref", "duckdb_cfg2_baseline")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe, str(rows), "3"] + [str(threshold(s)) for s in SELECTIVITIES], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, timeout=240)
        if r.returncode != 0:
            return {"error": r.stderr.strip()[-200:]}
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:  # the baseline is a report, never a reason to lose the bench line
        return {"error": str(e)[:200]}
    if checks is not None:
        assert [tuple(a) for a in d["answers"]] == [(int(c), int(t)) for c, t in checks], "DuckDB and the oracle disagree"
    return {"value": d["rows_per_s"], "unit": "rows/s", "cores": d["threads"], "kind": "reference",
            "sample": "reference DuckDB %s, SELECT count(*), sum(payload) FROM t WHERE v BETWEEN 10 AND 19 per sweep point on "
                      "%d rows per point, all host threads, median of 3; answers equal the oracle's on the same sample"
                      % (d["version"], rows)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    rows = args.cpu_rows
    value, med, checks = cpu_sweep(rows, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args.rows, args.gpus),
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port",
                         "sample": "same 6-point sweep on %d rows per point (bounded sample of the %d-row workload); "
                                   "CPU restatement oracle/cubit_oracle.c, %d pthreads; the reference tree has no "
                                   "CUBIT source to run" % (rows, args.rows, threads)},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    duck = duckdb_cpu_sweep(min(rows, 1 << 26), checks if rows <= (1 << 26) else None)
    if duck is not None:
        line["reference_duckdb_cpu"] = duck  # the vanilla reference's scan of the same sample, beside the CUBIT-style port
    print(json.dumps(line))
    return 0


def workload_config(rows, gpus):
    return {"workload": "cfg2: synthetic %d-row table per GPU, cardinality-100 CUBIT index, range predicate OR over "
                        "10 bitvectors (values 10..19), selectivity sweep %s; step = the 6-query sweep, each query "
                        "merge+decode→sorted int64 row IDs→probe int64 payload→COUNT,SUM" % (rows, ",".join(SELECTIVITIES)),
            "rows_per_gpu": rows, "cardinality": CARD, "k_bitvectors": HOT_N, "selectivities": SELECTIVITIES,
            "seg_bits": 65536, "l2_policy": "inputs larger than L2 (1.25 GB of bitvectors per query vs 126 MB L2)",
            "sharding": "row-range, %d shard(s)" % gpus}


# ------------------------------------------------------------------ GPU arm
def run_b200(args):
    import numpy as np
    import torch
    cubit = importlib.import_module("duckdb-cubit_b200")
    sharding = importlib.import_module("duckdb-cubit_b200.sharding")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUBIT GPU path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    rows = args.rows
    seg_bits = 65536
    rows_pad = (rows + seg_bits - 1) // seg_bits * seg_bits
    row_base = rank * rows_pad  # weak scaling: every rank owns `rows` rows of a world*rows table

    t = cubit.CubitTable(rows, row_base=row_base, seg_bits=seg_bits, device=local)
    stream = torch.cuda.current_stream()
    t.set_stream(stream.cuda_stream)
    t_build0 = time.time()
    t.synth_column(COL_PAYLOAD, 0)
    indexes, expect = [], []
    for s in SELECTIVITIES:
        t.synth_column(COL_VALUE, 1, seed=SEED, threshold=threshold(s), card=CARD, hot_lo=HOT_LO, hot_n=HOT_N)
        ix = t.create_index(CARD)
        t.build_index(ix, COL_VALUE, 0)
        indexes.append(ix)
        expect.append(sum(t.bitvector_count(ix, v) for v in range(HOT_LO, HOT_LO + HOT_N)))
    t.drop_column(COL_VALUE)
    build_s = time.time() - t_build0

    flags = cubit.Q_ROWIDS | cubit.Q_VALUES
    mk = lambda ix, extra: cubit.QueryPlan([[(ix, v) for v in range(HOT_LO, HOT_LO + HOT_N)]], flags | extra,  # noqa
                                           cols=[COL_PAYLOAD], agg=cubit.AGG_SUM, agg_a=COL_PAYLOAD)
    plans_async = [mk(ix, cubit.Q_ASYNC | cubit.Q_TIMING) for ix in indexes]
    plans_sync = [mk(ix, 0) for ix in indexes]
    n_words = t.n_words

    def step_device():
        res = [t.execute(p) for p in plans_async]
        for r in res:
            r.wait()
        agg = [(r.count, r.sum) for r in res]
        if world > 1:  # the one collective of the path: exact global COUNT/SUM of the six queries
            limbs = []
            for c, sm in agg:
                limbs += sharding.to_limbs(c, sm)
            tt = torch.tensor(limbs, dtype=torch.int64, device=dev)
            dist.all_reduce(tt)
            fl = tt.tolist()
            agg = [sharding.from_limbs(fl[i * 5:(i + 1) * 5]) for i in range(len(res))]
        infos = [(r.info.ms_scan, r.info.ms_probe, r.info.algo_bytes_scan, r.info.algo_bytes_probe, r.count,
                  r.info.n_launches, r.info.fused) for r in res]
        for r in res:
            r.free()
        return agg, infos

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        agg, infos = step_device()
    # correctness of what is being timed: COUNT equals Σ popcount of the (disjoint) value
    # bitvectors and SUM(payload) equals the closed form only via the global check below
    for inf, e in zip(infos, expect):
        assert inf[4] == e, "count %d != expected %d" % (inf[4], e)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    launches0 = t.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record(stream)
    # per-kernel accumulators: [ms, algorithmic bytes, launches]
    kscan, kprobe = [0.0, 0, 0], [0.0, 0, 0]
    per_sel = [dict(ms_scan=0.0, ms_probe=0.0, by_scan=0, by_probe=0, cnt=0, fused=0) for _ in SELECTIVITIES]
    for _ in range(args.steps):
        agg, infos = step_device()
        for i, (ms_s, ms_p, by_s, by_p, cnt, _nl, fused) in enumerate(infos):
            ps = per_sel[i]
            ps["cnt"], ps["fused"] = cnt, fused
            ps["ms_scan"] += ms_s
            ps["ms_probe"] += ms_p
            ps["by_scan"] += by_s
            ps["by_probe"] += by_p
            if ms_p > 0:        # separate probe kernel ran
                kscan[0] += ms_s
                kscan[1] += by_s
                kscan[2] += 1
                kprobe[0] += ms_p
                kprobe[1] += by_p
                kprobe[2] += 1
            else:               # probe fused into the scan kernel
                kscan[0] += ms_s
                kscan[1] += by_s + by_p
                kscan[2] += 1
    e1.record(stream)
    barrier()
    wall1 = time.time()
    launches = t.launch_count - launches0
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        tl = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(tl)
        launches = int(tl.item())
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms_per_step = ms / args.steps
    n_q = len(SELECTIVITIES)
    value = world * n_q * rows / (ms_per_step * 1e-3)

    # ---- e2e: the synchronous C-ABI call path with host buffers (what the table function does)
    chunk = 2048
    ids_host = np.empty(chunk, dtype=np.int64)
    val_host = [np.empty(chunk, dtype=np.int64)]
    scan_args_bytes = 64 * 8 * 3 + 160  # kernel-parameter block carrying the flattened predicate

    def step_e2e():
        out = []
        for p in plans_sync:
            with t.execute(p) as r:            # blocks until COUNT/SUM are on the host
                n = min(chunk, r.count)
                r.fetch(0, n, out_ids=ids_host, out_cols=val_host)   # first DataChunk (GetData call #1)
                out.append((r.count, r.sum))
        if world > 1:
            limbs = []
            for c, sm in out:
                limbs += sharding.to_limbs(c, sm)
            tt = torch.tensor(limbs, dtype=torch.int64, device=dev)
            dist.all_reduce(tt)
            fl = tt.tolist()
            out = [sharding.from_limbs(fl[i * 5:(i + 1) * 5]) for i in range(n_q)]
        return out

    for _ in range(max(1, args.warmup)):
        e2e_out = step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_out = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = world * n_q * rows / (e2e_s / args.steps)
    # global self-check: payload = global row id, so SUM(payload) must equal SUM(row ids);
    # verified exactly on rank 0's shard for the smallest query by fetching its row IDs
    with t.execute(plans_sync[0]) as r:
        ids, (vals,) = r.fetch()
        assert (ids == vals).all() and int(ids.sum()) == r.sum and (np.diff(ids) > 0).all()
    for (c, _s), e in zip(agg, expect):
        if world == 1:
            assert c == e

    # ---- full materialisation to host (every row ID + value over PCIe), reported beside e2e
    e2e_full = None
    if world == 1 and not args.no_materialize:
        cap = max(expect)
        pin_ids = torch.empty(cap, dtype=torch.int64, pin_memory=True).numpy()
        pin_val = [torch.empty(cap, dtype=torch.int64, pin_memory=True).numpy()]

        def step_full():
            for p in plans_sync:
                with t.execute(p) as r:
                    r.fetch(0, r.count, out_ids=pin_ids, out_cols=pin_val)
        step_full()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        reps = max(1, min(2, args.steps))
        for _ in range(reps):
            step_full()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - w0) / reps
        e2e_full = {"value": n_q * rows / dt, "unit": "rows/s", "d2h_bytes_per_step": 16 * sum(expect),
                    "note": "every selected row ID + payload value copied to pinned host memory (PCIe bound)"}

    if rank != 0:
        t.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp))
        except Exception:
            traffic = {}

    def roof(acc, kernel, tkey, formula):
        ms, by, n = acc
        ach = by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "frac_of_nominal_8TBs": ach / 8000.0, "traffic": traffic.get(tkey), "peak_source": peak_src,
                "launches": n, "bytes_per_launch_avg": by / n if n else 0, "ms_per_launch_avg": ms / n if n else 0,
                "share_of_step": ms / (ms_per_step * args.steps) if ms_per_step > 0 else 0, "bytes_formula": formula}

    roof_scan = roof(kscan, "cubit_scan_kernel<4,false,0> (segment merge + bit->row-ID decode, single pass)",
                     "cubit_scan_kernel_dram_bytes_per_launch",
                     "k*ceil(N/64)*8 + 8*M  [SURVEY 8d]")
    roof_probe = roof(kprobe, "cubit_probe_bits_kernel<4,1,true> (bit-driven probe: gather payload + SUM; "
                      "cubit_probe_kernel over the row-ID list for the two sparsest points)",
                      "cubit_probe_bits_kernel_dram_bytes_per_launch",
                      "8*M payload values read + ceil(N/64)*8 re-read of the merged bitvector  [SURVEY 8d: P]")
    dominant = roof_probe if kprobe[0] > kscan[0] else roof_scan
    sweep = []
    for s, ps in zip(SELECTIVITIES, per_sel):
        ms_tot = (ps["ms_scan"] + ps["ms_probe"]) / args.steps
        g_scan = ps["by_scan"] / (ps["ms_scan"] * 1e-3) / 1e9 if ps["ms_scan"] > 0 else 0.0
        sweep.append({"selectivity": s, "rows_selected": ps["cnt"], "probe_fused": bool(ps["fused"] and ps["ms_probe"] == 0),
                      "scan_ms": ps["ms_scan"] / args.steps, "probe_ms": ps["ms_probe"] / args.steps,
                      "rows_per_s": rows / (ms_tot * 1e-3) if ms_tot > 0 else 0.0,
                      "scan_algo_gbs": g_scan, "scan_frac_of_peak": g_scan / peak})
    kernel_bytes = kscan[1] + kprobe[1]
    line = {
        "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_config(rows, world),
        "hbm_gbs": world * (kernel_bytes / args.steps) / (ms_per_step * 1e-3) / 1e9,
        "roofline": dominant, "roofline_merge_decode": roof_scan, "roofline_probe": roof_probe,
        "sweep": sweep,
        "e2e": {"value": e2e_value, "unit": "rows/s",
                "h2d_bytes_per_step": n_q * scan_args_bytes,
                "d2h_bytes_per_step": n_q * (32 + 2 * 8 * chunk),
                "note": "synchronous cubit_gpu_query + cubit_gpu_fetch of the first 2048-row DataChunk per query; "
                        "host predicate structs in, COUNT/SUM row + chunk out"},
        "e2e_full_materialize": e2e_full,
        "gpu_launches": launches, "clocks": clocks, "index_build_s": build_s,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cv, cmed, checks = cpu_sweep(args.cpu_rows, 3, 1, threads)
        line["cpu_baseline"] = {"value": cv, "unit": "rows/s", "cores": threads, "kind": "port",
                                "sample": "same 6-point sweep on %d rows per point, %d pthreads, median of 3 "
                                          "(oracle/cubit_oracle.c; no reference CUBIT CPU source exists)"
                                          % (args.cpu_rows, threads)}
        duck = duckdb_cpu_sweep(min(args.cpu_rows, 1 << 26), checks if args.cpu_rows <= (1 << 26) else None)
        if duck is not None:
            line["reference_duckdb_cpu"] = duck
    print(json.dumps(line))
    t.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="table rows per GPU")
    ap.add_argument("--cpu-rows", type=int, default=1 << 26, help="rows per sweep point of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-materialize", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
