#!/usr/bin/env python
"""bench.py — bitmap-scan rows/s and HBM GB/s of the CUBIT scan hot path on B200.

Workload (BASELINE.json configs[1]), per GPU:
    synthetic table of --rows rows (default 10^9), one cardinality-100 CUBIT index per
    selectivity point s ∈ {1e-4, 1e-3, 1e-2, 0.1, 0.25, 0.5} (values 10..19 share mass s),
    int64 payload column = global row id.
One STEP = the whole selectivity sweep: six queries
    OR over the 10 value bitvectors 10..19 → sorted int64 row IDs → probe payload at those
    rows (values materialised) → COUNT, SUM(payload)
each as two sm_100a kernels: the single-pass fused merge+decode kernel, then a probe kernel (gather + SUM).  The
payload column is resident raw (int64) AND FOR-bit-packed (the form DuckDB itself stores numeric columns in,
SURVEY §8f rank 3); the library picks per query: gather over the row-ID list (sparsest points), bit-driven gather,
or — from ~2 % of the rows upward — the dense probe that streams the packed column through shared memory
(--payload-form raw|packed|both).  A second, 24-bit uniformly random payload (TPC-H l_extendedprice-like) is
probed in both forms after the timed region and reported under "payload_24bit".
    value  = table rows covered per second, inputs resident in HBM, device-timed (CUDA events)
    e2e    = the same sweep through the synchronous C-ABI call a DuckDB table function makes
             (host predicate structs in, aggregate row + first 2048-row DataChunk out)
Multi-GPU (torchrun, one rank per GPU): weak scaling by default — every rank owns --rows rows of a
row-range-sharded table (row_base = rank * rows); --scaling strong cuts ONE --rows-row table over the ranks
(config 5: --rows 16000000000 --workload cfg5).  The only collective is one NCCL all-reduce per step of the
(COUNT, SUM) limbs, which the library accumulates ON THE DEVICE behind each query (cubit_gpu_result_add_limbs): no
aggregate visits the host inside the timed region.

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


def workload_config(rows, gpus, scaling="weak", workload="cfg2"):
    if workload == "cfg5":
        return {"workload": "cfg5: ONE synthetic %d-row table row-range sharded over %d GPU(s), 8 predicates (OR over 8 value "
                            "bitvectors; AND of two OR-of-4 groups), 10 %% of the rows on values 10..19; step = the two "
                            "queries, each merge+decode→sorted int64 row IDs (+ payload probe and SUM when the payload "
                            "fits)" % (rows, gpus),
                "rows_total": rows, "k_bitvectors": 8, "seg_bits": 65536, "scaling": scaling,
                "l2_policy": "inputs larger than L2", "sharding": "row-range, %d shard(s)" % gpus}
    return {"workload": "cfg2: synthetic %d-row table %s, cardinality-100 CUBIT index, range predicate OR over "
                        "10 bitvectors (values 10..19), selectivity sweep %s; step = the 6-query sweep, each query "
                        "merge+decode→sorted int64 row IDs→probe int64 payload→COUNT,SUM"
                        % (rows, "per GPU" if scaling == "weak" else "sharded over the GPUs", ",".join(SELECTIVITIES)),
            "rows_per_gpu" if scaling == "weak" else "rows_total": rows, "cardinality": CARD, "k_bitvectors": HOT_N,
            "selectivities": SELECTIVITIES,
            "seg_bits": 65536, "l2_policy": "inputs larger than L2 (1.25 GB of bitvectors per query vs 126 MB L2)",
            "sharding": "row-range, %d shard(s)" % gpus}



# ------------------------------------------------------------------ DRAM traffic (ncu, same invocation)
KERNEL_FAMILIES = ("cubit_scan_kernel", "cubit_probe_dense_kernel", "cubit_probe_bits_kernel", "cubit_probe_kernel")


def measure_traffic(args):
    """DRAM bytes per launch of the step's kernels, measured by ncu on ONE step of the same workload in a child
    process (after the timed region; the parent has released its table).  None when ncu cannot run here."""
    import shutil
    import tempfile
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    log = tempfile.NamedTemporaryFile(prefix="cubit_ncu_", suffix=".csv", delete=False).name
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
           "--print-units", "base", "-k", "regex:cubit_(scan|probe)", "--csv", "--log-file", log,
           sys.executable, os.path.abspath(__file__), "--traffic-child", "--rows", str(args.rows),
           "--payload-form", args.payload_form]
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        if r.returncode != 0:
            return None, "ncu child failed: " + (r.stderr.strip() or r.stdout.strip())[-160:]
        return parse_ncu_traffic(log), None
    except Exception as e:
        return None, "ncu child: " + str(e)[:160]
    finally:
        try:
            os.unlink(log)
        except OSError:
            pass


def parse_ncu_traffic(path):
    import csv
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = next(r for r in rows if "Kernel Name" in r and "Metric Value" in r)
    ik, im, iv, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per_launch = {}
    for r in rows:
        if r is hdr or len(r) <= iv or not r[ii].strip().isdigit():
            continue
        fam = next((f for f in KERNEL_FAMILIES if f in r[ik]), None)
        if fam is None or not r[im].startswith("dram__bytes"):
            continue
        per_launch.setdefault((fam, int(r[ii])), 0.0)
        per_launch[(fam, int(r[ii]))] += float(r[iv].replace(",", ""))
    out = {}
    for (fam, _id), b in per_launch.items():
        d = out.setdefault(fam, {"launches": 0, "dram_bytes": 0.0})
        d["launches"] += 1
        d["dram_bytes"] += b
    for d in out.values():
        d["dram_bytes_per_launch"] = d["dram_bytes"] / d["launches"]
    return out


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
    seg_bits = 65536
    strong = args.scaling == "strong"
    cfg5 = args.workload == "cfg5"
    if strong:   # ONE table of --rows rows cut into contiguous row ranges of whole segments
        lo, hi = sharding.shard_ranges(args.rows, world, seg_bits)[rank]
        rows, row_base = hi - lo, lo
        total_rows = args.rows
    else:        # weak scaling: every rank owns --rows rows of a world * rows table
        rows = args.rows
        row_base = rank * ((rows + seg_bits - 1) // seg_bits * seg_bits)
        total_rows = world * rows

    t = cubit.CubitTable(rows, row_base=row_base, seg_bits=seg_bits, device=local)
    # ONE stream for the library's kernels, torch's limb arithmetic and NCCL's ordering: a real (non-default) torch
    # stream made current — the legacy default stream has handle 0, which the library reads as "use your own stream"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    t.set_stream(stream.cuda_stream)
    t_build0 = time.time()
    with_payload = rows <= 8_000_000_000  # 8 B/row payload + index + row IDs must fit 180 GB
    packed_bytes = None
    if with_payload:
        t.synth_column(COL_PAYLOAD, 0)
        if args.payload_form != "raw":   # FOR-bit-packed form next to (or instead of) the raw int64 array
            packed_bytes = t.pack_column(COL_PAYLOAD, keep_raw=(args.payload_form == "both"))
    indexes, expect, plans_async, plans_sync, labels = [], [], [], [], []
    flags = cubit.Q_ROWIDS | (cubit.Q_VALUES if with_payload else 0)
    kw = dict(cols=[COL_PAYLOAD], agg=cubit.AGG_SUM, agg_a=COL_PAYLOAD) if with_payload else {}
    if cfg5:
        t.synth_column(COL_VALUE, 1, seed=SEED, threshold=threshold("0.1"), card=CARD, hot_lo=HOT_LO, hot_n=HOT_N)
        ix = t.create_index(18)  # only the bitvectors the two queries read are kept (values 0..17)
        t.build_index(ix, COL_VALUE, 0)
        pop = [t.bitvector_count(ix, v) for v in range(18)]
        qs = [("or_of_8", [[(ix, v) for v in range(10, 18)]], sum(pop[10:18])),
              ("and_of_two_or_of_4", [[(ix, v) for v in (10, 11, 12, 13)], [(ix, v) for v in (12, 13, 14, 15)]], pop[12] + pop[13])]
        for name, groups, want in qs:
            labels.append(name)
            expect.append(want)
            plans_async.append(cubit.QueryPlan(groups, flags | cubit.Q_ASYNC | cubit.Q_TIMING, **kw))
            plans_sync.append(cubit.QueryPlan(groups, flags, **kw))
    else:
        for s in SELECTIVITIES:
            t.synth_column(COL_VALUE, 1, seed=SEED, threshold=threshold(s), card=CARD, hot_lo=HOT_LO, hot_n=HOT_N)
            ix = t.create_index(CARD)
            t.build_index(ix, COL_VALUE, 0)
            indexes.append(ix)
            expect.append(sum(t.bitvector_count(ix, v) for v in range(HOT_LO, HOT_LO + HOT_N)))
            groups = [[(ix, v) for v in range(HOT_LO, HOT_LO + HOT_N)]]
            labels.append(s)
            plans_async.append(cubit.QueryPlan(groups, flags | cubit.Q_ASYNC | cubit.Q_TIMING, **kw))
            plans_sync.append(cubit.QueryPlan(groups, flags, **kw))
    t.drop_column(COL_VALUE)
    build_s = time.time() - t_build0
    n_q = len(plans_async)

    # (COUNT, SUM) of every query of a step as int64 limbs ON THE DEVICE: the library adds them behind each query
    # (cubit_gpu_result_add_limbs), one NCCL all-reduce per step sums them over the ranks, a running total keeps the
    # steps — nothing is read back before the timed region ends
    step_limbs = torch.zeros(5 * n_q, dtype=torch.int64, device=dev)
    total_limbs = torch.zeros(5 * n_q, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def step_device():
        step_limbs.zero_()
        res = [t.execute(p) for p in plans_async]
        for i, r in enumerate(res):
            r.add_limbs(step_limbs.data_ptr() + 40 * i)
        if world > 1:  # the one collective of the path: exact global COUNT/SUM of the step's queries
            dist.all_reduce(step_limbs)
        total_limbs.add_(step_limbs)
        for r in res:
            r.wait()
        infos = [(r.info.ms_scan, r.info.ms_probe, r.info.algo_bytes_scan, r.info.algo_bytes_probe, r.count,
                  r.info.n_launches, r.info.fused, r.info.probe_path) for r in res]
        for r in res:
            r.free()
        return infos

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if args.traffic_child:   # under ncu (measure_traffic): exactly one step of the workload, nothing printed
        step_device()
        torch.cuda.synchronize()
        t.close()
        return 0
    for _ in range(args.warmup):
        infos = step_device()
    # correctness of what is being timed: COUNT equals Σ popcount of the (disjoint) value bitvectors; SUM(payload)
    # is checked against the row IDs below and globally through the limbs
    for inf, e in zip(infos, expect):
        assert inf[4] == e, "count %d != expected %d" % (inf[4], e)
    barrier()
    total_limbs.zero_()

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
    kpath = {}   # the same per probe kernel (cubit_result_info.probe_path)
    per_q = [dict(ms_scan=0.0, ms_probe=0.0, by_scan=0, by_probe=0, cnt=0, fused=0, path=0) for _ in range(n_q)]
    for _ in range(args.steps):
        infos = step_device()
        for i, (ms_s, ms_p, by_s, by_p, cnt, _nl, fused, path) in enumerate(infos):
            ps = per_q[i]
            ps["cnt"], ps["fused"], ps["path"] = cnt, fused, path
            if ms_p > 0:
                kp = kpath.setdefault(path, [0.0, 0, 0])
                kp[0] += ms_p
                kp[1] += by_p
                kp[2] += 1
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
            else:               # probe fused into the scan kernel (or no probe)
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
    value = total_rows * n_q / (ms_per_step * 1e-3)
    # the limbs, now that the timed region is over: steps x the global (COUNT, SUM) of every query
    tl_host = total_limbs.cpu().tolist()
    global_agg = []
    for i in range(n_q):
        c, l0, l1, l2, l3 = tl_host[5 * i:5 * i + 5]
        assert c % args.steps == 0
        global_agg.append((c // args.steps, (l0 + (l1 << 32) + (l2 << 64) + (l3 << 96)) // args.steps))
    if world == 1:
        for (c, _s), e in zip(global_agg, expect):
            assert c == e, "COUNT from the device limbs %s != expected %s (limbs %s)" % (
                [g[0] for g in global_agg], expect, tl_host)
    else:
        ge = torch.tensor(expect, dtype=torch.int64, device=dev)
        dist.all_reduce(ge)
        assert [c for c, _ in global_agg] == ge.tolist(), "global COUNT differs from the sum of the shards' popcounts"

    # ---- e2e: the synchronous C-ABI call path with host buffers (what the table function does when the aggregate is
    # pushed down): host predicate structs in, COUNT/SUM row + the first DataChunk out.  NOT a row-returning number.
    chunk = 2048
    ids_host = np.empty(chunk, dtype=np.int64)
    val_host = [np.empty(chunk, dtype=np.int64)] if with_payload else []
    scan_args_bytes = 64 * 8 * 3 + 160  # kernel-parameter block carrying the flattened predicate

    def step_e2e():
        step_limbs.zero_()
        for i, p in enumerate(plans_sync):
            with t.execute(p) as r:            # blocks until COUNT/SUM are on the host
                n = min(chunk, r.count)
                r.fetch(0, n, out_ids=ids_host, out_cols=val_host)   # first DataChunk (GetData call #1)
                if world > 1:
                    r.add_limbs(step_limbs.data_ptr() + 40 * i)
        if world > 1:
            dist.all_reduce(step_limbs)
            return step_limbs.cpu()            # the global aggregate row reaches the host: part of the call
        return None

    for _ in range(max(1, args.warmup)):
        step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = total_rows * n_q / (e2e_s / args.steps)
    # self-check on this rank's shard for the smallest query: payload = global row id, so the probed values must
    # equal the row IDs, SUM(payload) their sum, and the IDs must ascend inside the shard's row range
    with t.execute(plans_sync[0]) as r:
        if with_payload:
            ids, (vals,) = r.fetch()
            assert (ids == vals).all() and int(ids.sum()) == r.sum
        else:
            ids, _ = r.fetch()
        assert (np.diff(ids) > 0).all() and (len(ids) == 0 or (ids[0] >= row_base and ids[-1] < row_base + rows))

    # ---- row-returning e2e: EVERY selected row ID + payload value to page-locked host memory, on every rank, through
    # the asynchronous hand-off (two windows in flight per result: window i+1 crosses PCIe while window i is consumed)
    e2e_full = None
    if not args.no_materialize:
        win = 1 << 22
        ncols = 1 if with_payload else 0
        bufs = [(torch.empty(win, dtype=torch.int64, pin_memory=True).numpy(),
                 [torch.empty(win, dtype=torch.int64, pin_memory=True).numpy() for _ in range(ncols)]) for _ in range(2)]

        def step_full():
            moved = 0
            for p in plans_sync:
                with t.execute(p) as r:
                    cnt = r.count
                    nwin = (cnt + win - 1) // win
                    tk = [None, None]
                    if nwin:
                        tk[0] = r.fetch_async(0, min(win, cnt), bufs[0][0], bufs[0][1])
                    for w in range(nwin):
                        if w + 1 < nwin:
                            o = (w + 1) * win
                            tk[(w + 1) & 1] = r.fetch_async(o, min(win, cnt - o), bufs[(w + 1) & 1][0], bufs[(w + 1) & 1][1])
                        r.fetch_wait(tk[w & 1])   # window w is in host memory: the consumer would read it here
                    moved += cnt * 8 * (1 + ncols)
            return moved
        step_full()
        barrier()
        w0 = time.perf_counter()
        reps = max(1, min(2, args.steps))
        for _ in range(reps):
            moved = step_full()
        barrier()
        dt = (time.perf_counter() - w0) / reps
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            tm = torch.tensor([moved], dtype=torch.int64, device=dev)
            dist.all_reduce(tm)
            moved = int(tm.item())
        e2e_full = {"value": total_rows * n_q / dt, "unit": "rows/s", "d2h_bytes_per_step": moved,
                    "pcie_GBps_all_gpus": moved / dt / 1e9,
                    "note": "every selected row ID + payload value copied to page-locked host memory on every rank "
                            "(asynchronous double-buffered hand-off; PCIe bound)"}

    # ---- the same hand-off over the NARROW WIRE (include/cubit_gpu_wire.h): per DataChunk and stream the GPU ships a
    # base + the narrowest deltas straight into page-locked windows; host workers (cubit_gpu_drain) widen ONE DataChunk
    # at a time into cache-resident vectors and a consumer reads every value (wrapping sums, checked against the
    # device's own SUM) — every selected row is DELIVERED as int64 DataChunks, which is what GetData hands on.
    e2e_drain = None
    if not args.no_materialize:
        drain_threads = max(1, min(16, (os.cpu_count() or 1) // world))
        ncols = 1 if with_payload else 0

        def step_drain():
            wire = wide = 0
            for p in plans_sync:
                with t.execute(p) as r:
                    st = r.drain(rowids=True, n_cols=ncols, threads=drain_threads, window_rows=args.drain_window)
                    assert st.rows == r.count
                    if with_payload:  # payload = global row id: both streams must sum to the device's SUM
                        assert st.sum_rowids == st.sum_cols[0] == r.sum % (1 << 64), "drain checksum != device SUM"
                    wire += st.wire_bytes
                    wide += st.wide_bytes
            return wire, wide
        step_drain()
        barrier()
        w0 = time.perf_counter()
        reps = max(1, min(3, args.steps))
        for _ in range(reps):
            wire_b, wide_b = step_drain()
        barrier()
        dt = (time.perf_counter() - w0) / reps
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            tm = torch.tensor([wire_b, wide_b], dtype=torch.int64, device=dev)
            dist.all_reduce(tm)
            wire_b, wide_b = int(tm[0].item()), int(tm[1].item())
        e2e_drain = {"value": total_rows * n_q / dt, "unit": "rows/s", "d2h_bytes_per_step": wire_b,
                     "wide_bytes_per_step": wide_b, "pcie_GBps_all_gpus": wire_b / dt / 1e9,
                     "host_threads_per_gpu": drain_threads, "window_rows": args.drain_window or 128 * 2048,
                     "rows_delivered_per_s": wide_b / (8 * (1 + ncols)) / dt,
                     "note": "every selected row ID + payload value delivered to host consumers as 2048-row int64 "
                             "DataChunks: narrow wire (per-chunk base + 1/2/4/8-byte deltas chosen and written by "
                             "the GPU, zero-copy), widened chunk by chunk in cache by cubit_gpu_drain's workers, every "
                             "value read (checksums equal the device SUM)"}

    # ---- the result gather (north_star: "per-shard row-ID lists are concatenated with shard offsets, NCCL only for
    # the final aggregate or result gather"): every rank's sorted row IDs of the s = 1e-2 query to rank 0 over NCCL
    # send/recv, straight from the library's device buffers; concatenation in rank order is the globally sorted list
    rowid_gather = None
    if world > 1 and not cfg5:
        gi = SELECTIVITIES.index("1e-2")
        gplan = cubit.QueryPlan([[(indexes[gi], v) for v in range(HOT_LO, HOT_LO + HOT_N)]], cubit.Q_ROWIDS)
        best = None
        for _ in range(3):
            with t.execute(gplan) as r:
                loc = sharding.result_rowids_tensor(r, dev)
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(stream)
                full = sharding.gather_sorted(loc, dist, dst=0)
                g1.record(stream)
                torch.cuda.synchronize()
                ms_g = g0.elapsed_time(g1)
                if rank == 0:
                    assert bool((full[1:] > full[:-1]).all()), "gathered row IDs are not ascending"
                    n_full = int(full.numel())
                del full
            tg = torch.tensor([ms_g], dtype=torch.float64, device=dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            best = float(tg.item()) if best is None else min(best, float(tg.item()))
        if rank == 0:
            rowid_gather = {"query": "s = 1e-2", "rows_gathered": n_full, "ms": best,
                            "GBps_into_rank0": n_full * 8 * (world - 1) / world / (best * 1e-3) / 1e9,
                            "note": "NCCL send/recv of the per-shard row-ID lists to rank 0 (zero-copy views of the "
                                    "library's device buffers); checked strictly ascending"}

    if rank != 0:
        t.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- a realistically distributed payload in both forms (not part of the step): uniform values of 24 significant
    # bits over a base, like TPC-H l_extendedprice in cents; column 2 stays raw, column 3 holds the same values packed
    payload24 = None
    if world == 1 and with_payload and not cfg5 and not args.no_payload24:
        COL_P24_RAW, COL_P24_PK = 2, 3
        t.synth_column(COL_P24_RAW, 3, seed=0xFEED, threshold=1 << 24, hot_lo=90000)
        t.synth_column(COL_P24_PK, 3, seed=0xFEED, threshold=1 << 24, hot_lo=90000)
        pk_bytes = t.pack_column(COL_P24_PK, keep_raw=False)
        payload24 = {"distribution": "uniform int64 in [90000, 90000 + 2^24)", "raw_bytes": rows * 8,
                     "packed_bytes": pk_bytes, "points": []}
        names = {0: "none", 1: "fused", 2: "bit-driven gather", 3: "gather over row IDs", 4: "dense (streamed)"}
        for s_lbl, ix in zip(labels, indexes):
            groups = [[(ix, v) for v in range(HOT_LO, HOT_LO + HOT_N)]]
            rec = {"selectivity": s_lbl}
            sums = []
            for form, col in (("raw", COL_P24_RAW), ("packed", COL_P24_PK)):
                plan = cubit.QueryPlan(groups, flags | cubit.Q_TIMING, cols=[col], agg=cubit.AGG_SUM, agg_a=col)
                best = None
                for _ in range(4):
                    with t.execute(plan) as r:
                        ms_p, path, sm = r.info.ms_probe, r.info.probe_path, r.sum
                    best = ms_p if best is None or ms_p < best else best
                rec[form + "_probe_ms"] = best
                rec[form + "_kernel"] = names.get(path, str(path))
                sums.append(sm)
            assert sums[0] == sums[1], "raw and packed forms of the 24-bit payload disagree"
            payload24["points"].append(rec)
        t.drop_column(COL_P24_RAW)
        t.drop_column(COL_P24_PK)

    peak, peak_src = measured_peak()
    # ---- DRAM traffic of the step's kernels: ncu on one step of this very workload, in a child process, now that
    # the timed regions are over (the table is released first: two copies do not fit).  Falls back to the committed
    # capture of the same command when ncu cannot run here.
    traffic, traffic_src = None, None
    if world == 1 and not args.no_traffic and not cfg5:
        t.close()
        traffic, err = measure_traffic(args)
        traffic_src = "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on one step of this workload, child " \
                      "process of this run" if traffic else None
        if traffic is None:
            traffic_src = "unavailable in this run (%s)" % err
    if traffic is None:
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                static = json.load(open(tp))
                traffic = {k[:-len("_dram_bytes_per_launch")]: {"dram_bytes_per_launch": v} for k, v in static.items()
                           if k.endswith("_dram_bytes_per_launch")}
                traffic_src = (traffic_src + "; " if traffic_src else "") + "static: profiles/" + str(static.get("source"))
            except Exception:
                traffic = None

    def roof(acc, kernel, family, formula):
        ms, by, n = acc
        ach = by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        fams = family if isinstance(family, (list, tuple)) else [family]
        tr = None
        if traffic:   # bytes per launch, averaged over the launches of these kernel families in one step
            tb = sum(traffic[f].get("dram_bytes", traffic[f]["dram_bytes_per_launch"]) for f in fams if f in traffic)
            tn = sum(traffic[f].get("launches", 1) for f in fams if f in traffic)
            tr = tb / tn if tn else None
        return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "frac_of_nominal_8TBs": ach / 8000.0, "traffic": tr, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "launches": n, "bytes_per_launch_avg": by / n if n else 0, "ms_per_launch_avg": ms / n if n else 0,
                "share_of_step": ms / (ms_per_step * args.steps) if ms_per_step > 0 else 0, "bytes_formula": formula}

    roof_scan = roof(kscan, "cubit_scan_kernel<4,false,0,true,false> (segment merge + bit->row-ID decode, single pass)",
                     "cubit_scan_kernel", "k*ceil(N/64)*8 + 8*M  [SURVEY 8d]")
    path_kernel = {2: ("cubit_probe_bits_kernel", "bit-driven gather probe over the merged bitvector"),
                   3: ("cubit_probe_kernel", "gather probe over the row-ID list (sparsest points)"),
                   4: ("cubit_probe_dense_kernel", "dense probe: pack blocks of the bit-packed payload streamed "
                                                   "through shared memory by bulk async copies")}
    roof_paths = {}
    for path, acc in sorted(kpath.items()):
        fam, what = path_kernel.get(path, ("?", "?"))
        roof_paths[fam] = roof(acc, "%s (%s)" % (fam, what), fam,
                               "8*M payload values (+ 8*M row IDs re-read by the gather kernel)  [SURVEY 8d: P]")
    roof_probe = roof(kprobe, "probe kernels of the step: " + ", ".join(
                          "%s x%d" % (path_kernel.get(p, ("?",))[0], a[2] // max(1, args.steps)) for p, a in sorted(kpath.items())),
                      [path_kernel[p][0] for p in kpath if p in path_kernel],
                      "8*M payload values (+ 8*M row IDs re-read by the gather kernel); the merged bitvector the "
                      "probe re-reads is intermediate traffic and is not counted  [SURVEY 8d: P]")
    dominant = roof_probe if kprobe[0] > kscan[0] else roof_scan
    sweep = []
    for s, ps in zip(labels, per_q):
        ms_tot = (ps["ms_scan"] + ps["ms_probe"]) / args.steps
        g_scan = ps["by_scan"] / (ps["ms_scan"] * 1e-3) / 1e9 if ps["ms_scan"] > 0 else 0.0
        sweep.append({"query" if cfg5 else "selectivity": s, "rows_selected": ps["cnt"],
                      "probe_fused": bool(ps["fused"] and ps["ms_probe"] == 0),
                      "probe_kernel": path_kernel.get(ps["path"], ("none",))[0],
                      "scan_ms": ps["ms_scan"] / args.steps, "probe_ms": ps["ms_probe"] / args.steps,
                      "rows_per_s": rows / (ms_tot * 1e-3) if ms_tot > 0 else 0.0,
                      "scan_algo_gbs": g_scan, "scan_frac_of_peak": g_scan / peak})
    kernel_bytes = kscan[1] + kprobe[1]
    line = {
        "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args.rows, world, args.scaling, args.workload),
        "hbm_gbs": world * (kernel_bytes / args.steps) / (ms_per_step * 1e-3) / 1e9,
        "roofline": dominant, "roofline_merge_decode": roof_scan, "roofline_probe": roof_probe,
        "roofline_probe_kernels": roof_paths,
        "payload": {"column": "int64 = global row id (SURVEY 8d config 2; self-checking: SUM(payload) = SUM(rowid))",
                    "forms_resident": args.payload_form, "raw_bytes": rows * 8 if args.payload_form != "packed" else 0,
                    "packed_bytes": packed_bytes,
                    "note": "FOR-bit-packed in blocks of 1024 rows (10 bits per value for this column); the library "
                            "chooses the form per query; see payload_24bit for a column that packs to 24 bits"},
        "payload_24bit": payload24,
        "sweep": sweep,
        "aggregate_reduce": "device limbs (cubit_gpu_result_add_limbs) + one NCCL all-reduce per step" if world > 1
                            else "device limbs (cubit_gpu_result_add_limbs)",
        "e2e": {"value": e2e_value, "unit": "rows/s",
                "h2d_bytes_per_step": n_q * scan_args_bytes,
                "d2h_bytes_per_step": n_q * (32 + (1 + len(val_host)) * 8 * chunk),
                "path": "aggregate push-down + first DataChunk",
                "note": "synchronous cubit_gpu_query + cubit_gpu_fetch of the first 2048-row DataChunk per query; "
                        "host predicate structs in, COUNT/SUM row + chunk out.  This is the aggregate-push-down path: "
                        "the row-returning numbers (every row ID and value over PCIe) are e2e_full_materialize "
                        "(8-byte copies) and e2e_full_materialize_narrow_wire (delivered to host consumers)"},
        "e2e_full_materialize": e2e_full,
        "e2e_full_materialize_narrow_wire": e2e_drain,
        "rowid_gather": rowid_gather,
        "gpu_launches": launches, "clocks": clocks, "index_build_s": build_s,
    }
    if world == 1 and not args.no_cpu_baseline and not cfg5:
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
    ap.add_argument("--drain-window", type=int, default=0, help="rows per narrow-wire window (0 = library default)")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu child that measures DRAM bytes per kernel")
    ap.add_argument("--no-payload24", action="store_true", help="skip the 24-bit payload raw/packed comparison")
    ap.add_argument("--payload-form", default="both", choices=["both", "raw", "packed"],
                    help="forms of the payload column kept in HBM (both: the library picks per query)")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --rows rows per GPU; strong: ONE table of --rows rows cut over the GPUs")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg5"],
                    help="cfg2: the selectivity sweep (the metric's config); cfg5: 8 predicates on a 16e9-row table")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
