#!/usr/bin/env python3
"""bench.py -- CT entries/sec through the B200-native map/reduce hot path.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A "step" is one pass of the hot path over one batch of synthetic CT entries (BASELINE.json
configs[1] at N=1: 10 M x ~1.5 KB DER, SHA-256 fingerprint + KnownCertificates dedup, per-issuer
counts).  `value` is whole-job entries/s with the batch already resident in HBM; `e2e` is the same
metric through the host-buffer C-ABI call (pinned host memory in, host results out, copies inside
the timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NOW_SEC = 1767225600
NOW_NS = NOW_SEC * 10**9
SEED = 20260922

WORKLOADS = {
    # BASELINE.json configs[1]: SHA-256 fingerprint + KnownCertificates dedup (no CN filter, expired kept)
    "cfg2": dict(desc="configs[1]: 10M synthetic ~1.5KB DER certs per GPU, SHA-256 fingerprint + KnownCertificates dedup",
                 n=10_000_000, synth=dict(len_mode=0, len_lo=1436, len_hi=1564, dup_mode=0), filter=b"", log_expired=True,
                 out_bytes=42),
    # configs[2] shape (full map incl. issuerCNFilter + expiry, IssuerMetadata reduce) at a size that fits HBM at once
    "cfg3": dict(desc="configs[2] shape: full map (ASN.1 + issuerCNFilter + expiry) + reducers, 10M entries per GPU chunk",
                 n=10_000_000, synth=dict(len_mode=0, len_lo=1436, len_hi=1564, dup_mode=0), filter=b"Let's Encrypt, ISRG",
                 log_expired=False, out_bytes=54),
    # configs[4] shape: mixed 512 B-8 KB, 50 % duplicates, 256 issuers
    "cfg5": dict(desc="configs[4] shape: mixed 512B-8KB DER, 50% duplicates, 256 issuers, 5M entries per GPU chunk",
                 n=5_000_000, synth=dict(len_mode=1, len_lo=512, len_hi=8192, dup_mode=1), filter=b"Let's Encrypt, ISRG",
                 log_expired=False, out_bytes=54),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--entries", type=int, default=0, help="entries per GPU per step (default: the workload's)")
    ap.add_argument("--no-fingerprint", action="store_true",
                    help="CTMR_F_NO_FINGERPRINT: the reference-faithful path (it never hashes the leaf): parse+filter+dedup+counts")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="entries in the CPU sample (default: auto)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_traffic(entries_per_launch):
    """dram__bytes_read+write of K_map from the committed ncu --set full capture (profiles/), scaled from the
    captured launch size to this run's entries per launch; None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "r1_map_traffic.json")
    try:
        t = json.load(open(p))
        per_entry = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["entries"]
        return per_entry * entries_per_launch, t["source"]
    except Exception:
        return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------- reference arm
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_arm(n_sample, workload, steps, warmup, make_sample):
    """The reference's CPU algorithm (the oracle port: the Go toolchain is absent, oracle/_ref cannot exist)
    over a bounded sample of the same workload, map half on every host thread, reduce half sequential
    exactly like numThreads=1 over MockRemoteCache.  Returns (entries/s, seconds/step)."""
    from oracle import oracle
    blob, offs, idx, iblob, ioffs = make_sample(n_sample)
    thr = host_threads()
    times = []
    for it in range(warmup + steps):
        db = oracle.DB(workload["filter"], workload["log_expired"])
        t0 = time.perf_counter()
        db.process(blob, offs, iblob, ioffs, idx, NOW_NS, nthreads=thr)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        del db
    per = sum(times) / len(times)
    return n_sample / per, per, thr


def cpu_sample_maker(workload, rank_first=0):
    """Sample of the bench corpus for the CPU arm; generated by the CPU generator (no GPU involved)."""
    from oracle import oracle

    def make(n):
        cfg = oracle.synth_cfg(max(n, 2), seed=SEED, **workload["synth"])
        blob, offs, idx = oracle.synth_corpus(cfg, rank_first, n)
        iblob, ioffs = oracle.synth_issuers(cfg)
        return blob, offs, idx, iblob, ioffs
    return make


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    n_sample = args.cpu_sample or 400_000
    v, per, thr = cpu_arm(n_sample, wl, args.steps, max(args.warmup, 1), cpu_sample_maker(wl))
    line = {
        "impl": "reference", "metric": "ct_entries_per_sec", "value": v, "unit": "entries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": wl["desc"], "sample_entries_per_step": n_sample},
        "cpu_baseline": {"value": v, "unit": "entries/s", "cores": thr, "kind": "port",
                         "sample": f"{n_sample} entries of the same corpus per step; map half on {thr} threads, reduce half "
                                   "sequential (reference numThreads=1 over MockRemoteCache); C restatement, not the Go engine"},
        "e2e": {"value": v, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------- our arm
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from ct_mapreduce_b200 import build, capi, engine, sharded

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    build.build()
    wl = WORKLOADS[args.workload]
    n = args.entries or wl["n"]
    K, W = args.steps, args.warmup

    # ---- corpus: rank r owns entries [r*n, (r+1)*n) of one global corpus (weak scaling), generated in HBM
    cfg = capi.synth_cfg(world * n, seed=SEED, **wl["synth"])
    blob, offsets, issuer_idx, total_bytes = engine.synth_corpus_device(cfg, rank * n, n, dev)
    iblob, ioffs = engine.synth_issuers(cfg)
    flags = capi.F_NO_FINGERPRINT if args.no_fingerprint else 0
    db = engine.GpuCertDatabase(device=local, table_capacity=max(1 << 20, 2 * n), issuer_cn_filter=wl["filter"],
                                log_expired_entries=wl["log_expired"], flags=flags, max_issuers=4096)
    dense = db.register_issuers(iblob, ioffs)  # same order on every rank -> same dense indices
    assert (dense == np.arange(cfg.n_issuers)).all()

    status = torch.empty(n, dtype=torch.uint8, device=dev)
    sha = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    exp_hour = torch.empty(n, dtype=torch.int64, device=dev)
    was_unknown = torch.empty(n, dtype=torch.uint8, device=dev)
    first = torch.empty(n, dtype=torch.uint8, device=dev)
    keys = torch.empty((n, capi.KEY_BYTES), dtype=torch.uint8, device=dev)
    ops = sharded.GpuOps(db)
    # CTMR_FIXED_EXCHANGE=1: the sync-free fixed-capacity key exchange (opt-in until it has been measured on >= 2 GPUs)
    fixed_exchange = world > 1 and os.environ.get("CTMR_FIXED_EXCHANGE", "0") == "1"
    red = sharded.ShardedReducer(ops, dev, n_issuer_slots=cfg.n_issuers, max_keys=n if world > 1 else 0,
                                 fixed_capacity=fixed_exchange)
    # Two streams: the INT-bound map half of sub-batch k+1 (stream A) overlaps the latency-bound reduce
    # half -- and, at N>1, the key exchange -- of sub-batch k (stream B).  Every launch, event and
    # collective of the timed region lives on one of them.
    stream = torch.cuda.Stream(dev)      # B: reset, reduce, collectives; also brackets the timed region
    stream_a = torch.cuda.Stream(dev)    # A: K_map
    torch.cuda.synchronize(dev)
    torch.cuda.set_stream(stream)
    # overlap hides the exchange at N>1 (only the LAST sub-batch's reduce chain is exposed, so more, smaller
    # sub-batches shorten the step until launch overheads win); at N=1 the library pipelines internally
    NSUB = int(os.environ.get("CTMR_BENCH_NSUB", "4")) if (world > 1 and n >= (1 << 21)) else 1
    bounds = [n * k // NSUB for k in range(NSUB + 1)]

    step_no = [0]

    def dev_batch(lo, hi):
        b = capi.DevBatch()
        b.blob, b.blob_bytes, b.offsets, b.n = blob.data_ptr(), total_bytes, offsets.data_ptr() + 8 * lo, hi - lo
        b.issuer_idx, b.issuer_map, b.issuer_map_len = issuer_idx.data_ptr() + 4 * lo, None, 0
        b.first_index = (step_no[0] * world + rank) * n + lo
        b.now_unix_ns = NOW_NS
        return b

    def dev_out(lo):
        return capi.DevOut(status.data_ptr() + lo, sha.data_ptr() + 32 * lo, exp_hour.data_ptr() + 8 * lo, None, None, None,
                           None, keys.data_ptr() + capi.KEY_BYTES * lo)

    map_events = []

    fused_out = capi.DevOut(status.data_ptr(), sha.data_ptr(), exp_hour.data_ptr(), None, None, was_unknown.data_ptr(),
                            first.data_ptr(), keys.data_ptr())

    def step(timed):
        db.reset_device(stream.cuda_stream)               # every step sees an empty known-certificate set
        if world == 1:
            # the library's own single-GPU pipeline: K_map with the table insert fused in, then K_resolve / K_pairs
            db.process_device(dev_batch(0, n), fused_out, stream.cuda_stream)
            counts, stat = red.merged_histogram()
            step_no[0] += 1
            return counts, stat
        ev_reset = torch.cuda.Event()
        ev_reset.record(stream)
        stream_a.wait_event(ev_reset)                      # also orders A behind the previous step's reduces
        for k in range(NSUB):
            lo, hi = bounds[k], bounds[k + 1]
            with torch.cuda.stream(stream_a):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream_a)
                ops.map(dev_batch(lo, hi), dev_out(lo))    # K_map: length bucketing + DER walk + filter + SHA-256
                e1.record(stream_a)
            stream.wait_event(e1)
            red.reduce_chunk(keys[lo:hi], hi - lo, was_unknown[lo:hi], first[lo:hi])  # K_insert/K_resolve/K_pairs + exchange
            if timed:
                map_events.append((e0, e1))
        counts, stat = red.merged_histogram()              # one all-reduce of the histograms per chunk
        step_no[0] += 1
        return counts, stat

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(W):
        step(False)
    db.check_device(stream.cuda_stream)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(K):
        counts, stat = step(True)
    t1.record(stream)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    elapsed_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    # K_map's duration per step = sum over the step's sub-batch launches: torch events on stream A at N>1; at N=1
    # the library's own CUDA events around the map stage of the last timed step (ctmr_profile_last)
    if world == 1:
        map_ms = torch.tensor([db.profile_last()[0]], dtype=torch.float64, device=dev)
    else:
        map_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in map_events) / max(K, 1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(map_ms, op=dist.ReduceOp.MAX)
    elapsed_ms, map_ms = float(elapsed_ms.item()), float(map_ms.item())
    db.check_device(stream.cuda_stream)
    if red.check_overflow():
        raise RuntimeError("fixed-capacity key exchange overflowed: rerun without CTMR_FIXED_EXCHANGE")

    # ---- sanity inside the bench: the timed result is the real thing (cheap, size-independent checks)
    n_ok = int((status == 0).sum().item())
    n_unknown_local = int(was_unknown.sum().item())
    tot_counts = int(counts.sum().item())
    tot = torch.tensor([n_unknown_local, n_ok], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    assert tot_counts == int(tot[0].item()), "per-issuer counts must sum to the number of unknown entries"
    assert int(stat[0].item()) == int(tot[1].item()), "status counter OK mismatch"
    if wl["synth"]["dup_mode"] == 0:
        assert n_unknown_local == n_ok, "distinct corpus: every stored entry is unknown"
    # spot-check fingerprints of a few entries against hashlib on the host
    import hashlib
    offs_h = offsets[:9].cpu().numpy()
    blob_h = blob[: int(offs_h[8])].cpu().numpy()
    sha_h = sha[:8].cpu().numpy()
    for i in range(8):
        if not args.no_fingerprint:
            assert sha_h[i].tobytes() == hashlib.sha256(blob_h[offs_h[i]:offs_h[i + 1]].tobytes()).digest(), "fingerprint mismatch"

    value = world * n * K / (elapsed_ms / 1e3)
    peak, peak_src = measured_peaks()
    alg_bytes = total_bytes + wl["out_bytes"] * n  # sum(L_i) + 42*N (cfg2) / 54*N (cfg3-5), per launch
    achieved = alg_bytes / (map_ms / 1e3) / 1e9
    # N=1: 4 sub-batches x (len_order(3) + map(with insert) + resolve + pairs); N>1: + insert, partition(3), scatter
    # measured INT-pipe ceiling of the fingerprint on this GPU: register-only SHA-256 at K_map's occupancy
    int_ceiling_gbs = db.sha256_ceiling(iters=2000, rolled=True, ctas_per_sm=2)[0] if not args.no_fingerprint else None
    launches_per_step = 4 * 6 if world == 1 else NSUB * (4 + 3 + 4)
    map_launches = 4 if world == 1 else NSUB
    traffic, traffic_src = measured_traffic(n / map_launches)

    # ---- e2e through the host-buffer C ABI (pinned host memory -> results on the host)
    e2e = None
    if not args.no_e2e:
        ne = n
        hb = capi.PinnedBuffer(total_bytes + 64)
        ho = capi.PinnedBuffer((ne + 1) * 8)
        hi = capi.PinnedBuffer(ne * 4)
        out_bufs = {k: capi.PinnedBuffer(ne * sz) for k, sz in (("status", 1), ("sha256", 32), ("exp_hour", 8),
                                                                   ("was_unknown", 1), ("first", 1))}
        hblob = torch.from_numpy(hb.view(np.uint8, total_bytes))
        hblob.copy_(blob[:total_bytes])
        torch.from_numpy(ho.view(np.int64, ne + 1)).copy_(offsets)
        torch.from_numpy(hi.view(np.int32, ne)).copy_(issuer_idx)
        torch.cuda.synchronize(dev)
        res = engine.BatchResult(out_bufs["status"].view(np.uint8), out_bufs["sha256"].view(np.uint8).reshape(ne, 32),
                                 out_bufs["exp_hour"].view(np.int64), None, None, out_bufs["was_unknown"].view(np.uint8),
                                 out_bufs["first"].view(np.uint8))
        h2d = total_bytes + (ne + 1) * 8 + ne * 4
        d2h = ne * (1 + 32 + 8 + 1 + 1)

        def e2e_step():
            db.reset_device(None)
            o = capi.Out(res.status.ctypes.data, res.sha256.ctypes.data, res.exp_hour.ctypes.data, None, None,
                         res.was_unknown.ctypes.data, res.first_issuer_hour.ctypes.data)
            rc = db._lib.ctmr_process_batch(db.handle, hb.addr, ho.addr, ne, iblob.ctypes.data, ioffs.ctypes.data,
                                            cfg.n_issuers, hi.addr, NOW_NS, C.byref(o))
            db._check(rc)

        for _ in range(min(W, 2)):
            e2e_step()
        barrier()
        ts = time.perf_counter()
        for _ in range(K):
            e2e_step()
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - ts], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert args.no_fingerprint or np.array_equal(res.sha256[:8], sha_h), "e2e fingerprints differ from the device-resident run"
        e2e = {"value": world * ne * K / float(dt.item()), "unit": "entries/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h),
               "note": "ctmr_process_batch, pinned host buffers, 3-stage H2D/kernel/D2H pipeline"
                       + ("; per-rank known-certificate sets (no cross-GPU routing on the host-buffer call)" if world > 1 else "")}

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the box's host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_sample = args.cpu_sample or 400_000
        v, per, thr = cpu_arm(n_sample, wl, 1, 1, cpu_sample_maker(wl))
        cpu = {"value": v, "unit": "entries/s", "cores": thr, "kind": "port",
               "sample": f"first {n_sample} entries of the same corpus, 1 timed pass ({per:.2f} s); map half on {thr} threads, "
                         "reduce half sequential like the reference's numThreads=1; C restatement, not the Go engine"}

    if rank == 0:
        line = {
            "metric": "ct_entries_per_sec", "value": value, "unit": "entries/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": wl["desc"] + (" [fingerprint OFF: reference-faithful path]" if args.no_fingerprint else ""),
                       "entries_per_gpu_per_step": n, "bytes_per_gpu_per_step": int(total_bytes),
                       "issuers": cfg.n_issuers, "l2": "inputs (>=7 GB per step) far exceed the 126 MB L2; no explicit flush",
                       "parallelism": f"entry-index shards x{world}, key routing by hash(expDate, issuer), 1 all-reduce/chunk"
                                      if world > 1 else "single GPU"},
            "sha256_gbs": total_bytes * world * K / (elapsed_ms / 1e3) / 1e9,
            "roofline": {"bound": "hbm", "kernel": "map_stream_kernel (+ its 3 length-bucketing helper launches)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel_ms_per_step": map_ms, "kernel_ms": map_ms / map_launches,
                         "algorithmic_bytes_per_step": int(alg_bytes), "algorithmic_bytes_per_launch": int(alg_bytes / map_launches),
                         "launches_per_step": map_launches,
                         "sha256_int_ceiling": None if int_ceiling_gbs is None else {
                             "value": int_ceiling_gbs, "unit": "GB/s of message bytes",
                             "how": "register-only SHA-256 microbenchmark (ctmr_sha256_ceiling_device), 16 warps/SM, run in this process",
                             "k_map_frac": (total_bytes / (map_ms / 1e3) / 1e9) / int_ceiling_gbs},
                         "note": "SHA-256 is INT-pipe bound (DESIGN.md): see profiles/ for ALU-pipe utilisation; timed while the "
                                 "reduce kernels of the previous sub-batch share the GPU"},
            "gpu_launches": launches_per_step * K,
            "clocks": clk,
        }
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
