"""Throughput of the CT wire-format front end (ctmr_process_raw) on one GPU.

    python tools/bench_frontend.py [--entries 200000] [--steps 5] [--warmup 2]

Synthesises get-entries pages (2/3 x509 entries, 1/3 precert entries, chains of one or two certificates) around
the synthetic corpus with libctmr's threaded host generator, directly in pinned host memory, and times whole ctmr_process_raw calls (host text in,
host results out).  Prints one JSON line: end-to-end entries/s, and the library's own CUDA-event split of a call
into front-end kernels (base64, framing, Chain[0] identification) and the map/reduce path behind them.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ct_mapreduce_b200 import capi, engine, frontend as fe  # noqa: E402

NOW_NS = 1767225600 * 10**9


def make_pages(n, page=1000):
    """n synthetic entries as get-entries bodies, written by libctmr's host generator (ctmr_synth_raw_pages_host)
    straight into pinned memory; tests/test_frontend_oracle.py checks that generator against the Python encoders."""
    cfg = capi.synth_cfg(n)
    L = capi.load()
    need = L.ctmr_synth_raw_pages_host(C.byref(cfg), 0, n, page, None, 0, None, None, None, None)
    pin = capi.PinnedBuffer(need)
    lo, ll = np.zeros(n, np.uint64), np.zeros(n, np.uint32)
    xo, xl = np.zeros(n, np.uint64), np.zeros(n, np.uint32)
    got = L.ctmr_synth_raw_pages_host(C.byref(cfg), 0, n, page, pin.addr, need, capi.ptr(lo), capi.ptr(ll), capi.ptr(xo), capi.ptr(xl))
    assert got == need
    return pin, lo, ll, xo, xl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--entries", type=int, default=200000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    t0 = time.time()
    pin, lo, ll, xo, xl = make_pages(a.entries)
    gen_s = time.time() - t0
    tv = pin.view()
    text = tv
    chars = int(ll.sum()) + int(xl.sum())
    with engine.GpuCertDatabase(log_expired_entries=True, table_capacity=1 << 22) as db:
        fe_ms = path_ms = 0.0
        for it in range(a.warmup + a.steps):
            if it == a.warmup:
                t0 = time.perf_counter()
            r = db.store_raw_entries(tv, lo, ll, xo, xl, NOW_NS)
            if it >= a.warmup:
                f, p, launches = db.frontend_profile_last()
                fe_ms += f
                path_ms += p
        dt = (time.perf_counter() - t0) / a.steps
        ok = int((r.entry_status == 0).sum())
    assert ok == a.entries, (ok, a.entries)
    print(json.dumps({
        "metric": "ct_raw_entries_per_sec", "value": a.entries / dt, "unit": "entries/s", "entries": a.entries, "steps": a.steps,
        "ms_per_call": dt * 1e3, "text_bytes": len(text), "base64_chars": chars,
        "h2d_gbs": len(text) / dt / 1e9,
        "frontend_kernels_ms": fe_ms / a.steps, "frontend_kernels_chars_gbs": chars / (fe_ms / a.steps) / 1e6,
        "frontend_kernels_entries_per_sec": a.entries / (fe_ms / a.steps) * 1e3,
        "path_ms": path_ms / a.steps, "frontend_launches": launches, "page_synthesis_s": gen_s,
        "note": "value = whole ctmr_process_raw calls, pinned host text in, host results out, synchronous chunks",
    }))


if __name__ == "__main__":
    main()
