#!/usr/bin/env python3
"""Latency / throughput of ctmr_process_batch (host buffers) for the batch sizes a Go batcher would use
(entryChan depth 16384, cmd/ct-fetch/ct-fetch.go:132).  Run on a GPU box."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np
from ct_mapreduce_b200 import capi, engine

NOW_NS = 1767225600 * 10**9
cfg = capi.synth_cfg(1 << 20, dup_mode=1)
import torch
blob_d, offs_d, idx_d, total = engine.synth_corpus_device(cfg, 0, 1 << 20, "cuda:0")
iblob, ioffs = engine.synth_issuers(cfg)
hb = capi.PinnedBuffer(total + 64); hblob = hb.view(np.uint8, total); torch.from_numpy(hblob).copy_(blob_d[:total])
offs = offs_d.cpu().numpy().astype(np.uint64); idx = idx_d.cpu().numpy().astype(np.uint32)
db = engine.GpuCertDatabase(table_capacity=1 << 24, issuer_cn_filter=b"Let's Encrypt, ISRG", max_issuers=1024)
print("%10s %12s %14s %12s" % ("batch", "ms/call", "entries/s", "GB/s in"))
for n in (256, 1024, 4096, 16384, 65536, 262144, 1048576):
    reps = max(3, min(200, (1 << 21) // n))
    o = offs[: n + 1]
    res = None
    for it in range(reps + 3):
        if it == 3:
            t0 = time.perf_counter()
        res = db.store_batch(hblob, o, iblob, ioffs, idx[:n], NOW_NS, out=res)
    dt = (time.perf_counter() - t0) / reps
    print("%10d %12.3f %14.0f %12.2f" % (n, dt * 1e3, n / dt, float(o[n]) / dt / 1e9))
