#!/usr/bin/env python3
"""Measured INT-pipe ceiling of SHA-256 on this GPU (register-only, K_map's compression function)."""
import sys
sys.path.insert(0, ".")
from ct_mapreduce_b200 import engine
db = engine.GpuCertDatabase(table_capacity=1 << 12)
print("%8s %6s %10s" % ("variant", "warps", "GB/s"))
for rolled in (True, False):
    for ctas in (1, 2, 4, 8):
        gbs, ms = db.sha256_ceiling(iters=3000, rolled=rolled, ctas_per_sm=ctas)
        print("%8s %6d %10.0f   (%.2f ms)" % ("rolled" if rolled else "unrolled", ctas * 8, gbs, ms))
