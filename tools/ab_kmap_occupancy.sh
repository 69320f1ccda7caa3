# needs the experiments build (the product library carries one instantiation of the streaming kernel)
export CTMR_EXPERIMENTS=1
python -m ct_mapreduce_b200.build --force > /dev/null || exit 1
# K_map alone (4 M x 1.5 KB, map_device only) at different residency shapes; see DESIGN.md "Measured and rejected"
for v in "CTMR_MAP_CHUNK=128" "CTMR_MAP_CHUNK=64" "CTMR_MAP_CHUNK=128 CTMR_MAP_WARPS=6"; do echo "== $v"; env $v CTMR_FUSE_INSERT=0 python - <<'PY' 2>&1 | tail -1
import os, sys
sys.path.insert(0, ".")
import torch
from ct_mapreduce_b200 import capi, engine
n = 4_000_000
cfg = capi.synth_cfg(n)
dev = torch.device("cuda:0")
blob, offs, idx, total = engine.synth_corpus_device(cfg, 0, n, dev)
db = engine.GpuCertDatabase(table_capacity=1 << 23)
st = torch.empty(n, dtype=torch.uint8, device=dev); sha = torch.empty((n, 32), dtype=torch.uint8, device=dev)
eh = torch.empty(n, dtype=torch.int64, device=dev); keys = torch.empty((n, 64), dtype=torch.uint8, device=dev)
b = capi.DevBatch(); b.blob, b.blob_bytes, b.offsets, b.n = blob.data_ptr(), total, offs.data_ptr(), n
b.issuer_idx, b.first_index, b.now_unix_ns = idx.data_ptr(), 0, 1767225600 * 10**9
o = capi.DevOut(st.data_ptr(), sha.data_ptr(), eh.data_ptr(), None, None, None, None, keys.data_ptr())
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for _ in range(3): db.map_device(b, o, s.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(5): db.map_device(b, o, s.cuda_stream)
e1.record(s); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("K_map alone: %.3f ms per 4M entries = %.0f GB/s" % (ms, total / ms / 1e6))
PY
done
