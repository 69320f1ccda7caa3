import sys; sys.path.insert(0,'.')
import numpy as np, torch
from oracle import oracle as ora
from ct_mapreduce_b200 import capi, engine as eng
n, first = 300, 1234
cfg_o = ora.synth_cfg(50000); cfg_g = capi.synth_cfg(50000)
blob, offs, idx = ora.synth_corpus(cfg_o, first, n)
dblob, doffs, didx, total = eng.synth_corpus_device(cfg_g, first, n, "cuda:0")
got = dblob[:total].cpu().numpy()
bad = np.nonzero(got != blob)[0]
for b in bad[:40]:
    ci = np.searchsorted(offs, b, side='right')-1
    print("cert", ci, "off", int(b-offs[ci]), "len", int(offs[ci+1]-offs[ci]), "host %02x dev %02x" % (blob[b], got[b]), "issuer", idx[ci])
