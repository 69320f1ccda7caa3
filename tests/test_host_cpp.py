"""The C++ host mirror (ct_mapreduce_b200/host/ctmr_storage.hpp): the reference's plug-in surface for
this path (RemoteCache / StorageBackend / KnownCertificates / IssuerMetadata / the Store facade).

CPU part: the reference's own reducer tests transcribed to the C++ classes (tests/cpp/host_tests.cpp).
GPU part: GpuCertDatabase::StoreBatch over MockRemoteCache + MockBackend must leave exactly the
state the reference would have left -- computed here from the oracle, i.e. the sequential
insertCTWorker -> Store semantics -- while issuing only one SetInsert per NEW certificate."""
import base64
import os
import subprocess

import numpy as np
import pytest

from conftest import NOW_NS, README_FILTER, ROOT


@pytest.fixture(scope="module")
def host_bin(tmp_path_factory):
    from ct_mapreduce_b200 import build
    build.build()
    out = tmp_path_factory.mktemp("hostbin") / "host_tests"
    libdir = os.path.join(ROOT, "ct_mapreduce_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-o", str(out), os.path.join(ROOT, "tests", "cpp", "host_tests.cpp"),
                    "-L" + libdir, "-lctmr", "-Wl,-rpath," + libdir], check=True)
    return str(out)


def test_reference_reducer_tests_on_cpp_host(host_bin):
    r = subprocess.run([host_bin, "cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "7 passed" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["per_entry", "pipelined", "group_of_three_shards"])
def test_store_batch_drives_remote_cache_like_the_reference(host_bin, ora, tmp_path, mode):
    n = 6000
    cfg = ora.synth_cfg(n, len_mode=1, len_lo=512, len_hi=4096, dup_mode=1)
    blob, offs, idx = ora.synth_corpus(cfg, 0, n)
    iblob, ioffs = ora.synth_issuers(cfg)
    idx = idx.copy()
    idx[::97] = 0xFFFFFFFF  # some entries without a chain
    for name, arr in (("blob", blob), ("offsets", offs), ("issuer_blob", iblob), ("issuer_offsets", ioffs), ("issuer_idx", idx),
                      ("now_ns", np.array([NOW_NS], np.int64)), ("filter", np.frombuffer(README_FILTER, np.uint8))):
        np.ascontiguousarray(arr).tofile(tmp_path / f"{name}.bin")
    # "pipelined": the cache also implements BatchRemoteCache (SetInsertBatch / ExpireAtBatch): same final state,
    # but the serial inserts and expiries of a batch travel in one round trip each
    # "group_of_three_shards": the same StoreBatch over ctmr_group_* (three shards on device 0, 512-entry rounds)
    extra = {"per_entry": [], "pipelined": ["pipelined"], "group_of_three_shards": ["group"]}[mode]
    r = subprocess.run([host_bin, "gpu", str(tmp_path), "3"] + extra, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

    # what the reference would have left behind (sequential Store semantics from the oracle)
    want = ora.DB(README_FILTER, False).process(blob, offs, iblob, ioffs, idx, NOW_NS)
    issuer_ids = {}
    for k in range(cfg.n_issuers):
        der = iblob[ioffs[k]:ioffs[k + 1]].tobytes()
        rc, c = ora.parse_cert(der)
        issuer_ids[k] = ora.issuer_id(der[c.spki_off:c.spki_off + c.spki_len])[1]
    exp_sets, exp_expire, exp_dirty, exp_alloc, exp_pems = set(), {}, set(), set(), {}
    for i in np.nonzero(want.status == 0)[0]:
        hour = int(want.exp_hour[i])
        key = ora.serials_key(hour, issuer_ids[int(idx[i])])
        a = int(offs[i]) + int(want.serial_off[i])
        serial = blob[a:a + int(want.serial_len[i])].tobytes()
        exp_sets.add((key, serial.hex()))
        exp_dirty.add(ora.day_id(hour * 3600))
        if want.was_unknown[i]:
            exp_expire[key] = hour * 3600
            sid = base64.urlsafe_b64encode(serial).decode()
            exp_pems[f"{ora.expdate_id(hour)}/{issuer_ids[int(idx[i])]}/{sid}"] = blob[offs[i]:offs[i + 1]].tobytes()
        if want.first_issuer_hour[i]:
            exp_alloc.add((ora.expdate_id(hour), issuer_ids[int(idx[i])]))

    # IssuerMetadata string sets (issuermetadata.go:92-138): DN via pkix.Name.String(), http(s) CRL-DPs only
    import warnings
    from cryptography import x509
    warnings.filterwarnings("ignore")
    cn_of = {0: "Let's Encrypt Authority X%d", 1: "\\ ISRG Root X%d", 2: "ISRG Root X%d", 3: "Synth Trust Services CA %d"}
    exp_str = set()
    for i in np.nonzero(want.was_unknown)[0]:
        k = int(idx[i])
        iid = issuer_ids[k]
        exp_str.add(("issuer::" + iid, ("CN=" + cn_of[k % 4] % k + ",O=Synth CA %d,C=US" % k).encode().hex()))
        cert = x509.load_der_x509_certificate(blob[offs[i]:offs[i + 1]].tobytes())
        dp = cert.extensions.get_extension_for_oid(x509.oid.ExtensionOID.CRL_DISTRIBUTION_POINTS).value[0].full_name[0].value
        if dp.startswith("http"):
            exp_str.add(("crl::" + iid, dp.encode().hex()))

    got_sets, got_expire, got_dirty, got_alloc, got_pemkeys, got_pems, stats = set(), {}, set(), set(), set(), {}, {}
    got_str = set()
    for line in open(tmp_path / "state.txt"):
        f = line.rstrip("\n").split(" ")
        if f[0] == "STATS":
            stats = dict(zip(f[1::2], map(int, f[2::2])))
        elif f[0] == "SET" and f[1].startswith("serials::"):
            got_sets.add((f[1], f[2]))
        elif f[0] == "STR":
            got_str.add((f[1], f[2]))
        elif f[0] == "EXPIRE":
            got_expire[f[1]] = int(f[2])
        elif f[0] == "DIRTY":
            got_dirty.add(f[1])
        elif f[0] == "ALLOC":
            got_alloc.add((f[1], f[2]))
        elif f[0] == "PEMKEY":
            got_pemkeys.add(f[1])
        elif f[0] == "PEM":
            got_pems[f[1]] = line.rstrip("\n").split(" ", 2)[2]
    assert got_sets == exp_sets
    assert got_expire == exp_expire
    assert got_str == exp_str
    assert got_dirty == exp_dirty
    assert got_alloc == exp_alloc
    assert got_pemkeys == set(exp_pems)
    for k, one_line in got_pems.items():  # PEM body decodes back to the DER (pem.EncodeToMemory)
        body = "".join(one_line.split("|")[1:-2])
        assert base64.b64decode(body) == exp_pems[k]
    n_unknown = int(want.was_unknown.sum())
    assert stats["entries"] == n and stats["stored"] == int((want.status == 0).sum()) and stats["unknown"] == n_unknown
    # the point of the exercise: one cache round trip per NEW certificate instead of one per stored entry
    # (+ one "issuer::<id>" / "crl::<id>" insert per distinct string, from IssuerMetadata's memo)
    n_issuers_seen = len({int(idx[i]) for i in np.nonzero(want.was_unknown)[0]})
    assert stats["set_insert_calls"] == n_unknown + len(exp_str) < stats["stored"]  # + one per distinct DN / CRL string
    assert stats["pem_writes"] == n_unknown
    assert stats["round_trips"] == (2 * 3 if mode == "pipelined" else 0)  # 3 batches x (one SADD pipeline + one EXPIREAT pipeline)
    # the GPU's first-seen bits keep the host's string work at O(issuers), not O(new certificates)
    assert stats["dn_formats"] == n_issuers_seen and stats["crl_parses"] <= 2 * n_issuers_seen < n_unknown
