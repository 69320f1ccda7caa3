import base64
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
NOW_SEC = 1767225600  # 2026-01-01T00:00:00Z, the corpus' fixed "now" 
NOW_NS = NOW_SEC * 10**9
README_FILTER = b"Let's Encrypt, ISRG"  # reference README.md:29, split without trimming


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pem_to_der(path):
    body = "".join(l.strip() for l in open(path) if "-----" not in l)
    return base64.b64decode(body)


@pytest.fixture(scope="session")
def golden():
    fx = json.load(open(os.path.join(GOLDEN, "fixtures.json")))
    for name in fx:
        fx[name]["der"] = pem_to_der(os.path.join(GOLDEN, name + ".pem"))
    return fx


@pytest.fixture(scope="session")
def ora():
    from oracle import oracle
    oracle.lib()
    return oracle


def pack(ders):
    """list of bytes -> (blob u8, offsets u64)"""
    offsets = np.zeros(len(ders) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(d) for d in ders], dtype=np.uint64)
    blob = np.frombuffer(b"".join(ders), np.uint8).copy() if ders else np.zeros(0, np.uint8)
    return blob, offsets
