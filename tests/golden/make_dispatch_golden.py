"""Writes tests/golden/dispatch_streams.json from the UNMODIFIED reference host library (oracle/_ref/libnrd_ref.so).
Run in the build container, where /root/reference is mounted:  python tests/golden/make_dispatch_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nrd_sample_b200 import nrd_api as api  # noqa: E402
from oracle import runner  # noqa: E402
from tests.test_dispatch_stream import CASES, GOLDEN, capture  # noqa: E402

runner.build()
ref = api.NrdLibrary(runner.REF_LIB_PATH)
out = {label: capture(ref, den, w, h, kind) for label, den, w, h, kind in CASES}
json.dump(out, open(GOLDEN, "w"), separators=(",", ":"))
print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")
