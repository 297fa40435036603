"""bench.py's host-side helpers (no GPU): the clock-sampler row parser and the verdict packing used for the gathered-mask check."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["bench_module"] = m
    spec.loader.exec_module(m)
    return m


def test_clock_sampler_parses_rows_and_reasons():
    b = _bench()
    s = b.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    s.marks = [10.0, 10.2]
    ok = "0, 1965, 1965, 0x0000000000000001, Not Active, Not Active, Not Active, Not Active"
    capped = "0, 1800, 1965, 0x0000000000000004, Not Active, Not Active, Not Active, Active"
    s.rows = [(9.0, ok), (10.05, ok), (10.1, capped), (10.15, ok), (11.0, "0, 600, 1965, 0x1, Not Active, Not Active, Not Active, Not Active"),
              (10.12, "garbage")]
    out = s.stop()
    assert out["sm_max_mhz"] == 1965.0
    assert out["samples"] == 3 and out["samples_total"] == 5
    assert out["sm_mhz"] == 1965.0                      # median of the three samples inside the timed region
    assert out["reasons"] == ["sw_power_cap"]           # only reasons seen inside the region count


def test_pack_bits_matches_the_device_layout():
    b = _bench()
    from consensus_b200 import sharding
    rng = np.random.default_rng(3)
    ok = (rng.integers(0, 2, 65536)).astype(np.uint8)
    words = b.pack_bits(ok)
    assert words.dtype == np.uint32 and words.size == 2048
    assert np.array_equal(words, sharding.pack_bits(ok, 2048))          # bit i of word i/32, as k_pack_bits writes it
    assert np.array_equal(sharding.unpack_bits(words, ok.size), ok)
