"""Known-answer tests of the bit-exact building blocks, on the CPU.

Vectors: Random123's published threefry2x64-20 KATs (zeros, ones, pi digits) plus rows
generated from the reference's vendored Random123/threefry.h with the reference's keying
(ctr = {counter, 0}, key = {pkey, master_key}; omp3/neutral.c:632-652) - SURVEY.md 4.3.
The oracle port and the host build of the kernels' own source must both reproduce them.
"""
import ctypes as C
import math

import numpy as np
import pytest

from neutral_b200.decks import cross_section_table

# (c0, c1, k0, k1) -> (out0, out1)
THREEFRY_KAT = [
    ((0, 0, 0, 0), (0xc2b6e3a8c2c69865, 0x6f81ed42f350084d)),
    ((1, 0, 0, 0), (0xbaf51c00fb3a5957, 0xed553e57f10b3b42)),
    ((0, 0, 0, 1), (0x3386564ed9e958da, 0x5ec3797e073ce882)),
    ((0, 0, 1, 1), (0x23e5a526416cfd26, 0x0f866f9cf277ba2f)),
    ((7, 0, 123456, 3), (0x25ad74e65018596a, 0xcf56ec78d6e8a88a)),
    ((2, 0, 999999, 10), (0x44ca4c11b022d8cd, 0xda8150beb41400ea)),
    ((2**64 - 1, 2**64 - 1, 2**64 - 1, 2**64 - 1), (0xe02cb7c4d95d277a, 0xd06633d0893b8b68)),
    ((0x243f6a8885a308d3, 0x13198a2e03707344, 0xa4093822299f31d0, 0x082efa98ec4e6c89),
     (0x263c7d30bb0f0af1, 0x56be8361d3311526)),
]

# rn0, rn1 of the first six rows and glibc 2.39's -log(rn0) of the first four
UNIT_KAT = [
    ("0x1.856dc751858d3p-1", "0x1.be07b50bcd402p-2"),
    ("0x1.75ea3801f674bp-1", "0x1.daaa7cafe2167p-1"),
    ("0x1.9c32b276cf4acp-3", "0x1.7b0de5f81cf3ap-2"),
    ("0x1.1f2d29320b67fp-3", "0x1.f0cdf39e4ef74p-5"),
    ("0x1.2d6ba73280c2dp-3", "0x1.9eadd8f1add15p-1"),
    ("0x1.13293046c08b6p-2", "0x1.b502a17d6828p-1"),
]
NEGLOG_KAT = ["0x1.1836018e6b89cp-2", "0x1.41d6e6dd59d73p-2", "0x1.9a65c00648e35p+0",
              "0x1.f6eaeecf323a5p+0"]

# bracketing index of the reference table for a few energies
CS_INDEX_KAT = {1.0: 298, 1.0e3: 1685, 1.0e4: 2998, 2.5e4: 3771, 1.0e6: 9485}


def _host_threefry(lib, c0, c1, k0, k1):
    out = (C.c_uint64 * 2)()
    lib.nb200_host_threefry2x64_20(c0, c1, k0, k1, out)
    return int(out[0]), int(out[1])


@pytest.mark.parametrize("args,expect", THREEFRY_KAT)
def test_threefry_kat_oracle(port, args, expect):
    assert port.threefry(*args) == expect


@pytest.mark.parametrize("args,expect", THREEFRY_KAT)
def test_threefry_kat_kernel_source(lib, args, expect):
    assert _host_threefry(lib, *args) == expect


def test_unit_interval_and_neglog(port, lib):
    for (args, _), (h0, h1) in zip(THREEFRY_KAT[:6], UNIT_KAT):
        c0, _, k0, k1 = args
        r0, r1 = port.random_pair(k0, k1, c0)
        assert r0 == float.fromhex(h0) and r1 == float.fromhex(h1)
    for (args, _), h in zip(THREEFRY_KAT[:4], NEGLOG_KAT):
        c0, _, k0, k1 = args
        r0, _ = port.random_pair(k0, k1, c0)
        assert -math.log(r0) == float.fromhex(h)          # the libm the oracle links
        assert -lib.nb200_host_log(r0) == float.fromhex(h)  # the kernels' transliteration


def test_cs_index_kat(port):
    keys, values = cross_section_table()
    assert len(keys) == 29999
    assert np.all(np.diff(keys) > 0)
    for e, ind in CS_INDEX_KAT.items():
        assert port.cs_index(keys, e) == ind
        assert keys[ind] <= e < keys[ind + 1]
        t = (e - keys[ind]) / (keys[ind + 1] - keys[ind])
        assert port.cs_lookup(keys, values, e) == values[ind] + t * (values[ind + 1] - values[ind])
