"""CPU tests of the boundary: librsfusion.so loads, exports every symbol include/rsfusion.h declares,
host-callable functions agree with the oracle, and the product fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib
from conftest import ROOT, have_gpu


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rsfusion.h")).read()
    return sorted(set(re.findall(r"RSF_API\s+[\w\s\*]+?\b(rsf_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(rsf):
    lib = rsf.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 26
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in rsfusion.h but not exported"
        assert name in rsf.ABI, f"{name} has no ctypes binding"
    out = subprocess.check_output(["nm", "-D", "--defined-only", rsf.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (rsf_\w+)", out))
    assert exported == set(declared), exported ^ set(declared)
    assert lib.rsf_abi_version() == 2


def test_library_is_sm100a_native(rsf):
    out = subprocess.check_output(["cuobjdump", "-lelf", rsf.LIB_PATH], text=True)
    assert "sm_100a" in out


def test_host_hash_matches_oracle_and_golden(rsf, oracle):
    assert rsf.Hash(33, 180, 42) == rsf.Hash(61, 16, 170) == rsf.Hash(63, 171, 45) == 2097151
    assert rsf.Hash(0, 0, 0) == 0
    rng = np.random.default_rng(1)
    for x, y, z in rng.integers(-3000, 3000, size=(500, 3)):
        assert rsf.Hash(int(x), int(y), int(z)) == oracle_lib.Hash(int(x), int(y), int(z))


def test_shard_owner_matches_oracle_and_partitions(rsf, oracle):
    rng = np.random.default_rng(2)
    for world in (2, 4, 8):
        counts = np.zeros(world, int)
        for x, y, z in rng.integers(-500, 500, size=(800, 3)):
            o = rsf.shard_owner(int(x), int(y), int(z), 1, world)
            assert o == oracle.ora_shard_owner(int(x), int(y), int(z), 1, world)
            assert 0 <= o < world
            counts[o] += 1
        assert counts.min() > 0.5 * 800 / world      # reasonably balanced
    assert rsf.shard_owner(5, 6, 7, 1, 1) == 0


def test_default_options(rsf):
    lib = rsf.load_library()
    o = rsf.rsf_options()
    lib.rsf_default_options(C.byref(o))
    assert o.pool_blocks == 1 << 18 and o.table_slots_log2 == 22 and o.max_pixels == 1920 * 1080
    assert o.shard_world == 1 and o.device == -1


def test_invalid_arguments_are_reported(rsf):
    lib = rsf.load_library()
    assert lib.rsf_create(None, None) == rsf.RSF_ERR_INVALID
    assert b"NULL" in lib.rsf_last_error()
    assert lib.rsf_integrate(None, None, None, None, None, 1, 1, 1.0, None, None) == rsf.RSF_ERR_INVALID
    assert lib.rsf_destroy(None) == rsf.RSF_OK


@pytest.mark.skipif(have_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(rsf):
    with pytest.raises(rsf.RsfError) as e:
        rsf.TSDFGrid(0.01, 0.06)
    assert e.value.code == rsf.RSF_ERR_CUDA
    assert "no CPU fallback" in str(e.value)
