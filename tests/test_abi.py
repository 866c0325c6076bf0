"""The C-ABI library loads and exports every function include/lasgun_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "lasgun_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lgb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(native):
    names = declared_functions()
    assert len(names) >= 16
    lib = ctypes.CDLL(native.SO_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(native.ABI_SYMBOLS) == names


def test_struct_sizes_match_header(native):
    assert ctypes.sizeof(native.Node) == 32
    assert ctypes.sizeof(native.CameraDesc) == 12 * 8 + 3 * 8 + 8
    assert ctypes.sizeof(native.Stats) == 21 * 8 + 12 * 4 + 8
    assert ctypes.sizeof(native.KernelTime) == 40 + 8 + 11 * 8
    assert ctypes.sizeof(native.SceneDesc) % 8 == 0


def test_no_cpu_fallback(native):
    """Without a GPU lgb_init must fail with LGB_ERR_NO_DEVICE, never silently render on the CPU."""
    L = native.lib()
    if L.lgb_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(native.LasgunError) as e:
        native.Context(0)
    assert e.value.status == native.LGB_ERR_NO_DEVICE
    assert b"no sm_100" in L.lgb_status_string(native.LGB_ERR_NO_DEVICE)


def test_product_does_not_touch_oracle():
    """Nothing under lasgun_b200/ or include/ may import, include or link the oracle."""
    bad = []
    for base in ("lasgun_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(import|from)\s+oracle|pyoracle|lasgun_oracle|oracle/", txt) and "does not include or link anything under oracle" not in txt:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_host_block_cache_reuses_freed_blocks(native):
    """The host mirror's arrays (include/lasgun_host.hpp) come from lgh_block_alloc: a freed block of a megabyte or more is handed out
    again for a request it fits (capture(scene, film) sizes the same arrays every frame), small requests bypass the cache."""
    lib = ctypes.CDLL(native.SO_PATH)
    lib.lgh_block_alloc.restype = ctypes.c_void_p
    lib.lgh_block_alloc.argtypes = [ctypes.c_size_t]
    lib.lgh_block_free.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    big = 5 << 20
    p = lib.lgh_block_alloc(big)
    assert p
    ctypes.memset(p, 0x5A, big)
    lib.lgh_block_free(p, big)
    got = []                                         # (other tests of this process may have left fitting blocks in the cache: at most 64)
    for _ in range(70):
        got.append(lib.lgh_block_alloc(big - 4096))  # rounds to the same 2 MB multiple: the cached block fits
        if got[-1] == p:
            break
    assert p in got and all(got)
    for q in got:
        lib.lgh_block_free(q, big - 4096)
    small = lib.lgh_block_alloc(1000)
    assert small
    lib.lgh_block_free(small, 1000)
