"""CPU: the C-ABI library loads and exports every symbol include/ssd_b200.h declares; the three
Philox4x32-10 implementations (oracle numpy, oracle C, library host helper) agree with the published
known-answer vectors (Random123 kat_vectors: philox4x32 10 rounds)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "ssd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from contracts_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "libssd_b200.so does not export %s" % missing
    assert sorted(_lib.EXPORTS) == names, "contracts_b200/_lib.py EXPORTS out of sync with include/ssd_b200.h"
    assert lib.ssd_abi_version() == _lib.SSD_ABI_VERSION


def test_config_struct_matches_header():
    """ctypes mirror of ssd_config / ssd_step_io has the field order of the header."""
    from contracts_b200 import _lib
    text = open(os.path.join(ROOT, "include", "ssd_b200.h")).read()
    for struct, cls in (("ssd_config", _lib.ssd_config), ("ssd_step_io", _lib.ssd_step_io),
                        ("ssd_selfdrive_io", _lib.ssd_selfdrive_io), ("ssd_feat_io", _lib.ssd_feat_io),
                        ("ssd_host_layout", _lib.ssd_host_layout)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                fields += [re.sub(r"\[.*\]|[\*\s]", "", f.split()[-1]) for f in decl.split(",")]
        assert fields == [f[0] for f in cls._fields_], struct


def test_create_rejects_stale_config_struct():
    """abi_version and struct_size are checked before anything else (no GPU needed): a binding whose ssd_config
    declaration is shorter than the library's must not be read past its end."""
    from contracts_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    cfg = _lib.make_config(env_kind=0, num_envs=4, num_agents=2, map_h=1, map_w=1, ascii_map=b"P", horizon=10)
    cfg.struct_size = ctypes.sizeof(_lib.ssd_config) - 64            # e.g. a stub that stops before env_params
    assert lib.ssd_create(ctypes.byref(cfg), ctypes.byref(h)) == -1 and not h.value
    assert b"struct_size" in lib.ssd_last_error(None)
    cfg = _lib.make_config(env_kind=0, num_envs=4, num_agents=2, map_h=1, map_w=1, ascii_map=b"P", horizon=10)
    cfg.abi_version = _lib.SSD_ABI_VERSION - 1
    assert lib.ssd_create(ctypes.byref(cfg), ctypes.byref(h)) == -1 and not h.value
    assert b"abi_version" in lib.ssd_last_error(None)


def test_integration_md_stub_matches_binding():
    """The ctypes stub printed in INTEGRATION.md declares ssd_config with the same fields as the shipped binding."""
    from contracts_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class ssd_config\(ctypes\.Structure\):(.*?)\n\n", text, flags=re.S)
    assert m, "INTEGRATION.md must show the ssd_config ctypes stub"
    names = re.findall(r'\("([a-z_0-9]+)",', m.group(1))
    assert names == [f[0] for f in _lib.ssd_config._fields_]


def test_no_cpu_fallback_without_gpu():
    import torch
    from contracts_b200 import _lib
    from contracts_b200.batched import BatchedGridEnv
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.SsdError):
        BatchedGridEnv("cleanup_new", 4, 2)


@pytest.mark.parametrize("ctr,key,want", KATS)
def test_philox_kat(oracle_lib, ctr, key, want):
    from contracts_b200 import _lib
    from oracle import philox
    want = np.array(want, dtype=np.uint32)
    assert np.array_equal(philox.philox4x32_10(np.array(ctr), np.array(key)), want)
    assert np.array_equal(oracle_lib.philox4x32_10(ctr, key), want)
    lib = _lib.load()
    c = np.array(ctr, dtype=np.uint32)
    k = np.array(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib.ssd_philox4x32_10(c.ctypes.data_as(ctypes.c_void_p), k.ctypes.data_as(ctypes.c_void_p),
                          out.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(out, want)


def test_draw_addressing_matches_between_numpy_and_c(oracle_lib):
    from oracle import philox
    for (seed, env, ep, t, site, call, idx) in [(73907, 5, 0, 1, 3, 0, 7), (1, 4294967295, 3, 999, 6, 7, 18)]:
        got = philox.draws_u32(seed, env, ep, t, site, call, idx)[0]
        blk = oracle_lib.philox4x32_10((idx >> 2, site | (call << 8), t, ep), (seed, env))
        assert got == blk[idx & 3]
