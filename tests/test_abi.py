"""CPU: the C-ABI library loads and exports every symbol include/invpref_b200.h declares; struct layouts of
the ctypes mirror match; size queries (no GPU work) behave."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "invpref_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(invpref_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    from invpref_kdd_2022_b200 import _lib
    assert sorted(_lib.EXPORTS) == header_functions()


def test_library_exports_every_declared_symbol():
    from invpref_kdd_2022_b200 import _lib
    lib = _lib.load()
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.invpref_abi_version() == _lib.ABI_VERSION
    assert lib.invpref_strerror(0) == b"ok"
    assert b"dimension" in lib.invpref_strerror(-1)


def test_struct_layouts():
    from invpref_kdd_2022_b200 import _lib
    assert C.sizeof(_lib.Desc) == 40
    assert C.sizeof(_lib.Params) == 56
    assert C.sizeof(_lib.Adam) == 136
    assert C.sizeof(_lib.Batch) == 48
    assert C.sizeof(_lib.Hyper) == 128
    assert C.sizeof(_lib.Push) == 32
    assert C.sizeof(_lib.Dyn) == 16


def test_size_queries_and_argument_checks():
    from invpref_kdd_2022_b200 import _lib
    d = _lib.make_desc(10_000_000, 1_000_000, 4, 64, 0, 1, 0)
    ws, pl = _lib.workspace_bytes(d, 1 << 22), _lib.plan_bytes(d, 1 << 22)
    assert 100e6 < pl < 400e6 and ws > pl
    assert _lib.workspace_bytes(d, 0) > 0
    for bad in (_lib.make_desc(10, 10, 0, 8, 0, 0, 0), _lib.make_desc(10, 10, 9, 8, 0, 0, 0)):
        with pytest.raises(RuntimeError, match="environments"):
            _lib.workspace_bytes(bad, 16)
    for bad in (_lib.make_desc(10, 10, 2, 0, 0, 0, 0), _lib.make_desc(10, 10, 2, 257, 0, 0, 0),
                _lib.make_desc(10, 10, 2, 65, 0, 0, 0)):
        with pytest.raises(RuntimeError, match="dimension"):
            _lib.workspace_bytes(bad, 16)
    with pytest.raises(RuntimeError):
        _lib.workspace_bytes(_lib.make_desc(0, 10, 2, 8, 0, 0, 0), 16)
    with pytest.raises(RuntimeError):
        _lib.workspace_bytes(d, -1)


def test_no_cpu_fallback():
    import torch
    from invpref_kdd_2022_b200 import _lib
    with pytest.raises(RuntimeError, match="CUDA"):
        _lib.ptr(torch.zeros(4))
    from invpref_kdd_2022_b200.models import InvPrefExplicit
    m = InvPrefExplicit(5, 6, 2, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(3, dtype=torch.int64), torch.zeros(3, dtype=torch.int64), torch.zeros(3, dtype=torch.int64), 1.0)


def test_shipped_library_contains_no_cub_or_thrust_kernels():
    """The sort-segment plan is built by the library's own radix sort and prefix sums (csrc/sort.cuh): the device
    code of the shipped .so holds the psort:: kernels and no CUB / Thrust instantiation (round 1 called
    cub::DeviceRadixSort / DeviceScan there)."""
    from invpref_kdd_2022_b200 import _lib
    _lib.load()
    path = os.environ.get("INVPREF_LIB") or os.path.join(ROOT, "invpref_kdd_2022_b200", "libinvpref_b200.so")
    blob = open(path, "rb").read()
    for needle in (b"DeviceRadixSort", b"DeviceScan", b"thrust", b"cub17", b"3cub"):
        assert needle not in blob, needle
    for needle in (b"rs_scatter_kernel", b"rs_hist_kernel", b"rs_offsets_kernel", b"sc_scan_kernel", b"sc_mid_kernel"):
        assert needle in blob, needle
