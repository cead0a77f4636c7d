"""CPU tests: the C-ABI library loads and exports every symbol include/dj_b200.h declares
(no compute calls without a GPU), and the host-side size/plan logic is sane."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dj_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dj_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    import djb200

    lib = ctypes.CDLL(djb200.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in dj_b200.h but not exported"
    assert sorted(djb200.ABI_SYMBOLS) == declared


def test_version_and_workspace_queries_need_no_gpu():
    import djb200

    L = djb200.lib()
    assert L.dj_version() == 100
    small = L.dj_inner_join_workspace_bytes(1000, 1000)
    big = L.dj_inner_join_workspace_bytes(100_000_000, 100_000_000)
    assert 0 < small < big
    # two radix levels of scratch for 100M x 100M: at least 2 * 16 B * 200M rows
    assert big >= 2 * 16 * 200_000_000
    assert L.dj_distributed_inner_join_workspace_bytes(10**8, 10**8, 8, 1) > big


def test_missing_library_fails_loudly(monkeypatch):
    import djb200

    monkeypatch.setattr(djb200, "_lib", None)
    monkeypatch.setattr(djb200, "LIB_PATH", "/nonexistent/libdj_b200.so")
    try:
        djb200.lib()
    except djb200.DjError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")
