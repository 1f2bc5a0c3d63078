"""The C-ABI library builds, loads without a GPU and exports everything include/eitb200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from eitsynthai_b200 import build, cabi
    build.build()
    return cabi.load()


def _declared():
    with open(os.path.join(ROOT, "include", "eitb200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(eitb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from eitsynthai_b200 import cabi
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in eitb200.h but not exported"
        assert n in cabi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(cabi.SIGNATURES) == names


def test_argument_counts_match_header(lib):
    from eitsynthai_b200 import cabi
    with open(os.path.join(ROOT, "include", "eitb200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    for name, (_, args) in cabi.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_status_strings_and_version(lib):
    from eitsynthai_b200 import cabi
    assert lib.eitb_version() >= 100
    assert cabi.strerror(0) == "ok"
    assert "argument" in cabi.strerror(cabi.ERR_BAD_ARG)
    assert "workspace" in cabi.strerror(cabi.ERR_WORKSPACE)


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    # argument validation happens before any CUDA call, so it is checkable on a CPU-only box
    from eitsynthai_b200 import cabi
    assert lib.eitb_hu_window_nchw(0, 1, 512, 512, -160, 240, 1, 0, 0, 0, 0, 0, 0) == cabi.ERR_BAD_ARG
    assert lib.eitb_nms(0, 0, 1, 4, 32, 5376, 0.3, 0.7, 300, 7680.0, 0, 0, 0, 0, 0, 0) == cabi.ERR_BAD_ARG
    assert lib.eitb_tri_label(0, 0, 0, -1, 0, 0, 0, 0, 0, 4, 0, 0, 0, 0) == cabi.ERR_BAD_ARG
    assert lib.eitb_mask_decode(0, 0, 300, 0, 0, 0, 1, 32, 128, 128, 512, 512, 0, 0, 0, 0, 0, 0, 0) == cabi.ERR_BAD_ARG
    with pytest.raises(cabi.EitbError):
        cabi.call("eitb_minmax_u8", 0, 10, 0, 0, 0)


def test_ops_refuse_cpu_tensors():
    import torch
    from eitsynthai_b200 import ops
    with pytest.raises(ValueError):
        ops.hu_window(torch.zeros((1, 8, 8), dtype=torch.int16))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: without libeitb200.so the binding raises instead of computing elsewhere."""
    from eitsynthai_b200 import cabi
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", str(tmp_path / "libeitb200.so"))
    with pytest.raises(cabi.EitbLibraryError):
        cabi.load()
    with pytest.raises(cabi.EitbLibraryError):
        cabi.call("eitb_version")


def test_product_package_never_imports_the_oracle():
    import ast
    pkg = os.path.join(ROOT, "eitsynthai_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                tree = ast.parse(open(os.path.join(dirpath, f)).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom) and node.module:
                        names = [node.module]
                    assert not any(n == "oracle" or n.startswith("oracle.") for n in names), (f, names)
