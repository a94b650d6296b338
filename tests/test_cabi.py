"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every
symbol include/hypatia_b200.h declares; creating a context without a GPU fails loudly (there is
no CPU fallback); nothing under the product package imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from hypatia_b200 import capi
    return capi.load_library()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "hypatia_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hyp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    from hypatia_b200 import capi
    declared = _declared_symbols()
    assert len(declared) >= 30
    raw = ctypes.CDLL(capi.library_path())
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/hypatia_b200.h but not exported"
    assert set(declared) == set(capi.EXPORTED_SYMBOLS)


def test_version_and_timing_names(lib):
    assert lib.hyp_version() >= 100
    n = lib.hyp_timing_slots()
    names = [lib.hyp_timing_name(i).decode() for i in range(n)]
    assert "schur_syrk" in names and "potrf" in names and len(set(names)) == n


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hypatia_b200 import capi
    with pytest.raises(capi.HypatiaB200Error):
        capi.Context(0)
    from hypatia_b200.syssolver import QRCholDenseSystemSolver
    from hypatia_b200.host import instances as inst

    class S:
        pass
    s = S()
    s.model = inst.config("C3", 0.01).model
    with pytest.raises(capi.HypatiaB200Error):
        QRCholDenseSystemSolver().load(s)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hypatia.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "from oracle" not in text and "import oracle" not in text, f


def test_cone_codes_agree_between_header_host_mirror_and_julia_shim():
    """HYP_CONE_* / HYP_SSF_* of include/hypatia_b200.h, CONE_* / SSF_* of the Python host mirror and the cone_code
    methods of the Julia shim must name the same integers (the codes cross the ABI as plain ints)."""
    import re
    from pathlib import Path
    from hypatia_b200.host import models as M
    root = Path(__file__).resolve().parents[1]
    header = (root / "include" / "hypatia_b200.h").read_text()
    cones = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define HYP_CONE_(\w+)\s+(\d+)", header)}
    ssf = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define HYP_SSF_(\w+)\s+(\d+)", header)}
    assert len(cones) == len(M.CONE_NAMES) and sorted(cones.values()) == list(range(len(cones)))
    for name, code in cones.items():
        assert getattr(M, "CONE_" + name) == code, name
    for name, code in ssf.items():
        assert getattr(M, "SSF_" + name) == code, name
    # Julia shim: cone_code(::Cones.<Type>...) = Cint(<code>)
    julia = (root / "julia" / "HypatiaB200.jl").read_text()
    shim = {m.group(1).lower(): int(m.group(2))
            for m in re.finditer(r"cone_code\(::Cones\.(\w+)[^)]*\)\s*=\s*Cint\((\d+)\)", julia)}
    by_code = {v: k.lower().replace("_", "") for k, v in cones.items()}
    for jname, code in shim.items():
        assert by_code[code].startswith(jname[:8]), (jname, code, by_code[code])
    # the library's group table is sized for every code
    common = (root / "hypatia.jl_b200" / "csrc" / "common.cuh").read_text()
    assert int(re.search(r"#define HYP_NUM_CONE_TYPES\s+(\d+)", common).group(1)) == len(cones)
