"""The C-ABI boundary: every symbol include/spi_b200.h declares is exported by the built library and bound by
the Python loader; enum values in the header and their Python mirrors agree; argument errors come back as
negative return codes with a message (no compute is launched: no GPU needed)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from spi_active_b200 import _lib
from spi_active_b200 import go2_model as gm

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "spi_b200.h").read_text()


def _declared_functions():
    code = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(spi_b200_[a-z0-9_]+)\s*\(", code)))


def _enum_values():
    code = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    vals = {}
    for name, expr in re.findall(r"\b(SPI_[A-Z0-9_]+)\s*=\s*([^,}\n]+)", code):
        expr = expr.strip().replace("u", "")
        vals[name] = int(eval(expr))  # "1 << 3", "16", ...
    return vals


@pytest.fixture(scope="module")
def built_lib():
    _lib.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported_and_bound(built_lib):
    names = _declared_functions()
    assert len(names) >= 17
    assert sorted(_lib.SIGNATURES) == names, "python SIGNATURES must mirror include/spi_b200.h exactly"
    raw = C.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(raw, n), f"{n} not exported by {_lib.LIB_PATH.name}"


def test_version_and_header_constant(built_lib):
    assert built_lib.spi_b200_version() == int(re.search(r"#define SPI_B200_VERSION (\d+)", HEADER).group(1))


def test_enums_mirror_header():
    e = _enum_values()
    for k, v in gm.BLOB.items():
        assert e[f"SPI_BLOB_{k}"] == v, k
    for name, pid in (("mass", "MASS"), ("comx", "COMX"), ("comy", "COMY"), ("comz", "COMZ"), ("inertiax", "INERTIAX"),
                      ("inertiay", "INERTIAY"), ("inertiaz", "INERTIAZ"), ("inertiaxy", "INERTIAXY"),
                      ("inertiaxz", "INERTIAXZ"), ("inertiayz", "INERTIAYZ"), ("motor_model_hip_a", "MOTOR_HIP"),
                      ("motor_model_thigh_a", "MOTOR_THIGH"), ("motor_model_calf_a", "MOTOR_CALF"),
                      ("mass_scale", "MASS_SCALE")):
        assert gm.PARAM_IDS[name] == e[f"SPI_PARAM_{pid}"]
    assert gm.MOTOR_MODELS == {"none": e["SPI_MOTOR_NONE"], "act2tau_scalar": e["SPI_MOTOR_SCALAR"],
                               "act2tau_vec3": e["SPI_MOTOR_VEC3"], "act2tau_vec3_tanh": e["SPI_MOTOR_VEC3_TANH"]}
    assert (gm.FLAG_HIP_HALF, gm.FLAG_INERTIA_KEEP, gm.FLAG_STRICT_INERTIAY, gm.FLAG_TANH_BEFORE_CLIP) == (
        e["SPI_FLAG_HIP_HALF"], e["SPI_FLAG_INERTIA_KEEP"], e["SPI_FLAG_STRICT_INERTIAY"], e["SPI_FLAG_TANH_BEFORE_CLIP"])
    assert int(re.search(r"#define SPI_STATE_DIM (\d+)", HEADER).group(1)) == gm.STATE_DIM
    assert int(re.search(r"#define SPI_TARGET_DIM (\d+)", HEADER).group(1)) == gm.TARGET_DIM


def test_argument_errors_do_not_launch(built_lib):
    """NULL handle / bad blob -> negative rc + message; nothing here touches a device."""
    before = built_lib.spi_b200_launch_count()
    rc = built_lib.spi_b200_eval_candidates(None, None, 1, 0, None, None, None, None, None, None, 1, 1, 4, 0, 0, 0.0,
                                            None, None, None, None)
    assert rc < 0 and b"NULL" in built_lib.spi_b200_last_error()
    blob = gm.build_model_blob()
    bad = blob.copy(); bad[gm.BLOB["MAGIC"]] = 1.0
    h = C.c_void_p()
    rc = built_lib.spi_b200_model_create(bad.ctypes.data_as(C.POINTER(C.c_float)), int(bad.size), C.byref(h))
    assert rc < 0 and b"magic" in built_lib.spi_b200_last_error() and not h.value
    rc = built_lib.spi_b200_model_create(blob.ctypes.data_as(C.POINTER(C.c_float)), 10, C.byref(h))
    assert rc < 0 and b"short" in built_lib.spi_b200_last_error()
    assert built_lib.spi_b200_timing_enable(None, 1) < 0
    assert built_lib.spi_b200_launch_count() == before


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly, never fall back, when there is no CUDA device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from spi_active_b200.engine import RolloutEngine
    with pytest.raises(_lib.SpiB200Error):
        RolloutEngine()


def test_product_package_never_imports_oracle():
    pkg = ROOT / "spi_active_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        txt = p.read_text()
        assert "from oracle" not in txt and "import oracle" not in txt and "spi_oracle" not in txt, p
