"""The warp-specialised kernel's arithmetic (spi_active_b200/csrc/go2_ws.cuh, all __host__ __device__) executed
on the CPU role by role (tests/host_emulation/ws_emulate.cpp) against the fp64 oracle.  Same tolerances as the
GPU per-step parity test: the two only differ by fp32 rounding."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from spi_active_b200 import go2_model as gm

import synth

HERE = Path(__file__).resolve().parent
SRC = HERE / "host_emulation" / "ws_emulate.cpp"
HDR = HERE.parent / "spi_active_b200" / "csrc" / "go2_ws.cuh"
# the emulation is compiled with the product's own -DSPI_WS_* switches (spi_active_b200/_lib.py), so that it runs the same
# formulation as the shipped kernel; SPI_WS_EMU_DEFS adds switches for experiments
from spi_active_b200 import _lib as _product  # noqa: E402
DEFS = [f for f in _product.NVCC_FLAGS if f.startswith("-DSPI_WS_")] + os.environ.get("SPI_WS_EMU_DEFS", "").split()
LIB = HERE / "host_emulation" / "_build" / ("libws_emulate" + "".join(d.replace("-D", "_") for d in DEFS) + ".so")


@pytest.fixture(scope="module")
def emu():
    LIB.parent.mkdir(exist_ok=True)
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, HDR.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", *DEFS, "-x", "c++", str(SRC),
                        "-o", str(LIB)], check=True)
    return C.CDLL(str(LIB))


def _run(emu, blob, params, ids, init, actions, gains, H, motor=0, flags=0):
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    params = np.ascontiguousarray(params, np.float32); ids = np.ascontiguousarray(ids, np.int32)
    init = np.ascontiguousarray(init, np.float32); actions = np.ascontiguousarray(actions, np.float32)
    gains = np.ascontiguousarray(gains, np.float32)
    out = np.zeros((H, 37), np.float32); ff = np.zeros((4, 3), np.float32)
    rc = emu.ws_emulate_rollout(fp(blob), fp(params), C.c_int(params.size), ids.ctypes.data_as(C.POINTER(C.c_int)),
                                C.c_uint(flags), C.c_int(motor), fp(init), fp(actions), fp(gains), C.c_int(H), C.c_int(4),
                                fp(out), fp(ff))
    assert rc == 0
    return out, ff


@pytest.mark.parametrize("config,motor,flags", [("jump", 0, 0), ("sine", 3, 0), ("walk", 3, gm.FLAG_HIP_HALF), ("stand", 2, 0),
                                                ("sine", 1, 0), ("jump", 3, gm.FLAG_TANH_BEFORE_CLIP)])
def test_emulated_roles_match_oracle(emu, oracle_lib, blob, config, motor, flags):
    S, ds = synth.dataset(config, 5)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    names = ["mass", "comx", "comy", "comz", "inertiax", "inertiay", "inertiaz", "motor_model_hip_a",
             "motor_model_thigh_a", "motor_model_calf_a"]
    ids = [gm.PARAM_IDS[n] for n in names]
    rng = np.random.default_rng(2)
    nominal = gm.default_param_vector()[ids]
    for s in range(0, S, max(1, S // 12)):
        p = nominal.copy()
        p[0] += rng.uniform(-2, 3); p[1:4] += rng.uniform(-0.03, 0.03, 3); p[4:7] *= rng.uniform(0.7, 1.4, 3)
        p[7:10] = rng.uniform(0.7, 1.3, 3) if motor in (1, 2) else rng.uniform(12, 30, 3)
        out, _ = _run(emu, blob, p, ids, init[s], act[s], gains[s], 5, motor, flags)
        ref = oracle_lib.rollout_states(blob, p[None].astype(np.float32), ids, init[s:s + 1], act[s:s + 1], gains[s:s + 1],
                                        motor_model=motor, flags=flags)[0, 0]
        np.testing.assert_allclose(out[:, :7], ref[:, :7], atol=2e-4, rtol=0)
        np.testing.assert_allclose(out[:, 13:25], ref[:, 13:25], atol=2e-4, rtol=0)
        np.testing.assert_allclose(out[:, 7:13], ref[:, 7:13], atol=5e-3, rtol=0)
        np.testing.assert_allclose(out[:, 25:37], ref[:, 25:37], atol=2e-2, rtol=0)


def test_fast_path_rejects_foreign_joint_geometry(emu, nominal_model):
    """A URDF whose joint origins break the Go2 sparsity pattern must be refused by the specialised path
    (the engine then uses the generic leg-per-lane kernel)."""
    import copy
    m = copy.deepcopy(nominal_model)
    m.joint_origin[1] = [0.01, 0.0955, 0.0]
    b = gm.build_model_blob(m)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    z = np.zeros(64, np.float32); out = np.zeros((1, 37), np.float32)
    rc = emu.ws_emulate_rollout(fp(b), fp(z), C.c_int(0), np.zeros(1, np.int32).ctypes.data_as(C.POINTER(C.c_int)),
                                C.c_uint(0), C.c_int(0), fp(z), fp(z), fp(z), C.c_int(1), C.c_int(4), fp(out), None)
    assert rc == -2


def test_harmonic_calf_inertia_holds_for_an_asymmetric_calf(emu, oracle_lib, nominal_model):
    """SPI_WS_CALF_HARMONIC keeps only the harmonics of the calf angle that can occur behind a y-axis joint with a
    z-only joint offset and skips the products with the entries that vanish there (H[y][y], M[x][y], M[y][z]).  Those zeros are
    structural — the coupling block of a rigid leaf is skew, so U_l has no component along the joint axis — not a symmetry of
    the Go2 calf: a calf with a lateral centre-of-mass offset and extra products of inertia must run through the same path
    and still match the oracle at fp32 rounding."""
    import copy
    m = copy.deepcopy(nominal_model)
    for leg in range(4):
        calf = m.leg_bodies[3 * leg + 2]
        inertia = list(calf.inertia); inertia[3] += 2e-4; inertia[5] += 2e-4
        m.leg_bodies[3 * leg + 2] = type(calf)(calf.mass, [calf.com[0], calf.com[1] + 0.03, calf.com[2]], inertia)
    blob = gm.build_model_blob(m)
    S, ds = synth.dataset("sine", 5)
    init, act, tgt, gains, mask, denom = synth.pack_numpy(ds)
    ids = np.array([0], np.int32); p = np.array([6.921], np.float32)
    for s in (0, 7, 19):
        out, _ = _run(emu, blob, p, ids, init[s], act[s], gains[s], 5)
        ref = oracle_lib.rollout_states(blob, p[None].astype(np.float32), ids, init[s:s + 1], act[s:s + 1], gains[s:s + 1])[0, 0]
        np.testing.assert_allclose(out[:, :7], ref[:, :7], atol=5e-6, rtol=0)
        np.testing.assert_allclose(out[:, 13:25], ref[:, 13:25], atol=5e-6, rtol=0)
