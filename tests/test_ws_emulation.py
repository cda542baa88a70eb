"""The warp-specialised kernel's arithmetic (spi_active_b200/csrc/go2_ws.cuh, all __host__ __device__) executed
on the CPU role by role (tests/host_emulation/ws_emulate.cpp) against the fp64 oracle.  Same tolerances as the
GPU per-step parity test: the two only differ by fp32 rounding."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from spi_active_b200 import go2_model as gm

import synth

HERE = Path(__file__).resolve().parent
SRC = HERE / "host_emulation" / "ws_emulate.cpp"
LIB = HERE / "host_emulation" / "_build" / "libws_emulate.so"
HDR = HERE.parent / "spi_active_b200" / "csrc" / "go2_ws.cuh"


@pytest.fixture(scope="module")
def emu():
    LIB.parent.mkdir(exist_ok=True)
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, HDR.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", str(SRC),
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
