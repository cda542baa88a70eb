"""Loader of the in-tree CUDA library libspi_b200.so (C-ABI of include/spi_b200.h).

There is no CPU fallback: if the library is missing or cannot be loaded, every product entry point
raises.  `build()` compiles it with nvcc for sm_100a (used by __graft_entry__.build()).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("SPI_B200_LIB", _PKG / "libspi_b200.so"))  # override: kernel experiments only
SOURCES = [_PKG / "csrc" / "spi_b200.cu"]
HEADERS = [_PKG / "csrc" / "aba_leg.cuh", _PKG / "csrc" / "go2_ws.cuh", _PKG / "csrc" / "rollout_ws.cuh",
           _PKG / "csrc" / "fim_tc.cuh", _PKG / "csrc" / "active_step.cuh", _PKG / "csrc" / "mlp_tc.cuh", _PKG / "csrc" / "tiled_layout.cuh", _PKG / "csrc" / "pdl.cuh",
           _PKG.parent / "include" / "spi_b200.h"]
# SPI_WS_FAST_SINCOS: joint sin/cos through MUFU after a 2-constant reduction to [-pi, pi]; measured deviation from
# the fp64 oracle stays at the fp32 noise floor of the oracle itself (profiles/README.md, tests/tools/dev_accuracy.py)
# SPI_WS_FAST_TANH: the motor model's tanh through one ex2 + one reciprocal (|error| <= 2e-7 absolute, i.e. 4e-6 N m on a
# torque of ~20 N m = its fp32 rounding); +0.6 % throughput, deviation from the fp64 oracle unchanged (profiles/README.md)
# SPI_WS_CALF_HARMONIC: the calf's share of the thigh's articulated inertia as a trigonometric polynomial of the calf angle, its
# coefficients read from shared memory, and the structural zeros of that form skipped in the thigh's projection (go2_ws.cuh):
# leg loop 843 -> 814 instructions, 38.8 -> 38.3 ms; models without that structure fall back to the generic kernel
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-DSPI_WS_FAST_SINCOS",
              "-DSPI_WS_FAST_TANH", "-DSPI_WS_CALF_HARMONIC",
              # flush-to-zero fp32 (FFMA.FTZ ...): nothing on this path comes near 1e-38, and the rollout kernel is 1 % faster
              # (38.2 -> 37.9 ms; no change of the deviation from the fp64 oracle)
              "-ftz=true",
              "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "177"]

_lib = None


class SpiB200Error(RuntimeError):
    pass


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise SpiB200Error("nvcc not found; cannot build libspi_b200.so")


def _flags_stamp() -> Path:
    return LIB_PATH.with_suffix(".so.flags")


def needs_build() -> bool:
    """Rebuild when a source is newer than the library or when it was built with other compiler flags."""
    if not LIB_PATH.exists():
        return True
    newest = max(p.stat().st_mtime for p in SOURCES + HEADERS)
    if LIB_PATH.stat().st_mtime < newest:
        return True
    stamp = _flags_stamp()
    return not stamp.exists() or stamp.read_text() != " ".join(NVCC_FLAGS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), *map(str, SOURCES)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose:
        print(" ".join(cmd)); print(res.stdout)
    if res.returncode != 0:
        raise SpiB200Error(f"nvcc failed:\n{res.stdout}")
    _flags_stamp().write_text(" ".join(NVCC_FLAGS))
    return LIB_PATH


_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int)
_U8 = C.POINTER(C.c_ubyte)
_V = C.c_void_p

# name -> (restype, argtypes); every symbol include/spi_b200.h declares
SIGNATURES = {
    "spi_b200_version": (C.c_int, []),
    "spi_b200_last_error": (C.c_char_p, []),
    "spi_b200_launch_count": (C.c_longlong, []),
    "spi_b200_model_create": (C.c_int, [_F, C.c_int, C.POINTER(_V)]),
    "spi_b200_model_destroy": (C.c_int, [_V]),
    "spi_b200_eval_candidates": (C.c_int, [_V, _V, C.c_int, C.c_int, _I, _V, _V, _V, _V, _V, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_uint, C.c_float, _V, _V, _V, _V]),
    "spi_b200_eval_candidates_host": (C.c_int, [_V, _F, C.c_int, C.c_int, _I, _F, _F, _F, _F, _U8, C.c_int,
                                                C.c_int, C.c_int, C.c_int, C.c_uint, C.c_float, _F, _I, _V]),
    "spi_b200_rollout_states": (C.c_int, [_V, _V, C.c_int, C.c_int, _I, _V, _V, _V, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_uint, _V, _V]),
    "spi_b200_env_step": (C.c_int, [_V, _V, C.c_int, _I, _V, _V, _V, _V, C.c_int, C.c_int, C.c_int, C.c_uint, _V]),
    "spi_b200_sim_step": (C.c_int, [_V, _V, C.c_int, _I, C.c_uint, _V, _V, C.c_int, C.c_int, _V, _V]),
    "spi_b200_sim_step_ext": (C.c_int, [_V, _V, C.c_int, _I, C.c_uint, _V, _V, _V, C.c_int, C.c_int, _V, _V]),
    "spi_b200_body_states": (C.c_int, [_V, _V, C.c_int, _V, _V]),
    "spi_b200_compute_torques": (C.c_int, [_V, _V, _V, _V, _V, _V, C.c_int, C.c_int, C.c_uint, _V, _V]),
    "spi_b200_fim_reward": (C.c_int, [_V, _V, C.c_int, C.c_int, C.c_float, C.c_int, _V, _V, _V]),
    "spi_b200_active_post_step": (C.c_int, [_V, _V, _V, _V, _V, C.c_int,                       # model .. main_commands, T
                                            _V, _V, _V, _V, _V, _V, _V, _V, C.c_int, C.c_int,  # commands .. obs_lo, obs_stride, ring_slots
                                            _V, _V, _V, _V, _V, _V, C.c_float,                 # hist_index, fim_hist, fim_live, dead_steps, fim_jtj, fim_trace, fim_delta
                                            _V, C.c_int, _V, _V, C.c_int, C.c_int,             # schedule, schedule_rows, counter, ctrl, M, P1
                                            C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _F, _V]),
    "spi_b200_policy_create": (C.c_int, [_I, C.POINTER(_F), C.POINTER(_F), C.POINTER(_V)]),
    "spi_b200_policy_destroy": (C.c_int, [_V]),
    "spi_b200_policy_input_layout": (C.c_int, [_V, C.c_int, _I, _I]),
    "spi_b200_policy_split_input": (C.c_int, [_V, _V, C.c_int, _V, _V, _V]),
    "spi_b200_policy_unsplit_input": (C.c_int, [_V, _V, _V, C.c_int, _V, _V]),
    "spi_b200_policy_forward": (C.c_int, [_V, _V, _V, C.c_int, _V, _V]),
    "spi_b200_policy_enable_ring": (C.c_int, [_V, _I, C.c_int]),
    "spi_b200_policy_forward_ring": (C.c_int, [_V, _V, _V, C.c_int, _V, _V, _V]),
    "spi_b200_fim_contract": (C.c_int, [_V, _V, _V, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _V, _V, _V]),
    "spi_b200_cem_refit": (C.c_int, [_V, _V, _V, C.c_int, C.c_int, C.c_int, C.c_float, _V, _V, _V, _V, _V]),
    "spi_b200_cem_sample": (C.c_int, [_V, _V, _V, _V, _V, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_int,
                                      _V, _V]),
    "spi_b200_weighted_cost": (C.c_int, [_V, _V, C.c_int, C.c_float, C.c_float, C.c_float, _V, _V]),
    "spi_b200_fp32_peak": (C.c_int, [C.c_int, _F, _F, _V]),
    "spi_b200_model_set_kernel": (C.c_int, [_V, C.c_int]),
    "spi_b200_timing_enable": (C.c_int, [_V, C.c_int]),
    "spi_b200_timing_read": (C.c_int, [_V, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]),
}


def lib():
    """The loaded library; raises SpiB200Error (never falls back) if it is not there."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise SpiB200Error(
                f"{LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the rollout engine)")
        try:
            handle = C.CDLL(str(LIB_PATH))
        except OSError as e:  # pragma: no cover
            raise SpiB200Error(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().spi_b200_last_error()
        raise SpiB200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
