"""ctypes binding of include/ttsb200.h.  Plain pointers and sizes only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TTSB_LIB") or os.path.join(_HERE, "lib", "libttsb200.so")     # TTSB_LIB: A/B experiments only

OK, E_INVALID, E_CUDA, E_STATE, E_NOMEM, E_TOO_LONG, E_RANGE = 0, -1, -2, -3, -4, -5, -6
MEM_HOST, MEM_DEVICE = 0, 1

_p = C.c_void_p
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)

# name -> (restype, argtypes); mirrors include/ttsb200.h one to one (tests check the two stay in sync)
SIGNATURES = {
    "ttsb_abi_version": (C.c_int, []),
    "ttsb_last_error": (C.c_char_p, []),
    "ttsb_engine_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "ttsb_engine_destroy": (C.c_int, [_p]),
    "ttsb_engine_set_tensor": (C.c_int, [_p, C.c_char_p, _p, _i64p, C.c_int]),
    "ttsb_engine_finalize": (C.c_int, [_p]),
    "ttsb_engine_set_precision": (C.c_int, [_p, C.c_int]),
    "ttsb_engine_get_precision": (C.c_int, [_p]),
    "ttsb_engine_enable_taps": (C.c_int, [_p, C.c_int]),
    "ttsb_engine_launch_count": (C.c_int64, [_p]),
    "ttsb_engine_last_step_ms": (C.c_int, [_p, _f32p]),
    "ttsb_engine_workspace_peak": (C.c_int, [_p, _i64p]),
    "ttsb_engine_profile": (C.c_int, [_p, C.c_int]),
    "ttsb_engine_profile_read": (C.c_int, [_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), _i64p]),
    "ttsb_synthesize_begin": (C.c_int, [_p, _p, _p, C.c_int, _p, _p, C.c_int, _p, _p, _p, _p]),
    "ttsb_engine_voice_pack_set": (C.c_int, [_p, C.c_int, _p, C.c_int]),
    "ttsb_synthesize_begin_voices": (C.c_int, [_p, _p, _p, C.c_int, _p, _p, _p, C.c_int, _p, _p, _p, _p]),
    "ttsb_synthesize_finish": (C.c_int, [_p, _p, C.c_int, C.c_uint64, _p, _p, C.c_int]),
    "ttsb_vocoder": (C.c_int, [_p, _p, _p, _p, _p, C.c_int, C.c_int, _p, C.c_uint64, _p, C.c_int]),
    "ttsb_engine_read_tap": (C.c_int, [_p, C.c_char_p, _p, _i64p, _i64p]),
    "ttsb_op_istft": (C.c_int, [_p, _p, C.c_int, _p, C.c_int, _p, _f32p]),
    "ttsb_op_adain_snake": (C.c_int, [_p, _p, _p, C.c_int, _p, C.c_int, _p, _p, _f32p, _f32p]),
    "ttsb_op_length_regulate": (C.c_int, [_p, _p, _p, C.c_int, _p, C.c_int, _p, _p, _p, _f32p]),
    "ttsb_op_conv1d": (C.c_int, [_p, _p, _p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, C.c_int, C.c_int, C.c_int, _f32p]),
    "ttsb_op_bilstm": (C.c_int, [_p, _p, C.c_int, _p, _p, _p, _p, _p, C.c_int, _p, _f32p]),
}

_lib = None


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ttsb200 error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into lib/libttsb200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", _HERE, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libttsb200.so failed (see output above)")
    return LIB_PATH


def load():
    """dlopen the engine library.  Fails loudly: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C {_HERE}` (or __graft_entry__.build()). "
            "illufly-tts_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if lib.ttsb_abi_version() != 1:
        raise ImportError("libttsb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(code: int):
    if code != OK:
        msg = load().ttsb_last_error()
        raise EngineError(code, msg.decode("utf-8", "replace") if msg else "")
