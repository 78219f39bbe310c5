"""ctypes binding of libganb200.so (declared in include/ganb200.h).

There is no CPU fallback: if the library is missing, loading raises and every layer op fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GANB200_LIB") or os.path.join(HERE, "libganb200.so")   # override: kernel A/B experiments
HEADER_PATH = os.path.join(HERE, "..", "include", "ganb200.h")

OK = 0
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
F32, BF16 = 0, 1

_ACT_CODES = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "lrelu": ACT_LRELU, "tanh": ACT_TANH}


class GanbError(RuntimeError):
    """Raised when a libganb200 entry point returns a negative status."""


_lib = None


def header_symbols() -> list[str]:
    """Every function name declared in include/ganb200.h."""
    with open(HEADER_PATH) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ganb_[a-z0-9_]+)\s*\(", text)))


def lib() -> ctypes.CDLL:
    """Loads libganb200.so (once). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GanbError(
                f"{LIB_PATH} is missing: build it with `python -m gan_lib_tensorflow_b200.build` "
                "(there is no CPU or PyTorch fallback for the layer ops)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.ganb_last_error.restype = c_char_p
        _lib.ganb_conv2d_wgrad_workspace.restype = c_int64
        _lib.ganb_upconv_wgrad_workspace.restype = c_int64
    return _lib


_TRACE = os.environ.get("GANB_TRACE") == "1"   # debugging aid: print every C-ABI call as it returns


def check(rc: int, what: str) -> None:
    if _TRACE:
        print(f"[ganb] {what} -> {rc}", flush=True)
    if rc != OK:
        msg = lib().ganb_last_error().decode(errors="replace")
        raise GanbError(f"{what} failed with status {rc}: {msg}")


def act_code(name) -> int:
    try:
        return _ACT_CODES[name]
    except KeyError:
        raise ValueError(f"unknown activation {name!r}") from None


def ptr(t) -> c_void_p:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr() -> c_void_p:
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)


__all__ = ["lib", "check", "ptr", "stream_ptr", "act_code", "GanbError", "header_symbols", "c_int", "c_float",
           "c_int64", "c_void_p"]
