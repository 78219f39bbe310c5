"""Step-by-step probe of the sub-pixel UpsampleConv path (sync after every launch). Not a pytest file."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import framework  # noqa: E402
from gan_lib_tensorflow_b200 import kernels as K  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32


def step(name, fn):
    try:
        r = fn()
        torch.cuda.synchronize()
        print("ok  ", name, flush=True)
        return r
    except Exception as e:  # noqa: BLE001
        print("FAIL", name, str(e)[:200], flush=True)
        raise


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    h = w = 16
    c = 256
    framework.reset_default_graph("cuda")
    dev = "cuda"
    x = torch.randn(n, h, w, c, device=dev).to(BF16)
    wf = torch.randn(3, 3, c, c, device=dev) * 0.02
    we_t = torch.empty(16, c, c, dtype=BF16, device=dev)
    we_n = torch.empty(16, c, c, dtype=BF16, device=dev)
    wt = torch.empty(9, c, c, dtype=BF16, device=dev)
    step("pack", lambda: K.upconv_pack(wf, we_t, we_n, c, c))
    bias = torch.zeros(c, device=dev)
    y = step("upconv_fprop bf16", lambda: K.upconv_fprop(x, we_t, n, h, w, c, c, None, bias, None, BF16))
    step("upconv_fprop again", lambda: K.upconv_fprop(x, we_t, n, h, w, c, c, None, bias, None, BF16))
    mean, rstd = step("bn_stats", lambda: K.bn_stats(y, n, 4 * h * w, c, 1, 1e-5))
    a2 = step("norm_act quad", lambda: K.norm_act_fwd(y, n, 2 * h, 2 * w, c, mean, rstd, 1, None, None, None, "relu", 2, BF16))
    # plain 3x3 conv on the full-resolution tensor through the same pair-kernel instantiation
    wt.copy_(wf.reshape(9, c, c).transpose(1, 2).to(BF16))
    res = torch.randn(n, h, w, c, device=dev)
    step("conv2 pair (no residual)", lambda: K.conv_igemm(a2, wt, n, 2 * h, 2 * w, c, 2 * h, 2 * w, c, 3, 3, 1, 1, False, None, bias, None, None, F32))
    step("conv2 pair residual_up2", lambda: K.conv_igemm(a2, wt, n, 2 * h, 2 * w, c, 2 * h, 2 * w, c, 3, 3, 1, 1, False, None, bias, res, None, F32, residual_up2=True))
    dy = torch.randn(n, 2 * h, 2 * w, c, device=dev).to(BF16)
    step("upconv_dgrad", lambda: K.upconv_dgrad(dy, we_n, n, h, w, c, c, None, BF16))
    dw = torch.zeros(3, 3, c, c, device=dev)
    step("upconv_wgrad", lambda: K.upconv_wgrad(x, dy, dw, n, h, w, c, c, None, 0.0))
    step("norm_act_bwd quad", lambda: K.norm_act_bwd(y, dy, 0, n, 2 * h, 2 * w, c, mean, rstd, 1, None, None, None, "relu", 2, None, None, None, BF16))
    step("conv2 pair again", lambda: K.conv_igemm(a2, wt, n, 2 * h, 2 * w, c, 2 * h, 2 * w, c, 3, 3, 1, 1, False, None, bias, None, None, F32))
    print("all ok")


if __name__ == "__main__":
    main()
