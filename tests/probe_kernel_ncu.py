"""Launches the dominant kernels a few times for `ncu --set full` captures (fprop / dgrad / wgrad of the largest
layer: 3x3 256->256 at 128x32x32)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import kernels as K  # noqa: E402

n, h, w, cin, cout, k = 128, 32, 32, 256, 256, 3
dev = torch.device("cuda")
x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
wp = (torch.randn(k * k, cout, cin, device=dev) * 0.02).to(torch.bfloat16)
dy = torch.randn(n, h, w, cout, device=dev).to(torch.bfloat16)
bias = torch.zeros(cout, device=dev)
dw = torch.zeros(k, k, cin, cout, device=dev)
which = sys.argv[1] if len(sys.argv) > 1 else "fprop"
for _ in range(4):
    if which == "fprop":
        K.conv_igemm(x, wp, n, h, w, cin, h, w, cout, k, k, 1, 1, False, None, bias, None, None, torch.float32)
    elif which == "upconv":      # sub-pixel UpsampleConv 256->256 from 16x16 (G.Block.3.Conv1), fprop + dgrad + wgrad
        xl = x[:, :16, :16, :].contiguous()
        we = (torch.randn(16, cout, cin, device=dev) * 0.02).to(torch.bfloat16)
        yq = K.upconv_fprop(xl, we, n, 16, 16, cin, cout, None, bias, None, torch.bfloat16)
        K.upconv_dgrad(dy, we, n, 16, 16, cin, cout, None, torch.bfloat16)
        K.upconv_wgrad(xl, dy, dw, n, 16, 16, cin, cout, None, 0.0)
    elif which == "norm":        # conditional-BN forward / backward kernels at G.Block.3's size (HBM-bound)
        xb = x
        mean, rstd = K.bn_stats(xb, n, h * w, cin, 2, 1e-5)
        a = K.norm_act_fwd(xb, n, h, w, cin, mean, rstd, 2, None, None, None, "relu", False, torch.bfloat16)
        K.norm_act_bwd(xb, dy, 0, n, h, w, cin, mean, rstd, 2, None, None, None, "relu", False, None, None, None,
                       torch.bfloat16)
    elif which == "dgrad":
        K.conv_igemm(dy, wp, n, h, w, cout, h, w, cin, k, k, 1, 1, True, None, None, None, None, torch.float32)
    else:
        K.conv_wgrad(x, dy, dw, n, h, w, cin, h, w, cout, k, k, 1, 1, None, 0.0)
torch.cuda.synchronize()
print("done", which)
