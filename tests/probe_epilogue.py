"""Launches the epilogue-bound convolution shapes of the headline step in isolation (for ncu --set full captures and
CUDA-event timing): K = 32 im2col routes on the RGB side, 1x1 shortcuts.  Not a pytest file."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import kernels as K  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32
SHAPES = [  # n, h, cin, cout, k, out dtype, tag
    (128, 32, 32, 128, 1, BF16, "D.Block.1.Conv1 fprop (im2col K=32)"),
    (128, 32, 32, 256, 1, BF16, "G.Output dgrad (im2col K=32)"),
    (128, 16, 256, 256, 1, F32, "G shortcut 1x1 256->256 @16x16"),
    (128, 16, 256, 128, 1, F32, "D shortcut 1x1 256->128 @16x16"),
    (128, 8, 128, 128, 3, BF16, "D.Block.3 3x3 128->128 @8x8"),
    (128, 32, 128, 128, 3, BF16, "D.Block.1.Conv2 3x3 128->128 @32x32"),
]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    for (n, h, cin, cout, k, out, tag) in SHAPES:
        x = torch.randn(n, h, h, cin, device="cuda").to(BF16)
        w = (torch.randn(k * k, cout, cin, device="cuda") * 0.05).to(BF16)
        b = torch.zeros(cout, device="cuda")
        pad = k // 2
        fn = lambda: K.conv_igemm(x, w, n, h, h, cin, h, h, cout, k, k, pad, pad, False, None, b, None, None, out)  # noqa: E731
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if reps > 1:     # launches back to back inside a CUDA graph: no host launch gaps between the short kernels
            g = torch.cuda.CUDAGraph()
            s_ = torch.cuda.Stream()
            s_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_):
                with torch.cuda.graph(g):
                    for _ in range(reps):
                        fn()
            torch.cuda.current_stream().wait_stream(s_)
            run = g.replay
        else:
            run = fn
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        out_mb = n * h * h * cout * (2 if out == BF16 else 4) / 1e6
        in_mb = n * h * h * cin * 2 / 1e6
        gf = 2.0 * n * h * h * cin * k * k * cout / 1e9
        print(f"{tag:42s} {us:7.1f} us  {gf / us / 1e3:7.1f} TFLOP/s  in+out {in_mb + out_mb:6.1f} MB -> "
              f"{(in_mb + out_mb) / us * 1e3 / 1e3:5.2f} TB/s")


if __name__ == "__main__":
    main()
