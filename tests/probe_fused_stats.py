"""Times ganb_conv2d_igemm with and without the fused batch statistics on the dominant layers (not a pytest file)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import kernels as K  # noqa: E402

BF16 = torch.bfloat16


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


def main():
    for (n, h, cin, cout, k, out) in [(128, 32, 256, 256, 3, BF16), (64, 32, 256, 256, 3, BF16), (128, 16, 256, 256, 3, BF16),
                                      (128, 8, 1024, 256, 3, BF16), (128, 16, 256, 256, 1, BF16)]:
        x = torch.randn(n, h, h, cin, device="cuda").to(BF16)
        w = (torch.randn(k * k, cout, cin, device="cuda") * 0.02).to(BF16)
        b = torch.zeros(cout, device="cuda")
        pad = k // 2
        t0 = timeit(lambda: K.conv_igemm(x, w, n, h, h, cin, h, h, cout, k, k, pad, pad, False, None, b, None, None, out))
        t1 = timeit(lambda: K.conv_igemm_stats(x, w, n, h, h, cin, h, h, cout, k, k, pad, pad, False, None, b, None, None,
                                               out, 2))
        t2 = timeit(lambda: K.bn_stats(x, n, h * h, cin, 2, 1e-5))
        print(f"n{n} {h}x{h} {cin}->{cout} k{k}: plain {t0:7.1f} us  +stats {t1:7.1f} us  (separate bn_stats {t2:6.1f} us)"
              f"  dbg={os.environ.get('GANB_STATS_DBG', '0')}")


if __name__ == "__main__":
    main()
