"""Achieved bandwidth of the bandwidth-bound kernels added for the secondary variants / edges (csrc/variants.cu), timed
with CUDA events on the launching stream (20 launches after 3 warm-ups, inputs > L2 or an L2 flush in between) against
the measured HBM copy bandwidth of MEASURED_PEAKS.json.  Not a pytest file:  python tests/probe_variants_bw.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import kernels as K  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32
dev = torch.device("cuda")
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    PEAK = 6650.0
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # 2x the 126 MB L2


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ms = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / reps


def report(name, ms, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "us": round(ms * 1e3, 1), "algorithmic_MB": round(nbytes / 1e6, 1),
                      "GBps": round(gbs, 0), "frac_of_measured_hbm": round(gbs / PEAK, 3)}), flush=True)


def main():
    # layer norm of the critic at D.Block.2's size (batch 128, 16x16x128 would be tiny): 128 x 32x32x128 bf16
    n, h, w, c = 128, 32, 32, 128
    x = torch.randn(n, h, w, c, device=dev).to(BF16)
    dy = torch.randn(n, h, w, c, device=dev).to(BF16)
    gam, bet = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    y, mr = K.layer_norm_fwd(x, gam, bet, 1e-12, "relu", BF16)
    report("layer_norm_fwd (partial + apply)", timed(lambda: K.layer_norm_fwd(x, gam, bet, 1e-12, "relu", BF16)),
           x.numel() * 2 * 3)                      # 2 reads + 1 write of 2 B
    dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    report("layer_norm_bwd (partial + apply + 2 colsum)",
           timed(lambda: K.layer_norm_bwd(x, dy, mr, gam, bet, "relu", BF16, dg, db)), x.numel() * 2 * 5)
    # depthwise 4x4 stride-2 conv at Pix2Pix encoder_3's size: 32 x 64x64x128 -> 32x32
    n, h, c, k = 32, 64, 128, 4
    x = torch.randn(n, h, h, c, device=dev).to(BF16)
    f = torch.randn(k, k, c, 1, device=dev) * 0.1
    pt, _, ho = K.same_pads(h, k, 2)
    y = K.depthwise_fwd(x, f, None, ho, ho, 2, pt, pt, BF16)
    gy = torch.randn(n, ho, ho, c, device=dev)
    report("depthwise_fwd 4x4 s2", timed(lambda: K.depthwise_fwd(x, f, None, ho, ho, 2, pt, pt, BF16)),
           x.numel() * 2 + y.numel() * 2)
    report("depthwise_bwd_input 4x4 s2", timed(lambda: K.depthwise_bwd_input(gy, f, h, h, 2, pt, pt, BF16)),
           gy.numel() * 4 + x.numel() * 2)
    df = torch.zeros_like(f)
    report("depthwise_bwd_filter 4x4 s2 (+ colsum)", timed(lambda: K.depthwise_bwd_filter(x, gy, df, k, k, 1, 2, pt, pt)),
           gy.numel() * 4 + x.numel() * 2)
    # fade-in blend at PGGAN 256x256: 16 x 256x256x16 fp32
    a = torch.randn(16, 256, 256, 16, device=dev)
    b = torch.randn(16, 256, 256, 16, device=dev)
    al = torch.full((1,), 0.3, device=dev)
    report("lerp_fwd", timed(lambda: K.lerp_fwd(a, b, al)), a.numel() * 4 * 3)
    report("lerp_bwd (2 outputs)", timed(lambda: K.lerp_bwd(a, al, F32, F32)), a.numel() * 4 * 3)
    # weight-norm transform of a 3x3 512->512 filter (9.4 MB): latency-sized
    wt = torch.randn(3, 3, 512, 512, device=dev)
    g = torch.ones(512, device=dev)
    we, norms = torch.empty_like(wt), torch.empty(512, device=dev)
    report("weight_transform_fwd 3x3x512x512", timed(lambda: K.weight_transform_fwd(wt, g, None, we, norms, 9 * 512, 512, 1)),
           wt.numel() * 4 * 3)
    # sample grid: 100 x 32x32x3 (generate_image) and 64 x 256x256x3
    for shape in ((100, 32, 32, 3), (64, 256, 256, 3)):
        s = torch.tanh(torch.randn(*shape, device=dev))
        report(f"sample_grid {shape}", timed(lambda: K.sample_grid(s, 10 if shape[0] == 100 else 8)), s.numel() * 5)
    # nearest half-resize of an RGB batch: 16 x 256x256x3 fp32
    img = torch.randn(16, 256, 256, 3, device=dev)
    report("subsample2d fwd (RGB)", timed(lambda: K.subsample2d(img, 2)), img.numel() * 4 // 4 * 2)


if __name__ == "__main__":
    main()
