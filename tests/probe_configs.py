"""Full-size timing of the other BASELINE.json configs on one B200 (synthetic data, eager steps, CUDA events):
  2  ACGAN CIFAR-10 (batch 64, with the gradient penalty)         3  SNGAN ImageNet-128 (batch 32 per GPU, CUDA graphs)
  4  Pix2Pix U-Net + PatchGAN at 256x256 (batch 32)                5  PGGAN 256x256 (block_count 6, fade-in, batch 16)
Prints ms per critic step / generator step and the implied reference iterations per second.  Not a pytest file."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import framework  # noqa: E402


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def imagenet():
    from gan_lib_tensorflow_b200.SNGAN import gan_imagNet_resnet as P
    framework.reset_default_graph("cuda")
    tr = P.Trainer(batch_size=32, seed=0)
    rs = np.random.RandomState(0)
    tr.set_real_batch(rs.randint(0, 256, size=(32, 49152)).astype("int32"), rs.randint(0, 1000, size=32).astype("int32"))
    for _ in range(2):
        tr.sample_noise(); tr.d_step(1); tr.g_step(1)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        tr.capture()
    torch.cuda.current_stream().wait_stream(s)
    d = timed(lambda: (tr.sample_noise(), tr.d_step(1)))
    g = timed(lambda: (tr.sample_noise(), tr.g_step(1)))
    return dict(config="3 SNGAN ImageNet-128, batch 32 per GPU, CUDA graphs", d_ms=d, g_ms=g, n_critic=5,
                gflop_pair=6145.0)


def acgan():
    from gan_lib_tensorflow_b200.ACGAN import train as AT
    framework.reset_default_graph("cuda")
    tr = AT.Trainer(batch_size=64, gradient_penalty=True, seed=0)
    rs = np.random.RandomState(0)
    real = tr.preprocess(torch.from_numpy(rs.randint(0, 256, size=(64, 3072)).astype("int32")).cuda(), None)
    labels = torch.from_numpy(rs.randint(0, 10, size=64).astype("int32")).cuda()
    d = timed(lambda: tr.d_step(real, labels, *tr._noise()))
    g = timed(lambda: tr.g_step(*tr._noise()))
    tr.capture()
    dg = timed(lambda: tr.d_step(real, labels, *tr._noise()), warm=3, reps=20)
    gg = timed(lambda: tr.g_step(*tr._noise()), warm=3, reps=20)
    return dict(config="2 ACGAN CIFAR-10, batch 64, with the WGAN-GP gradient penalty, CUDA graphs", d_ms=dg, g_ms=gg,
                n_critic=5, eager_d_ms=d, eager_g_ms=g, launches_d=tr.players.launches("d"),
                launches_g=tr.players.launches("g"), gflop_pair=650.0 + 764.0)


def pix2pix():
    from gan_lib_tensorflow_b200.Pix2Pix import train as PT
    framework.reset_default_graph("cuda")
    tr = PT.Trainer(ngf=64, ndf=64, size=256, seed=0)
    x = torch.rand(32, 256, 256, 3, device="cuda") * 2 - 1
    t = torch.rand(32, 256, 256, 3, device="cuda") * 2 - 1
    masks = lambda: [(torch.rand(32, s, s, 512, device="cuda") < 0.5).float() for s in (2, 4, 8)]  # noqa: E731
    d = timed(lambda: tr.d_step(x, t, masks()))
    g = timed(lambda: tr.g_step(x, t, masks()))
    tr.capture(x, t, masks())
    dg = timed(lambda: tr.d_step(x, t, masks()), warm=3, reps=10)
    gg = timed(lambda: tr.g_step(x, t, masks()), warm=3, reps=10)
    return dict(config="4 Pix2Pix unet_g + unet_d 256x256, batch 32, ngf = ndf = 64, CUDA graphs", d_ms=dg, g_ms=gg,
                n_critic=5, eager_d_ms=d, eager_g_ms=g, launches_d=tr.players.launches("d"),
                launches_g=tr.players.launches("g"),
                note="encoder_8's instance norm sees one pixel at 256x256 (reference quirk, DESIGN.md section 2)")


def pggan():
    from gan_lib_tensorflow_b200.PGGAN import train as PT
    framework.reset_default_graph("cuda")
    tr = PT.Trainer(6, True, inputs_norm=True, batch_size=16, seed=0)
    real = torch.rand(16, 256, 256, 3, device="cuda") * 2 - 1
    z = lambda: torch.randn(16, 512, device="cuda")  # noqa: E731
    d = timed(lambda: tr.d_step(real, z(), 0.5))
    g = timed(lambda: tr.g_step(z(), 0.5))
    tr.capture()
    dg = timed(lambda: tr.d_step(real, z(), 0.5), warm=3, reps=10)
    gg = timed(lambda: tr.g_step(z(), 0.5), warm=3, reps=10)
    return dict(config="5 PGGAN nvidia 256x256, block_count 6, trans, inputs_norm, batch 16, alpha 0.5, CUDA graphs "
                       "(alpha in a device scalar)", d_ms=dg, g_ms=gg, n_critic=5, eager_d_ms=d, eager_g_ms=g,
                launches_d=tr.players.launches("d"), launches_g=tr.players.launches("g"))


def main():
    which = sys.argv[1:] or ["imagenet", "acgan", "pix2pix", "pggan"]
    for name in which:
        t0 = time.time()
        try:
            r = globals()[name]()
            r["pair_ms"] = r["d_ms"] + r["g_ms"]
            r["reference_iterations_per_s"] = 1000.0 / (r["g_ms"] + r["n_critic"] * r["d_ms"])
            if "gflop_pair" in r:
                r["pair_tflops_algorithmic"] = r["gflop_pair"] / r["pair_ms"]
            r["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
        except Exception as e:  # noqa: BLE001
            r = dict(config=name, error=repr(e)[:300])
        r["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(r), flush=True)
        framework.set_store(None)
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
