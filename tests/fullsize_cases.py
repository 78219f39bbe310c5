"""Parity cases at the BASELINE.json shapes of configs 3-5 (SNGAN ImageNet-128 at full width, Pix2Pix at ngf = ndf = 64,
PGGAN model_nvidia at 256x256 with fade-in): shared by tests/golden/make_fullsize.py, which runs the CPU oracle once
(minutes per case) and freezes compact summaries under tests/golden/, and by tests/test_gpu_fullsize.py, which runs the
CUDA path at the same sizes and compares against those summaries.

A summary of a tensor = its L2 norm and a fixed pseudo-random sample of SAMPLE elements (indices from the tensor's size
and name); relative errors are evaluated on the sample."""
import zlib

import numpy as np

SAMPLE = 4096


def sample_index(name: str, size: int) -> np.ndarray:
    if size <= SAMPLE:
        return np.arange(size)
    rs = np.random.RandomState(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return np.sort(rs.choice(size, SAMPLE, replace=False))


def summarize(name: str, t) -> dict:
    a = np.asarray(t, dtype=np.float32).reshape(-1)
    return {"norm": np.float64(np.linalg.norm(a.astype(np.float64))), "sample": a[sample_index(name, a.size)].copy()}


def flatten(prefix: str, summary: dict, out: dict) -> None:
    out[prefix + "::norm"] = np.asarray(summary["norm"])
    out[prefix + "::sample"] = summary["sample"]


def rel(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


# ------------------------------------------------------------------------------------------------ inputs (seeded)
def pggan_inputs():
    """PGGAN model_nvidia at block_count 6 = 256x256 (config 5), fade-in on, batch 2."""
    rs = np.random.RandomState(601)
    return dict(bc=6, trans=True, alpha=0.3, n=2,
                z=rs.standard_normal((2, 512)).astype("float32"),
                x=rs.uniform(-1, 1, size=(2, 256, 256, 3)).astype("float32"),
                cot_g=rs.standard_normal((2, 256, 256, 3)).astype("float32"),
                cot_d=rs.standard_normal((2,)).astype("float32"))


def pix2pix_inputs():
    """Pix2Pix unet_g / unet_d at ngf = ndf = 64 (config 4 widths), one 512x512 image (every instance norm of the
    generator then sees at least 2x2 pixels, see tests/test_gpu_wide.py::test_pix2pix_unet_generator)."""
    rs = np.random.RandomState(602)
    size, ngf = 512, 64
    return dict(n=1, size=size, ngf=ngf, ndf=64,
                x=rs.uniform(-1, 1, size=(1, size, size, 3)).astype("float32"),
                tgt=rs.uniform(-1, 1, size=(1, size, size, 3)).astype("float32"),
                masks=[(rs.uniform(size=(1, s, s, ngf * 8)) < 0.5).astype("float32") for s in (4, 8, 16)],
                cot_g=rs.standard_normal((1, size, size, 3)).astype("float32"))


def imagenet_inputs():
    """SNGAN ImageNet-128 at full width (DIM_G = DIM_D = 128, config 3), batch 4: one critic step and one generator step."""
    rs = np.random.RandomState(603)
    b = 4
    return dict(batch=b, data=rs.randint(0, 256, size=(b, 49152)).astype("int32"),
                labels=rs.randint(0, 1000, size=b).astype("int32"),
                z_d=rs.standard_normal((b, 128)).astype("float32"),
                deq=rs.uniform(0, 1 / 128, size=(b, 49152)).astype("float32"),
                z_g=rs.standard_normal((2 * b, 128)).astype("float32"),
                fl=rs.randint(0, 1000, size=2 * b).astype("int32"))


CASES = ("pggan_g", "pggan_d", "pix2pix_g", "pix2pix_d", "imagenet_step")
