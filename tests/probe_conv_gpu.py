"""GPU bring-up probe for the tensor-core convolution kernels (not a pytest file; run under gpurun).

Compares ganb_conv2d_igemm / ganb_conv2d_wgrad with torch fp32 convolutions on bf16-rounded inputs and
prints error statistics plus a coarse error map, so descriptor / swizzle mistakes can be localised from
one run.
"""
import ctypes
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import cabi  # noqa: E402

L = cabi.lib()
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def errmap(got, ref, name):
    diff = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    rel = diff.max().item() / scale
    ok = rel < 2e-2 and bool(torch.isfinite(got).all())
    print(f"  {name}: max_abs={diff.max().item():.4e} ref_max={scale:.4e} rel={rel:.3e} -> {'OK' if ok else 'FAIL'}")
    if not ok:
        flat = diff.reshape(-1, diff.shape[-1])
        rows = flat.shape[0]
        rb = max(1, rows // 16)
        cb = max(1, flat.shape[1] // 8)
        print("   error map (row blocks x col blocks, max abs):")
        for r0 in range(0, min(rows, 16 * rb), rb):
            line = " ".join(f"{flat[r0:r0 + rb, c0:c0 + cb].max().item():9.2e}" for c0 in range(0, flat.shape[1], cb))
            print(f"   r{r0:6d}: {line}")
        print("   got[0,:8] =", got.reshape(-1, got.shape[-1])[0, :8].tolist())
        print("   ref[0,:8] =", ref.reshape(-1, ref.shape[-1])[0, :8].tolist())
    return ok


def run_fprop(n, h, w, cin, cout, k, pad, flip=False, bias=True, res=False, alpha=None, act=0, bf16_out=False):
    g = torch.Generator(device="cpu").manual_seed(1234 + n + h + cin + cout + k)
    x = torch.randn(n, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(k, k, cin, cout, generator=g) / (k * (cin ** 0.5))).to(dev).to(torch.bfloat16)  # HWIO
    ho = h + 2 * pad - k + 1
    wo = w + 2 * pad - k + 1
    b = torch.randn(cout, generator=g).to(dev) if bias else None
    r = torch.randn(n, ho, wo, cout, generator=g).to(dev) if res else None
    a = torch.tensor([alpha], device=dev) if alpha is not None else None
    wp = wt.permute(0, 1, 3, 2).contiguous().reshape(k * k, cout, cin)  # [tap][cout][cin]
    y = torch.full((n, ho, wo, cout), float("nan"), device=dev, dtype=torch.bfloat16 if bf16_out else torch.float32)
    rc = L.ganb_conv2d_igemm(cabi.ptr(x), cabi.ptr(wp), cabi.ptr(y), n, h, w, cin, ho, wo, cout, k, k, 1, pad, pad,
                             0, cabi.ptr(a), cabi.ptr(b), cabi.ptr(r), 0, act, 1 if bf16_out else 0, stream())
    cabi.check(rc, "igemm")
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(3, 2, 0, 1), padding=pad).permute(0, 2, 3, 1)
    if alpha is not None:
        ref = ref * alpha
    if bias:
        ref = ref + b
    if res:
        ref = ref + r
    if act == 3:
        ref = torch.tanh(ref)
    return errmap(y.float(), ref, f"fprop n{n} {h}x{w} cin{cin} cout{cout} k{k} pad{pad}")


def run_dgrad(n, h, w, cin, cout, k, pad):
    g = torch.Generator(device="cpu").manual_seed(99 + n + h + cin + cout + k)
    ho = h + 2 * pad - k + 1
    wo = w + 2 * pad - k + 1
    dy = torch.randn(n, ho, wo, cout, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(k, k, cin, cout, generator=g) / (k * (cout ** 0.5))).to(dev).to(torch.bfloat16)  # HWIO
    dx = torch.full((n, h, w, cin), float("nan"), device=dev)
    wp = wt.reshape(k * k, cin, cout)  # HWIO viewed as [tap][cin][cout]
    rc = L.ganb_conv2d_igemm(cabi.ptr(dy), cabi.ptr(wp), cabi.ptr(dx), n, ho, wo, cout, h, w, cin, k, k, 1,
                             k - 1 - pad, k - 1 - pad, 1, None, None, None, 0, 0, 0, stream())
    cabi.check(rc, "dgrad")
    torch.cuda.synchronize()
    xx = torch.zeros(n, cin, h, w, device=dev, requires_grad=True)
    yy = F.conv2d(xx, wt.float().permute(3, 2, 0, 1), padding=pad)
    yy.backward(dy.float().permute(0, 3, 1, 2))
    ref = xx.grad.permute(0, 2, 3, 1)
    return errmap(dx, ref, f"dgrad n{n} {h}x{w} cin{cin} cout{cout} k{k} pad{pad}")


def run_wgrad(n, h, w, cin, cout, k, pad, time_it=False):
    g = torch.Generator(device="cpu").manual_seed(7 + n + h + cin + cout + k)
    ho = h + 2 * pad - k + 1
    wo = w + 2 * pad - k + 1
    x = torch.randn(n, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
    dy = (torch.randn(n, ho, wo, cout, generator=g) / ((n * ho * wo) ** 0.5)).to(dev).to(torch.bfloat16)
    dw = torch.full((k, k, cin, cout), float("nan"), device=dev)
    L.ganb_conv2d_wgrad_workspace.restype = ctypes.c_int64
    ws_bytes = L.ganb_conv2d_wgrad_workspace(n, h, w, cin, ho, wo, cout, k, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    rc = L.ganb_conv2d_wgrad(cabi.ptr(x), cabi.ptr(dy), cabi.ptr(dw), cabi.ptr(ws), n, h, w, cin, ho, wo, cout, k, k,
                             pad, pad, None, ctypes.c_float(0.0), stream())
    cabi.check(rc, "wgrad")
    torch.cuda.synchronize()
    wt = torch.zeros(cout, cin, k, k, device=dev, requires_grad=True)
    yy = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=pad)
    yy.backward(dy.float().permute(0, 3, 1, 2))
    ref = wt.grad.permute(2, 3, 1, 0)
    ok = errmap(dw, ref, f"wgrad n{n} {h}x{w} cin{cin} cout{cout} k{k} pad{pad} (ws {ws_bytes >> 20} MiB)")
    if time_it:
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        for _ in range(3):
            L.ganb_conv2d_wgrad(cabi.ptr(x), cabi.ptr(dy), cabi.ptr(dw), cabi.ptr(ws), n, h, w, cin, ho, wo, cout, k,
                                k, pad, pad, None, ctypes.c_float(0.0), stream())
        e0.record()
        for _ in range(10):
            L.ganb_conv2d_wgrad(cabi.ptr(x), cabi.ptr(dy), cabi.ptr(dw), cabi.ptr(ws), n, h, w, cin, ho, wo, cout, k,
                                k, pad, pad, None, ctypes.c_float(0.0), stream())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * n * ho * wo * cin * cout * k * k
        print(f"    wgrad time {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
    return ok


def time_fprop(n, h, w, cin, cout, k, pad):
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
    wp = torch.randn(k * k, cout, cin, device=dev).to(torch.bfloat16)
    y = torch.empty(n, h, w, cout, device=dev)
    args = (cabi.ptr(x), cabi.ptr(wp), cabi.ptr(y), n, h, w, cin, h, w, cout, k, k, 1, pad, pad, 0, None, None, None,
            0, 0, 0)
    for _ in range(3):
        L.ganb_conv2d_igemm(*args, stream())
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        L.ganb_conv2d_igemm(*args, stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * n * h * w * cin * cout * k * k
    print(f"  fprop n{n} {h}x{w} {cin}->{cout} k{k}: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s")


def main():
    print(torch.cuda.get_device_name(0))
    results = []
    t0 = time.time()
    # --- single tile, single k-iteration
    results.append(run_fprop(2, 8, 8, 64, 64, 1, 0, bias=False))
    results.append(run_fprop(2, 8, 8, 128, 64, 1, 0, bias=False))
    results.append(run_fprop(2, 8, 8, 64, 128, 1, 0))
    results.append(run_fprop(2, 8, 8, 64, 256, 1, 0))
    results.append(run_fprop(2, 8, 8, 64, 32, 1, 0))
    results.append(run_fprop(2, 8, 8, 64, 3, 1, 0))
    results.append(run_fprop(2, 8, 8, 72, 64, 1, 0))
    # --- 3x3 with halo / OOB fill
    results.append(run_fprop(2, 8, 8, 64, 64, 3, 1))
    results.append(run_fprop(3, 16, 16, 128, 128, 3, 1, res=True, alpha=0.37))
    results.append(run_fprop(5, 4, 4, 64, 64, 3, 1))
    results.append(run_fprop(4, 32, 32, 256, 256, 3, 1))
    results.append(run_fprop(4, 32, 32, 256, 3, 3, 1, act=3))
    results.append(run_fprop(2, 32, 32, 128, 128, 3, 1, bf16_out=True))
    results.append(run_fprop(64, 32, 32, 256, 256, 3, 1))
    # --- CTA-pair (cta_group::2) path: >= 74 pair tiles, Cout >= 128; odd tile count; residual / alpha / bf16 out
    results.append(run_fprop(75, 16, 16, 128, 128, 3, 1, res=True, alpha=0.37))
    results.append(run_fprop(151, 16, 8, 128, 256, 3, 1))
    results.append(run_fprop(40, 32, 32, 64, 384, 3, 1, bf16_out=True))
    results.append(run_dgrad(64, 32, 32, 256, 256, 3, 1))
    # --- dgrad through the same kernel
    results.append(run_dgrad(2, 8, 8, 64, 64, 3, 1))
    results.append(run_dgrad(4, 16, 16, 256, 128, 3, 1))
    results.append(run_dgrad(4, 16, 16, 256, 128, 1, 0))
    # --- wgrad
    results.append(run_wgrad(1, 8, 8, 64, 64, 1, 0))
    results.append(run_wgrad(2, 8, 8, 128, 64, 1, 0))
    results.append(run_wgrad(2, 8, 8, 128, 256, 3, 1))
    results.append(run_wgrad(8, 16, 16, 256, 128, 3, 1))
    results.append(run_wgrad(64, 32, 32, 256, 256, 3, 1, time_it=True))
    print(f"checks: {sum(results)}/{len(results)} OK in {time.time() - t0:.1f}s")
    # --- timing
    time_fprop(64, 32, 32, 256, 256, 3, 1)
    time_fprop(128, 32, 32, 256, 256, 3, 1)
    time_fprop(128, 32, 32, 256, 3, 3, 1)
    time_fprop(128, 16, 16, 256, 256, 3, 1)
    time_fprop(128, 32, 32, 128, 128, 3, 1)
    time_fprop(64, 16, 16, 256, 256, 3, 1)
    time_fprop(64, 8, 8, 1024, 256, 3, 1)
    time_fprop(128, 8, 8, 1024, 256, 3, 1)
    time_fprop(128, 8, 8, 256, 256, 3, 1)
    time_fprop(128, 8, 8, 128, 128, 3, 1)
    time_fprop(128, 8, 8, 256, 1024, 3, 1)
    time_fprop(128, 16, 16, 256, 128, 1, 0)
    return 0 if all(results) else 1


if __name__ == "__main__":
    sys.exit(main())
