"""Config 3 of BASELINE.json -- SNGAN ImageNet-128 ResNet (conditional, projection-free critic of gan_imagNet_resnet.py),
"batch 256 sharded across 8 x B200" = 32 per GPU -- under torchrun at 1 / 2 / 4 / 8 ranks: D+G pairs per second (weak
scaling, device-timed, max over ranks), same schedule as the headline (Trainer.pair_step, gradient all-reduce between
the captured halves).  Not a pytest file.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/probe_imagenet_scaling.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import framework  # noqa: E402
from gan_lib_tensorflow_b200.SNGAN import gan_imagNet_resnet as P  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    wire = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    framework.reset_default_graph("cuda")
    allreduce = (lambda g: dist.all_reduce(g)) if world > 1 else None
    tr = P.Trainer(batch_size=32, seed=0, world_size=world, grad_allreduce=allreduce, grad_wire=wire)
    rs = np.random.RandomState(100 + rank)
    tr.set_real_batch(rs.randint(0, 256, size=(32, 49152)).astype("int32"), rs.randint(0, 1000, size=32).astype("int32"))
    for _ in range(2):
        tr.sample_noise(); tr.d_step(1); tr.g_step(1)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        tr.capture()
    torch.cuda.current_stream().wait_stream(s)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(5):
        tr.sample_noise(); tr.pair_step(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(steps):
        tr.sample_noise(); tr.pair_step(1)
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        m = float(ms.item())
        print(json.dumps({"config": "3 SNGAN ImageNet-128, 32 per GPU", "n_gpus": world, "ms_per_pair": m,
                          "pairs_per_s": world * 1e3 / m, "images_per_s": world * 32 * 1e3 / m, "gradient_wire": tr.grad_wire,
                          "tflops_per_gpu": 6145.0 / m, "d_loss": float(tr.d_loss.item()), "g_loss": float(tr.g_loss.item())}),
              flush=True)
    tr._graphs.clear()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
