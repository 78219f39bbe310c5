"""Config 2 of BASELINE.json on N GPUs: ACGAN CIFAR-10 (conditional BN generator, batch-normed two-headed critic, WGAN-GP
penalty), batch 64 per GPU, one rank per GPU with an NCCL all-reduce of the flat gradients between the captured compute
and update graphs.  Device-timed, max over ranks.  Not a pytest file:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/probe_acgan_multi.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import framework  # noqa: E402
from gan_lib_tensorflow_b200.ACGAN import train as AT  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    framework.reset_default_graph("cuda")
    allreduce = (lambda g: dist.all_reduce(g)) if world > 1 else None
    tr = AT.Trainer(batch_size=64, gradient_penalty=True, seed=0, world_size=world, grad_allreduce=allreduce)
    rs = np.random.RandomState(100 + rank)
    real = tr.preprocess(torch.from_numpy(rs.randint(0, 256, size=(64, 3072)).astype("int32")).cuda(), None)
    labels = torch.from_numpy(rs.randint(0, 10, size=64).astype("int32")).cuda()

    def pair():
        tr.d_step(real, labels, *tr._noise())
        tr.g_step(*tr._noise())

    for _ in range(2):
        pair()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        tr.capture()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(5):
        pair()
    steps = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        pair()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    if rank == 0:
        print(json.dumps({"config": "2 ACGAN CIFAR-10 (WGAN-GP penalty), batch 64 per GPU, CUDA graphs + NCCL gradient all-reduce",
                          "n_gpus": world, "ms_per_pair": ms, "pairs_per_s": world * 1000.0 / ms,
                          "images_per_s": world * 64 * 1000.0 / ms, "launches_d": tr.players.launches("d"),
                          "launches_g": tr.players.launches("g")}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
