"""2-GPU check of the cross-GPU batch-norm statistic reduction (torchrun --nproc-per-node 2): with bn_sync the two ranks
(batch 32 each, every statistic tower split over the ranks) must reproduce ONE GPU at batch 64 -- same losses, same
gradients up to summation order.  Rank 0 also runs the single-GPU reference.  Not a pytest file; run under
gpurun --gpus 2."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import framework  # noqa: E402
from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def split_towers(t, rank, world, towers=2):
    """rows of rank `rank`: its share of every statistic tower (tower = contiguous chunk of the global batch)"""
    per_tower = t.shape[0] // towers
    share = per_tower // world
    return torch.cat([t[k * per_tower + rank * share: k * per_tower + (rank + 1) * share] for k in range(towers)])


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    B = 64
    rs = np.random.RandomState(0)
    data = torch.from_numpy(rs.randint(0, 256, size=(B, 3072)).astype("int32"))
    labels = torch.from_numpy(rs.randint(0, 10, size=B).astype("int32"))
    z_d = torch.from_numpy(rs.standard_normal((B, 128)).astype("float32"))
    deq = torch.from_numpy(rs.uniform(0, 1 / 128, size=(B, 3072)).astype("float32"))
    z_g = torch.from_numpy(rs.standard_normal((2 * B, 128)).astype("float32"))
    fl = torch.from_numpy(rs.randint(0, 10, size=2 * B).astype("int32"))

    def run(batch, sync, pick):
        store = framework.reset_default_graph("cuda", u_seed=2)
        allreduce = (lambda g: dist.all_reduce(g)) if sync else None
        tr = P.Trainer(batch_size=batch, seed=0, world_size=world if sync else 1, grad_allreduce=allreduce, bn_sync=sync)
        tr.set_real_batch(pick(data).numpy(), pick(labels).numpy())
        tr.z_d.copy_(pick(z_d)); tr.deq_noise.copy_(pick(deq)); tr.z_g.copy_(pick(z_g)); tr.fake_labels.copy_(pick(fl))
        tr.disc_opt.set_lr(0.0); tr.gen_opt.set_lr(0.0)
        tr._d_compute()
        if sync:
            dist.all_reduce(store.flat["Discriminator"].grads)
            store.flat["Discriminator"].grads.div_(world)
        dg = store.flat["Discriminator"].grads.clone()
        tr._g_compute()
        if sync:
            dist.all_reduce(store.flat["Generator"].grads)
            store.flat["Generator"].grads.div_(world)
        gg = store.flat["Generator"].grads.clone()
        losses = torch.stack([tr.d_loss.clone(), tr.g_loss.clone()]).reshape(-1)
        if sync:
            dist.all_reduce(losses)
            losses /= world
        framework.set_store(None)
        return dg, gg, losses

    dg2, gg2, l2 = run(B // world, True, lambda t: split_towers(t, rank, world))
    dgp, ggp, lp = run(B // world, False, lambda t: split_towers(t, rank, world))     # per-rank statistics
    dist.all_reduce(dgp); dist.all_reduce(ggp); dist.all_reduce(lp)
    dgp /= world; ggp /= world; lp /= world
    if rank == 0:
        dg1, gg1, l1 = run(B, False, lambda t: t)
        print("losses 1 GPU      :", l1.tolist())
        print("losses 2 GPU sync :", l2.tolist())
        print("losses 2 GPU local:", lp.tolist())
        print(f"synced   vs 1 GPU: d_grads {rel(dg2, dg1):.3e}  g_grads {rel(gg2, gg1):.3e}  losses {rel(l2, l1):.3e}")
        print(f"per-rank vs 1 GPU: d_grads {rel(dgp, dg1):.3e}  g_grads {rel(ggp, gg1):.3e}  losses {rel(lp, l1):.3e}")
        ok = rel(l2, l1) < 1e-3 and rel(gg2, gg1) < 3e-2 and rel(dg2, dg1) < 3e-2
        print("BN_SYNC_OK" if ok else "BN_SYNC_MISMATCH")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
