"""Worker of tests/test_gpu_multi.py (one process per GPU, launched with RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*):
peer-memory all-reduce (csrc/peer.cu) against NCCL, and the synced-BatchNorm trainer -- N ranks x 64/N with all-reduced
statistics against ONE GPU at batch 64, eagerly and as captured CUDA graphs.  Prints MULTI_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gan_lib_tensorflow_b200 import framework  # noqa: E402
from gan_lib_tensorflow_b200.peer import PeerComm  # noqa: E402
from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def split_towers(t, rank, world, towers=2):
    per_tower = t.shape[0] // towers
    share = per_tower // world
    return torch.cat([t[k * per_tower + rank * share: k * per_tower + (rank + 1) * share] for k in range(towers)])


def check_allreduce(peer, rank, world):
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    for count in (1, 7, 1024, 4096, 12032, 30000):
        x = torch.randn(count, device="cuda", generator=g)
        ref = x.clone()
        dist.all_reduce(ref)
        for rep in range(5):                       # epochs 1..5: both parities, repeated use of one site
            out = torch.empty_like(x)
            peer.allreduce(f"t{count}", x, out, scale=0.5)
            torch.cuda.synchronize()
            assert rel(out, 0.5 * ref) < 1e-6, (count, rep, rel(out, 0.5 * ref))
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        assert all(torch.equal(gathered[0], t) for t in gathered), "peer all-reduce must be bit-identical on every rank"
    # inside a captured graph, replayed
    x = torch.randn(2048, device="cuda", generator=g)
    out = torch.zeros_like(x)
    peer.allreduce("graph", x, out)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr):
            peer.allreduce("graph", x, out)
    torch.cuda.current_stream().wait_stream(s)
    for rep in range(6):
        x.normal_(generator=g)
        ref = x.clone()
        dist.all_reduce(ref)
        gr.replay()
        torch.cuda.synchronize()
        assert rel(out, ref) < 1e-6, rep
    # batch-norm moments: [mean | rstd] of equal shares -> global statistics
    c = 512
    xs = torch.randn(4096, c, device="cuda", generator=g) * (1 + rank) + rank
    mean, var = xs.mean(0), xs.var(0, unbiased=False)
    rstd = torch.rsqrt(var + 1e-5)
    peer.bn_moments("bn", mean, rstd, 1e-5)
    allx = [torch.empty_like(xs) for _ in range(world)]
    dist.all_gather(allx, xs)
    full = torch.cat(allx).double()
    assert float((mean.double() - full.mean(0)).abs().max()) < 1e-5
    assert rel(rstd, torch.rsqrt(full.var(0, unbiased=False) + 1e-5)) < 1e-5


def check_minibatch_std(peer, rank, world):
    """PGGAN's minibatch-stddev with the statistics over the global batch: `world` ranks x b/world == one rank x b,
    values and input gradients (PGGAN/model_nvidia.py:20-28)."""
    from gan_lib_tensorflow_b200 import functional as F

    b, h, w, c = 8, 4, 4, 512
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(b, h, w, c, device="cuda", generator=g)
    cot = torch.randn(b, h, w, c + 1, device="cuda", generator=g)
    share = b // world

    def run(xs, cots, sync):
        store = framework.reset_default_graph("cuda")
        store.bn_sync = sync
        xv = F.Var(xs.contiguous(), requires_grad=True)
        with store.gradient_tape() as tape:
            out = F.minibatch_std(xv)
            tape.backward(out, grad=cots.contiguous())
        torch.cuda.synchronize()
        framework.set_store(None)
        return out.data, xv.grad

    o_full, g_full = run(x, cot, None)
    sl = slice(rank * share, (rank + 1) * share)
    o_sync, g_sync = run(x[sl], cot[sl], peer)
    assert rel(o_sync, o_full[sl]) < 1e-6, rel(o_sync, o_full[sl])
    assert rel(g_sync, g_full[sl]) < 1e-5, rel(g_sync, g_full[sl])
    o_loc, _ = run(x[sl], cot[sl], None)                    # per-rank statistics give a different scalar
    assert float((o_loc[..., -1] - o_full[sl][..., -1]).abs().max()) > 1e-4


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    peer = PeerComm()
    check_allreduce(peer, rank, world)
    check_minibatch_std(peer, rank, world)

    B = 64
    rs = np.random.RandomState(0)
    data = torch.from_numpy(rs.randint(0, 256, size=(B, 3072)).astype("int32"))
    labels = torch.from_numpy(rs.randint(0, 10, size=B).astype("int32"))
    z_d = torch.from_numpy(rs.standard_normal((B, 128)).astype("float32"))
    deq = torch.from_numpy(rs.uniform(0, 1 / 128, size=(B, 3072)).astype("float32"))
    z_g = torch.from_numpy(rs.standard_normal((2 * B, 128)).astype("float32"))
    fl = torch.from_numpy(rs.randint(0, 10, size=2 * B).astype("int32"))
    allreduce = lambda g: dist.all_reduce(g)  # noqa: E731

    def make(batch, sync, pick):
        framework.reset_default_graph("cuda", u_seed=2)
        tr = P.Trainer(batch_size=batch, seed=0, world_size=world if sync else 1,
                       grad_allreduce=allreduce if sync else None, bn_sync=sync, peer=peer if sync else None)
        tr.set_real_batch(pick(data).numpy(), pick(labels).numpy())
        tr.z_d.copy_(pick(z_d)); tr.deq_noise.copy_(pick(deq)); tr.z_g.copy_(pick(z_g)); tr.fake_labels.copy_(pick(fl))
        return tr

    # ---- gradients of one critic step + one generator step: 2 x 32 synced == 1 x 64
    def grads(tr, sync):
        st = tr.store
        tr.disc_opt.set_lr(0.0); tr.gen_opt.set_lr(0.0)
        tr._d_compute()
        if sync:
            dist.all_reduce(st.flat["Discriminator"].grads)
            st.flat["Discriminator"].grads.div_(world)
        dg = st.flat["Discriminator"].grads.clone()
        tr._g_compute()
        if sync:
            dist.all_reduce(st.flat["Generator"].grads)
            st.flat["Generator"].grads.div_(world)
        gg = st.flat["Generator"].grads.clone()
        losses = torch.stack([tr.d_loss.clone(), tr.g_loss.clone()]).reshape(-1)
        if sync:
            dist.all_reduce(losses)
            losses /= world
        return dg, gg, losses

    tr = make(B // world, True, lambda t: split_towers(t, rank, world))
    assert tr.bn_sync and tr.bn_sync_in_graph
    dg2, gg2, l2 = grads(tr, True)
    framework.set_store(None)
    ok = True
    if rank == 0:
        dg1, gg1, l1 = grads(make(B, False, lambda t: t), False)
        framework.set_store(None)
        print(f"synced {world} x {B // world} vs 1 x {B}: d_grads {rel(dg2, dg1):.3e} g_grads {rel(gg2, gg1):.3e} "
              f"losses {rel(l2, l1):.3e}")
        ok = rel(l2, l1) < 1e-3 and rel(gg2, gg1) < 3e-2 and rel(dg2, dg1) < 3e-2

    # ---- training with the statistic exchanges inside captured graphs == the same steps run eagerly
    params = []
    for use_graphs in (False, True):
        tr = make(B // world, True, lambda t: split_towers(t, rank, world))
        for it in range(2):
            tr.d_step(1)
            tr.g_step(1)
        if use_graphs:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                tr.capture()
            torch.cuda.current_stream().wait_stream(s)
            assert "pair_fork" in tr._graphs
        for it in range(3):
            if use_graphs:
                tr.pair_step(1)
            else:
                tr.d_step(1)
                tr.g_step(1)
        torch.cuda.synchronize()
        st = tr.store
        params.append((st.flat["Generator"].params.clone(), st.flat["Discriminator"].params.clone()))
        tr._graphs.clear()
        framework.set_store(None)
    same = torch.equal(params[0][0], params[1][0]) and torch.equal(params[0][1], params[1][1])
    # replicated parameters stay identical on all ranks (bit-identical statistics, summed gradients)
    gp = [torch.empty_like(params[1][0]) for _ in range(world)]
    dist.all_gather(gp, params[1][0])
    replicated = all(torch.equal(gp[0], t) for t in gp)
    flag = torch.tensor([int(ok and same and replicated)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"graphs == eager: {same}; parameters identical on all ranks: {replicated}")
        print("MULTI_OK" if int(flag.item()) == 1 else "MULTI_MISMATCH")
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
