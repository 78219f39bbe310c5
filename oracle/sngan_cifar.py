"""Oracle restatement of SNGAN/gan_cifar_resnet.py: Generator, Discriminator, hinge losses, LR decay, TF Adam
and the two training ops, with every TF random tensor (noise, fake labels, dequantisation noise) injected."""
from __future__ import annotations

import functools

import numpy as np
import torch

from . import ops, resnet_block as rb, tfshim

BATCH_SIZE = 64          # gan_cifar_resnet.py:38
GEN_BS_MULTIPLE = 2      # :39
ITERS = 100000           # :40
DIM_G = 128              # :41
DIM_D = 128              # :42
NORMALIZATION_G = True   # :43
NORMALIZATION_D = False  # :44
OUTPUT_DIM = 3072        # :45
LR = 0.0002              # :46
N_CRITIC = 5             # :48
CONDITIONAL = True       # :51
ACGAN = False            # :52
VOCAB_SIZE = 10          # :59
EMBEDDING_DIM = 300      # :60
N_TOWERS = 2             # DEVICES always has two entries (:70-75)


def Normalize(g, name, inputs, labels=None):
    """gan_cifar_resnet.py:88-109 (the script's own dispatch: layer_norm in D only if NORMALIZATION_D)."""
    with g.variable_scope(name):
        if not CONDITIONAL:
            labels = None
        if CONDITIONAL and ACGAN and ("D." in name):
            labels = None
        if ("D." in name) and NORMALIZATION_D:
            return ops.layer_norm(g, name, [1, 2, 3], inputs)
        elif ("G." in name) and NORMALIZATION_G:
            if labels is not None:
                return ops.cond_batchnorm(g, name, [0, 1, 2], inputs, labels=labels, n_labels=10)
            return ops.batch_norm(g, inputs, fused=True)
        else:
            return inputs


def _block(g, inputs, input_dim, output_dim, filter_size, name, **kw):
    norm = lambda nm, x, labels=None: Normalize(g, nm, x, labels)  # noqa: E731
    return rb.ResidualBlock(g, inputs, input_dim, output_dim, filter_size, name, normalize=norm, **kw)


def Generator(g, n_samples_, labels, noise, reuse=False):
    """gan_cifar_resnet.py:237-263; `noise` [n,128] must be given (tf.random_normal is not reproducible)."""
    with g.variable_scope("Generator", reuse=reuse):
        output = ops.Linear(g, noise, 128, 4 * 4 * DIM_G * 8, "G.Input")
        output = output.reshape(-1, 4, 4, DIM_G * 8)
        output = _block(g, output, DIM_G * 8, DIM_G * 2, 3, "G.Block.1", resample="up", labels=labels, biases=True)
        output = _block(g, output, DIM_G * 2, DIM_G * 2, 3, "G.Block.2", resample="up", labels=labels, biases=True)
        output = _block(g, output, DIM_G * 2, DIM_G * 2, 3, "G.Block.3", resample="up", labels=labels, biases=True)
        output = Normalize(g, "G.OutputNorm", output, labels)
        output = rb.nonlinearity(output)
        output = ops.Conv2D(g, output, DIM_G * 2, 3, 3, 1, "G.Output", he_init=False)
        output = torch.tanh(output)
        return output.reshape(-1, OUTPUT_DIM)


def Discriminator(g, inputs, labels, update_collection=None, reuse=False):
    """gan_cifar_resnet.py:266-313"""
    with g.variable_scope("Discriminator", reuse=reuse):
        output = inputs.reshape(-1, 32, 32, 3)
        output = rb.OptimizedResBlockDisc1(g, output, DIM_D=DIM_D, spectral_normed=True,
                                           update_collection=update_collection, biases=True, prefix="D.Block.1")
        embedding_y = ops.embed_y(g, labels, VOCAB_SIZE, EMBEDDING_DIM)
        embedding_y = ops.Linear(g, embedding_y, EMBEDDING_DIM, DIM_D, "D.Embedding_y", spectral_normed=True,
                                 update_collection=update_collection, biases=True)
        embedding_y = embedding_y[:, None, None, :].expand(-1, output.shape[1], output.shape[2], -1)
        output = torch.cat([output, embedding_y], dim=3)
        output = _block(g, output, DIM_D * 2, DIM_D, 3, "D.Block.2", spectral_normed=True,
                        update_collection=update_collection, resample="down", labels=labels, biases=True)
        output = _block(g, output, DIM_D, DIM_D, 3, "D.Block.3", spectral_normed=True,
                        update_collection=update_collection, resample=None, labels=labels, biases=True)
        output = _block(g, output, DIM_D, DIM_D, 3, "D.Block.4", spectral_normed=True,
                        update_collection=update_collection, resample=None, labels=labels, biases=True)
        output = rb.nonlinearity(output)
        output = output.mean(dim=(1, 2))
        output_wgan = ops.Linear(g, output, DIM_D, 1, "D.Output", spectral_normed=True,
                                 update_collection=update_collection)
        return output_wgan.reshape(-1), None


def preprocess_real(real_int, deq_noise, dtype):
    """gan_cifar_resnet.py:334-337: int [B,3072] CHW -> float NHWC-flattened in [-1,1) + U(0,1/128)."""
    x = 2 * ((real_int.to(dtype) / 256.0) - 0.5)
    x = x + deq_noise
    return x.reshape(-1, 3, 32, 32).permute(0, 2, 3, 1).reshape(-1, OUTPUT_DIM)


def lr_decay(iteration):
    """gan_cifar_resnet.py:454-457"""
    return max(0.0, 1.0 - iteration / 100000.0) if iteration < 50000 else 0.5


class Adam:
    """tf.train.AdamOptimizer(lr, beta1, beta2, eps=1e-8): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; theta -= lr_t * m / (sqrt(v) + eps)   (SURVEY 8(c) item 6)."""

    def __init__(self, beta1=0.0, beta2=0.9, eps=1e-8):
        self.beta1, self.beta2, self.eps = beta1, beta2, eps
        self.t = 0
        self.m: dict[str, torch.Tensor] = {}
        self.v: dict[str, torch.Tensor] = {}

    def apply(self, named_params, grads, lr):
        self.t += 1
        lr_t = lr * np.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)
        with torch.no_grad():
            for (name, p), gr in zip(named_params, grads):
                if gr is None:
                    continue
                m = self.m.setdefault(name, torch.zeros_like(p))
                v = self.v.setdefault(name, torch.zeros_like(p))
                m.mul_(self.beta1).add_((1 - self.beta1) * gr)
                v.mul_(self.beta2).add_((1 - self.beta2) * gr * gr)
                p.sub_(lr_t * m / (torch.sqrt(v) + self.eps))


class SNGANCifar:
    """Graph + optimisers; build() mirrors the reference's graph-construction order so that the NumPy RNG
    stream is consumed identically (Appendix A of SURVEY.md)."""

    # model-specific pieces (overridden by oracle.sngan_imagenet.SNGANImageNet: same graph structure)
    G = staticmethod(lambda *a, **k: Generator(*a, **k))
    D = staticmethod(lambda *a, **k: Discriminator(*a, **k))
    preprocess = staticmethod(lambda real_int, deq_noise, dtype: preprocess_real(real_int, deq_noise, dtype))
    lr = LR
    decay = staticmethod(lambda it: lr_decay(it))

    def __init__(self, dtype=torch.float32, u_seed=2):
        self.g = tfshim.Graph(dtype=dtype, u_seed=u_seed)
        self.dtype = dtype
        self.gen_opt = Adam(0.0, 0.9)
        self.disc_opt = Adam(0.0, 0.9)
        self.built = False

    def build(self):
        """Creates all variables in reference order: G tower 0, G tower 1 (reuse: draws discarded), D."""
        z = torch.zeros(2, 128, dtype=self.dtype)
        lab = torch.zeros(2, dtype=torch.int64)
        self.g.draw_on_reuse = True
        with torch.no_grad():
            fake = self.G(self.g, 2, lab, z)
            self.G(self.g, 2, lab, z, reuse=True)
            self.D(self.g, fake, lab, update_collection=ops.NO_OPS)
            for _ in range(N_TOWERS):  # G-step towers (gan_cifar_resnet.py:464-482): G and D draw-and-discard
                self.D(self.g, self.G(self.g, 2, lab, z, reuse=True), lab, update_collection=ops.NO_OPS, reuse=True)
        self.g.draw_on_reuse = False
        self.built = True

    # ------------------------------------------------------------------------------------------ losses
    def disc_forward(self, real_int, real_labels, noises, deq_noise, update_collection=None):
        """gan_cifar_resnet.py:322-381 with one physical device (two towers of BATCH/2)."""
        g = self.g
        labels_splits = torch.chunk(real_labels, N_TOWERS)
        fake_splits = [self.G(g, real_int.shape[0] // N_TOWERS, labels_splits[i], noises[i], reuse=i > 0)
                       for i in range(N_TOWERS)]
        all_real = self.preprocess(real_int, deq_noise, self.dtype)
        real_splits = torch.chunk(all_real, N_TOWERS)
        real_and_fake = torch.cat([real_splits[0], real_splits[1], fake_splits[0], fake_splits[1]], dim=0)
        labels = torch.cat([labels_splits[0], labels_splits[1], labels_splits[0], labels_splits[1]], dim=0)
        disc_all, _ = self.D(g, real_and_fake, labels, update_collection=update_collection, reuse=True)
        n_real = real_int.shape[0]
        disc_real, disc_fake = disc_all[:n_real], disc_all[n_real:]
        cost = torch.relu(1.0 - disc_real).mean() + torch.relu(1.0 + disc_fake).mean()   # :376-378
        return cost

    def gen_forward(self, noises, fake_labels):
        """gan_cifar_resnet.py:462-498: two towers, each G(64) -> D(NO_OPS); cost = mean of -mean(D_fake)."""
        g = self.g
        costs = []
        for i in range(N_TOWERS):
            n_samples = noises[i].shape[0]
            fake = self.G(g, n_samples, fake_labels[i], noises[i], reuse=True)
            disc_fake, _ = self.D(g, fake, fake_labels[i], update_collection=ops.NO_OPS, reuse=True)
            costs.append(-disc_fake.mean())
        return sum(costs) / N_TOWERS

    # ------------------------------------------------------------------------------------------ train ops
    def disc_grads(self, real_int, real_labels, noises, deq_noise, update_collection=None):
        params = self.g.trainable_variables("Discriminator")
        cost = self.disc_forward(real_int, real_labels, noises, deq_noise, update_collection)
        grads = torch.autograd.grad(cost, [p for _, p in params], allow_unused=True)
        return cost.detach(), params, grads

    def gen_grads(self, noises, fake_labels):
        params = self.g.trainable_variables("Generator")
        cost = self.gen_forward(noises, fake_labels)
        grads = torch.autograd.grad(cost, [p for _, p in params], allow_unused=True)
        return cost.detach(), params, grads

    def disc_train_op(self, iteration, real_int, real_labels, noises, deq_noise):
        cost, params, grads = self.disc_grads(real_int, real_labels, noises, deq_noise, update_collection=None)
        self.disc_opt.apply(params, grads, self.lr * self.decay(iteration))
        return cost

    def gen_train_op(self, iteration, noises, fake_labels):
        cost, params, grads = self.gen_grads(noises, fake_labels)
        self.gen_opt.apply(params, grads, self.lr * self.decay(iteration))
        return cost


def synthetic_batch(seed=0, batch=BATCH_SIZE):
    """SURVEY 8(d) synthetic inputs: int32 [B,3072] uniform 0..255 and labels uniform 0..9."""
    rs = np.random.RandomState(seed)
    data = rs.randint(0, 256, size=(batch, OUTPUT_DIM)).astype("int32")
    labels = rs.randint(0, 10, size=(batch,)).astype("int32")
    return data, labels
