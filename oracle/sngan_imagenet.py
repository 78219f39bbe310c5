"""TEST INFRASTRUCTURE -- CPU oracle (torch) restating the Generator / Discriminator of
SNGAN/gan_imagNet_resnet.py:88-112, 216-334 on top of oracle.ops / oracle.resnet_block.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.  Parity unpinned by the reference (no upstream
tests or golden vectors; TensorFlow 1.5 not installable)."""
from __future__ import annotations

import torch

from . import ops
from . import resnet_block as rb

DIM_G = 128
DIM_D = 128
NORMALIZATION_G = True
NORMALIZATION_D = False
CONDITIONAL = True
ACGAN = False
VOCAB_SIZE = 1000
EMBEDDING_DIM = 300
OUTPUT_DIM = 49152


def Normalize(g, name, inputs, labels=None):
    """gan_imagNet_resnet.py:88-112"""
    with g.variable_scope(name):
        if not CONDITIONAL:
            labels = None
        if CONDITIONAL and ACGAN and ("D." in name):
            labels = None
        if ("D." in name) and NORMALIZATION_D:
            return ops.layer_norm(g, name, [1, 2, 3], inputs)
        elif ("G." in name) and NORMALIZATION_G:
            if labels is not None:
                return ops.cond_batchnorm(g, name, [0, 1, 2], inputs, labels=labels, n_labels=1000)   # :104
            return ops.batch_norm(g, inputs, fused=True)
        return inputs


def _block(g, inputs, input_dim, output_dim, filter_size, name, **kw):
    norm = lambda nm, x, labels=None: Normalize(g, nm, x, labels)  # noqa: E731
    return rb.ResidualBlock(g, inputs, input_dim, output_dim, filter_size, name, normalize=norm, **kw)


def Generator(g, n_samples_, labels, noise, reuse=False):
    """gan_imagNet_resnet.py:241-271"""
    with g.variable_scope("Generator", reuse=reuse):
        output = ops.Linear(g, noise, 128, 4 * 4 * DIM_G * 8, "G.Input")
        output = output.reshape(-1, 4, 4, DIM_G * 8)
        dims = [(DIM_G * 8, DIM_G * 8), (DIM_G * 8, DIM_G * 4), (DIM_G * 4, DIM_G * 2), (DIM_G * 2, DIM_G),
                (DIM_G, DIM_G // 2)]
        for i, (din, dout) in enumerate(dims):
            output = _block(g, output, din, dout, 3, "G.Block.%d" % (i + 1), resample="up", labels=labels, biases=True)
        output = Normalize(g, "G.OutputNorm", output, labels)
        output = rb.nonlinearity(output)
        output = ops.Conv2D(g, output, DIM_G // 2, 3, 3, 1, "G.Output", he_init=False)
        return torch.tanh(output).reshape(-1, OUTPUT_DIM)


def Discriminator(g, inputs, labels, update_collection=None, reuse=False):
    """gan_imagNet_resnet.py:274-334"""
    kw = dict(spectral_normed=True, update_collection=update_collection, labels=labels, biases=True)
    with g.variable_scope("Discriminator", reuse=reuse):
        output = inputs.reshape(-1, 128, 128, 3)
        output = rb.OptimizedResBlockDisc1(g, output, DIM_D=DIM_D // 2, spectral_normed=True,
                                           update_collection=update_collection, biases=True, prefix="D.Block.1")
        output = _block(g, output, DIM_D // 2, DIM_D, 3, "D.Block.2", resample="down", **kw)
        output = _block(g, output, DIM_D, DIM_D * 2, 3, "D.Block.3", resample="down", **kw)
        e = ops.embed_y(g, labels, VOCAB_SIZE, EMBEDDING_DIM)
        e = ops.Linear(g, e, EMBEDDING_DIM, DIM_D, "D.Embedding_y", spectral_normed=True,
                       update_collection=update_collection, biases=True)
        e = e[:, None, None, :].expand(-1, output.shape[1], output.shape[2], -1)
        output = torch.cat([output, e], dim=3)
        output = _block(g, output, DIM_D * 3, DIM_D * 4, 3, "D.Block.4", resample="down", **kw)
        output = _block(g, output, DIM_D * 4, DIM_D * 8, 3, "D.Block.5", resample="down", **kw)
        output = _block(g, output, DIM_D * 8, DIM_D * 8, 3, "D.Block.6", resample=None, **kw)
        output = rb.nonlinearity(output).mean(dim=(1, 2))
        out = ops.Linear(g, output, DIM_D * 8, 1, "D.Output", spectral_normed=True, update_collection=update_collection)
        return out.reshape(-1), None


def preprocess_real(real_int, deq_noise, dtype):
    """gan_imagNet_resnet.py:354-356: int [B, 49152] -> float in [-1, 1) + U(0, 1/128); NO transpose -- the script
    reshapes the flat vector straight to [-1, 128, 128, 3] (:275)."""
    return 2 * ((real_int.to(dtype) / 256.0) - 0.5) + deq_noise


def lr_decay(iteration):
    """gan_imagNet_resnet.py:473-476"""
    return 1.0 if iteration < 400000 else max(0.0, 1.0 - iteration / 450000.0)


def _trainer_class():
    from . import sngan_cifar

    class SNGANImageNet(sngan_cifar.SNGANCifar):
        """Losses / train ops of gan_imagNet_resnet.py:336-500: the same two-tower structure as the CIFAR script
        (real + fake concatenated through D, hinge losses :376-380 / :497-500, Adam(2e-4, 0, 0.9) :521-526)."""
        G = staticmethod(lambda *a, **k: Generator(*a, **k))
        D = staticmethod(lambda *a, **k: Discriminator(*a, **k))
        preprocess = staticmethod(preprocess_real)
        lr = 0.0002
        decay = staticmethod(lr_decay)
    return SNGANImageNet


SNGANImageNet = _trainer_class()
