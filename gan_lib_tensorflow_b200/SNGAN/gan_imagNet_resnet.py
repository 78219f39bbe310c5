"""SNGAN ImageNet 128x128 ResNet (conditional) on the B200 layer ops: the Generator / Discriminator of the reference
script SNGAN/gan_imagNet_resnet.py:241-334 (config 3 of BASELINE.json) with its constants (:40-68), its own Normalize
dispatch with 1000 classes (:88-112), hinge losses (:376-380, :497-500) and LR decay (:473-476).

Same building blocks and scheduling as SNGAN/gan_cifar_resnet.py (fused normalise+activation(+upsample) kernels,
1x1 shortcuts at the low resolution, residual sums in the GEMM epilogue, bf16 storage of every batch-norm input); the
conditioning of D is the reference's embed -> SN-Linear -> tile -> concat at 16x16 (:296-306), not a projection head."""
from __future__ import annotations

import torch

from .. import functional as F
from .. import kernels as K
from ..common import resnet_block as rb
from ..common.ops import conv2d as conv2d_ops
from ..common.ops import embedding as embedding_ops
from ..common.ops import linear as linear_ops
from ..framework import get_store

BATCH_SIZE = 32
GEN_BS_MULTIPLE = 2
ITERS = 450000
DIM_G = 128
DIM_D = 128
NORMALIZATION_G = True
NORMALIZATION_D = False
OUTPUT_DIM = 49152
LR = 0.0002
DECAY = True
N_CRITIC = 5
CONDITIONAL = True
ACGAN = False
VOCAB_SIZE = 1000
EMBEDDING_DIM = 300

BF16 = torch.bfloat16


def _normalize_kind(name, labels):
    """gan_imagNet_resnet.py:88-112"""
    if not CONDITIONAL:
        labels = None
    if CONDITIONAL and ACGAN and ('D.' in name):
        labels = None
    if ('D.' in name) and NORMALIZATION_D:
        return 'ln'
    elif ('G.' in name) and NORMALIZATION_G:
        return 'cbn' if labels is not None else 'bn'
    return None


def _block(inputs, input_dim, output_dim, filter_size, name, labels=None, **kw):
    return rb.ResidualBlock(inputs, input_dim, output_dim, filter_size, name, labels=labels, n_labels=VOCAB_SIZE,
                            normalize_kind=lambda nm: _normalize_kind(nm, labels), **kw)


def Generator(n_samples_, labels, noise=None, reuse=False):
    """gan_imagNet_resnet.py:241-271. Returns Var [n, 49152] (NHWC-flattened 128x128 images in (-1, 1))."""
    store = get_store()
    with store.variable_scope("Generator", reuse=reuse):
        if noise is None:
            noise = torch.randn(n_samples_, 128, device=store.device)
        output = linear_ops.Linear(F.as_var(noise), 128, 4 * 4 * DIM_G * 8, 'G.Input', out_dtype=BF16)
        output = F.reshape(output, (-1, 4, 4, DIM_G * 8))
        dims = [(DIM_G * 8, DIM_G * 8), (DIM_G * 8, DIM_G * 4), (DIM_G * 4, DIM_G * 2), (DIM_G * 2, DIM_G),
                (DIM_G, DIM_G // 2)]
        for i, (din, dout) in enumerate(dims):
            output = _block(output, din, dout, 3, 'G.Block.%d' % (i + 1), resample='up', labels=labels, biases=True,
                            out_dtype=BF16, out_bn_stats=True)
        output, _ = rb._norm_act('G.OutputNorm', output, labels, _normalize_kind('G.OutputNorm', labels), 'relu',
                                 n_labels=VOCAB_SIZE)
        output = conv2d_ops.Conv2D(output, DIM_G // 2, 3, 3, 1, 'G.Output', he_init=False)
        output = F.activation(output, 'tanh')
        return F.reshape(output, (-1, OUTPUT_DIM))


def Discriminator(inputs, labels, update_collection=None, reuse=False):
    """gan_imagNet_resnet.py:274-334. Returns (output_wgan Var [n], None)."""
    store = get_store()
    kw = dict(spectral_normed=True, update_collection=update_collection, labels=labels, biases=True)
    with store.variable_scope("Discriminator", reuse=reuse):
        output = F.reshape(F.as_var(inputs), (-1, 128, 128, 3))
        output = rb.OptimizedResBlockDisc1(output, DIM_D=DIM_D // 2, spectral_normed=True,
                                           update_collection=update_collection, biases=True,
                                           name_prefix='D.Block.1')                           # 64 x 64 x 64
        output = _block(output, DIM_D // 2, DIM_D, 3, 'D.Block.2', resample='down', **kw)   # 32 x 32 x 128
        output = _block(output, DIM_D, DIM_D * 2, 3, 'D.Block.3', resample='down', **kw)    # 16 x 16 x 256
        embedding_y = embedding_ops.embed_y(labels, VOCAB_SIZE, EMBEDDING_DIM)
        embedding_y = linear_ops.Linear(embedding_y, EMBEDDING_DIM, DIM_D, 'D.Embedding_y', spectral_normed=True,
                                        update_collection=update_collection, biases=True)
        pre = F.concat_label_map(output, embedding_y, act='relu')                            # 16 x 16 x 384
        output = _block(None, DIM_D * 3, DIM_D * 4, 3, 'D.Block.4', resample='down', pre_activated=pre, **kw)
        output = _block(output, DIM_D * 4, DIM_D * 8, 3, 'D.Block.5', resample='down', **kw)
        output = _block(output, DIM_D * 8, DIM_D * 8, 3, 'D.Block.6', resample=None, **kw)
        output = F.act_mean_hw(output, 'relu')
        output_wgan = linear_ops.Linear(output, DIM_D * 8, 1, 'D.Output', spectral_normed=True,
                                        update_collection=update_collection)
        return F.reshape(output_wgan, (-1,)), None


def lr_decay(iteration: int) -> float:
    """gan_imagNet_resnet.py:473-476"""
    if not DECAY:
        return 1.0
    return 1.0 if iteration < 400000 else max(0.0, 1.0 - iteration / 450000.0)


def _trainer_class():
    from . import gan_cifar_resnet as cifar

    class Trainer(cifar.Trainer):
        """D / G training steps of gan_imagNet_resnet.py:336-526: the graph structure of the CIFAR script (two towers,
        real + fake concatenated through D, hinge losses, Adam(2e-4, 0, 0.9), N_CRITIC = 5) around this module's
        Generator / Discriminator; BATCH_SIZE = 32 per process (:40), 1000 classes, no CHW -> NHWC transpose (:275)."""
        generator = staticmethod(lambda *a, **k: Generator(*a, **k))
        discriminator = staticmethod(lambda *a, **k: Discriminator(*a, **k))
        output_dim = OUTPUT_DIM
        image_hw = 128 * 128
        n_classes = VOCAB_SIZE
        gen_bs_multiple = GEN_BS_MULTIPLE
        base_lr = LR
        lr_schedule = staticmethod(lambda it: lr_decay(it))

        def __init__(self, batch_size: int = BATCH_SIZE, **kw):
            super().__init__(batch_size=batch_size, **kw)

        def _preprocess_real(self, b):
            # the flat vector is already NHWC: dequantise element-wise (hw = 1 makes the kernel's transpose the identity)
            return K.preprocess_real(self.real_int, self.deq_noise, b * self.image_hw, 1).reshape(b, self.output_dim)
    return Trainer


Trainer = _trainer_class()


SAMPLE_LABELS = [3, 547, 671, 833, 147, 218, 283, 292, 493, 555, 668, 698, 780, 985, 977, 963]   # :549-555


def train(iters: int = ITERS, **kw):
    """The training loop of gan_imagNet_resnet.py:546-705 -- the CIFAR script's loop (gan_cifar_resnet.train) around
    this module's Trainer, with the ImageNet script's sample function: 25 fixed-noise samples of ONE class drawn
    uniformly from SAMPLE_LABELS at every call (tf.multinomial over equal logits, :557-563), written as
    samples_<iteration>_<label>.png (:573); the dev-set cost is commented out in this script (:688-694)."""
    import numpy as np

    from . import gan_cifar_resnet as cifar

    def fixed_labels_fn():
        label = int(SAMPLE_LABELS[np.random.randint(len(SAMPLE_LABELS))])
        return np.full(25, label, dtype='int32'), label

    kw.setdefault('batch_size', BATCH_SIZE)
    if kw.get('trainer') is None:
        kw['trainer'] = Trainer(batch_size=kw['batch_size'], seed=kw.get('seed', 0))
    kw.setdefault('dev_gen', False)
    return cifar.train(iters, n_fixed=25, fixed_labels_fn=fixed_labels_fn, **kw)
