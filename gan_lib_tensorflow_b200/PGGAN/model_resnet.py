"""PGGAN based on the ResNet architecture (PGGAN/model_resnet.py:14-70): the same two methods as model_nvidia.PGGAN
over common.resnet_block.Generator_PGGAN / Discriminator_PGGAN (common/resnet_block.py:192-349).  Selected by the
reference's --model flag (PGGAN/train.py:61-66)."""
from __future__ import annotations

from .. import functional as F
from ..common import resnet_block
from ..framework import get_store


class PGGAN(object):
    def __init__(self, args=None, block_count=None, trans=None, inputs_norm=None):
        """args: an argparse-like namespace with block_count / trans / inputs_norm (model_resnet.py:15-22); the
        keywords override it."""
        self.bc = block_count if block_count is not None else args.block_count
        self.trans = trans if trans is not None else args.trans
        self.inputs_norm = inputs_norm if inputs_norm is not None else args.inputs_norm

    def get_generator(self, z_var, alpha, training=True, reuse=False):
        """:24-39.  Returns Var [n, 4 * 2^bc, 4 * 2^bc, 3] in (-1, 1)."""
        store = get_store()
        with store.variable_scope('g_net', reuse=reuse):
            z_var_ = F.as_var(z_var)
            z_var_ = F.reshape(z_var_, (z_var_.shape[0], -1))
            return resnet_block.Generator_PGGAN(z_var_, self.bc, self.trans, alpha, self.inputs_norm,
                                                training=training)

    def get_discriminator(self, x_var, alpha, labels=None, update_collection=None, reuse=False):
        """:41-70.  Returns the logits Var [n]."""
        store = get_store()
        with store.variable_scope('d_net', reuse=reuse):
            return resnet_block.Discriminator_PGGAN(x_var, labels, self.bc, self.trans, alpha, self.inputs_norm,
                                                    update_collection=update_collection, reuse=reuse)
