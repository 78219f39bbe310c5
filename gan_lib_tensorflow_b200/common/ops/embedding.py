"""Label embedding with the reference's signature (common/ops/embedding.py:12-51)."""
from __future__ import annotations

import numpy as np
import torch

from ... import functional as F
from ...framework import get_store


def embed_y(inputs, vocab_size=1000, embedding_dim=300, word2vec_file=None,
            spectral_normed=False, update_collection=None, reuse=False):
    """inputs: int32 labels (batch,). Returns Var (batch, embedding_dim).  spectral_normed / update_collection /
    reuse are accepted and ignored exactly as in the reference."""
    store = get_store()
    with store.variable_scope("Embedding.Label"):
        if word2vec_file is None:
            embedding_map = store.get_variable(
                name='embedding_map', trainable=True,
                initializer=lambda _s: np.random.uniform(low=-0.08, high=0.08,
                                                         size=(vocab_size, embedding_dim)).astype('float32'))
        else:
            embedding_map = store.get_variable(name='embedding_map', trainable=False,
                                               initializer=np.asarray(word2vec_file, dtype='float32'))
        labels = inputs.data if isinstance(inputs, F.Var) else inputs
        labels = labels.to(device=store.device, dtype=torch.int32).reshape(-1).contiguous()
        return F.embedding(embedding_map, labels)
