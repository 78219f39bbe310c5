"""The step either side of the hot path, with the reference's names (common/misc.py): the sample-grid writer
(save_images :215-244, fed by generate_image of SNGAN/gan_cifar_resnet.py:536-539), get_z (:247-259), get_loss
(:310-394) and optimistic_restore (:275-307).

Sample grids: the int conversion and the tiling run on the GPU (ganb_sample_grid, one pass, uint8 out: 4x fewer bytes
cross PCIe than the fp32 samples); the host keeps only what scipy.misc.imsave did to the assembled array -- the
min/max contrast stretch of scipy's bytescale -- and the PNG encoder (PIL)."""
from __future__ import annotations

import errno
import os

import numpy as np
import torch

from .. import functional as F
from .. import kernels as K
from ..framework import get_store


def mkdir_p(path):
    """common/misc.py:204-212"""
    try:
        os.makedirs(path)
    except OSError as exc:
        if exc.errno == errno.EEXIST and os.path.isdir(path):
            pass
        else:
            raise


def grid_shape(n_samples: int):
    """rows x columns of save_images (:221-226): the largest divisor of n not above sqrt(n) is the row count."""
    rows = int(np.sqrt(n_samples))
    while n_samples % rows != 0:
        rows -= 1
    return rows, int(n_samples / rows)


def _bytescale(img: np.ndarray) -> np.ndarray:
    """scipy.misc.bytescale as imsave applies it to a non-uint8 array: stretch [min, max] onto [0, 255]."""
    data = img.astype(np.float64)
    cmin, cmax = data.min(), data.max()
    cscale = cmax - cmin
    if cscale == 0:
        cscale = 1.0
    bytedata = (data - cmin) * (255.0 / cscale)
    return (bytedata.clip(0, 255) + 0.5).astype(np.uint8)


def sample_grid(samples, stretch: bool = True) -> np.ndarray:
    """generate_image's conversion + save_images' tiling for generator samples in (-1, 1): a device tensor / Var
    [n, h, w, c] (or [n, h*w*c] with square c = 3 images, as Generator returns them) -> uint8 host array
    [rows*h, cols*w, c].  stretch=True applies imsave's contrast stretch (what the reference's PNG holds)."""
    t = samples.data if isinstance(samples, F.Var) else samples
    if t.dim() == 2:
        side = int(round((t.shape[1] // 3) ** 0.5))
        t = t.reshape(t.shape[0], side, side, 3)
    t = t.contiguous()
    rows, cols = grid_shape(t.shape[0])
    grid = K.sample_grid(t, cols).cpu().numpy()
    return _bytescale(grid) if stretch else grid


def save_images(X, save_path, stretch: bool = True):
    """common/misc.py:215-244.  X: generator samples in (-1, 1) on the device (tensor / Var; the product path), or a
    host array exactly as the reference passes it (ints in [0, 255], or floats in [0, 1] which are scaled by 255.99)."""
    from PIL import Image

    if isinstance(X, (torch.Tensor, F.Var)):
        img = sample_grid(X, stretch=stretch)
    else:
        X = np.asarray(X)
        if np.issubdtype(X.dtype, np.floating):
            X = (255.99 * X).astype('uint8')
        rows, cols = grid_shape(X.shape[0])
        if X.ndim == 2:
            side = int(np.sqrt(X.shape[1]))
            X = np.reshape(X, (X.shape[0], side, side))
        h, w = X[0].shape[:2]
        img = np.zeros((h * rows, w * cols) + X.shape[3:])
        for n, x in enumerate(X):
            j, i = n // cols, n % cols
            img[j * h:j * h + h, i * w:i * w + w] = x
        img = _bytescale(img) if stretch else img.astype(np.uint8)
    if img.ndim == 3 and img.shape[2] == 1:
        img = img[:, :, 0]
    Image.fromarray(img).save(save_path)
    return img


def get_z(batchsize, n_hidden=128):
    """common/misc.py:247-259 (host NumPy RNG, like the reference)."""
    return np.random.normal(size=(batchsize, n_hidden)).astype(np.float32)


def get_loss(disc_real, disc_fake, loss_type='HINGE', player=None):
    """common/misc.py:310-394: (d_loss, g_loss) of the seven loss types as device scalars, from one fused kernel per
    player (functional.gan_loss).  A recording tape differentiates ONE player's loss (loss Vars push their logit
    gradients when the tape unwinds), so inside a gradient_tape pass player='d' or 'g' and take that element of the
    pair; the other one is None.  Without a tape (evaluation) both are returned."""
    store = get_store()
    real, fake = F.as_var(disc_real), F.as_var(disc_fake)
    if store.tape is not None and player is None:
        raise ValueError("get_loss inside a gradient tape needs player='d' or 'g' (one tape, one player's loss)")
    d_loss = g_loss = None
    if player in (None, 'd'):
        flat = lambda v: F.reshape(v, (-1,))  # noqa: E731
        d_loss = F.gan_loss(F.concat_rows(flat(real), flat(fake)), 'd', n_real=real.data.numel(), loss_type=loss_type)
    if player in (None, 'g'):
        g_loss = F.gan_loss(fake, 'g', loss_type=loss_type)
    return d_loss, g_loss


def optimistic_restore(session, save_file):
    """common/misc.py:275-307: restore every variable whose NAME and SHAPE match the checkpoint, skip the rest.
    `session` is ignored (the variable store is global state like TF's default graph); save_file is the prefix of a
    TensorFlow-1 checkpoint (`<prefix>.index` + `<prefix>.data-*`, read by common/tf_checkpoint.py), a state dict
    {name: array}, or the path of one written with numpy.savez / torch.save.  Returns the list of restored names (the
    reference prints them)."""
    store = get_store()
    if isinstance(save_file, (str, os.PathLike)) and os.path.exists(os.fspath(save_file) + '.index'):
        # a TensorFlow-1 tensor-bundle checkpoint (what saver.save wrote): read without TensorFlow, only the
        # variables whose names exist here (tf_checkpoint.CheckpointReader = tf.train.NewCheckpointReader)
        from .tf_checkpoint import CheckpointReader
        reader = CheckpointReader(os.fspath(save_file))
        shapes = reader.get_variable_to_shape_map()
        state = reader.state_dict([k for k in shapes if k in store.vars
                                   and list(store.vars[k].data.shape) == shapes[k]])
    elif isinstance(save_file, (str, os.PathLike)):
        path = os.fspath(save_file)
        if path.endswith('.npz'):
            with np.load(path) as z:
                state = {k: z[k] for k in z.files}
        else:
            state = torch.load(path, map_location='cpu', weights_only=False)   # the caller's own state dict
            state = {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in state.items()}
    else:
        state = save_file
    restored = store.load_state_dict(state, strict=False)
    print('\n--------variables to restore:--------')
    for name in restored:
        print(name)
    return restored


# ------------------------------------------------------------------------------------------------ checkpoints
def checkpoint_state(optimizers=()):
    """Everything `tf.train.Saver()` of the reference scripts would write (SNGAN/gan_cifar_resnet.py:588, 651-656), under
    TF-1 names: every variable of the store, and per optimiser (in creation order -- gen_opt before disc_opt, :520-526)
    the Adam slots `<variable>/Adam` (m), `<variable>/Adam_1` (v) and the non-slot scalars `beta1_power` /
    `beta2_power` (`beta1_power_1` ... for the second optimiser), which hold beta^(t+1) after t steps.
    optimizers: training.AdamState instances."""
    store = get_store()
    state = store.state_dict()
    for i, opt in enumerate(optimizers):
        flat = opt.flat
        for v, off in zip(flat.variables, flat.offsets):
            n = v.data.numel()
            state[v.key + '/Adam'] = flat.m[off:off + n].reshape(v.data.shape).cpu().numpy().copy()
            state[v.key + '/Adam_1'] = flat.v[off:off + n].reshape(v.data.shape).cpu().numpy().copy()
        suffix = '' if i == 0 else '_%d' % i
        state['beta1_power' + suffix] = np.float32(opt.beta1 ** (opt.t + 1))
        state['beta2_power' + suffix] = np.float32(opt.beta2 ** (opt.t + 1))
        # beta2^(t+1) underflows in fp32 after ~1000 steps, so the step count is also stored as an integer (not a TF name)
        state['ganb200/adam_step' + suffix] = np.int64(opt.t)
    return state


MAX_TO_KEEP = 5   # tf.train.Saver(max_to_keep=5), the default of the reference's savers


def _checkpoint_files(prefix: str):
    """Files that make up the checkpoint `prefix` (.npz, or a tensor bundle's .index / .data-* shards)."""
    d, base = os.path.split(prefix)
    out = []
    for f in os.listdir(d or '.'):
        if f == base + '.npz' or f == base + '.index' or f.startswith(base + '.data-'):
            out.append(os.path.join(d, f))
    return out


def _kept_checkpoints(directory: str):
    """all_model_checkpoint_paths of the directory's `checkpoint` state file, oldest first."""
    kept = []
    state_file = os.path.join(directory, 'checkpoint')
    if os.path.exists(state_file):
        with open(state_file) as fh:
            for line in fh:
                if line.startswith('all_model_checkpoint_paths:'):
                    kept.append(line.split(':', 1)[1].strip().strip('"'))
    return kept


def _rotate_checkpoints(path: str, max_to_keep: int, kept) -> None:
    """What tf.train.Saver does next to every save: the `checkpoint` state file of the directory
    (model_checkpoint_path + all_model_checkpoint_paths, oldest first) and deletion of everything but the last
    `max_to_keep` checkpoints.  kept: the list before this save (_kept_checkpoints)."""
    d = os.path.dirname(path) or '.'
    base = os.path.basename(path[:-4] if path.endswith('.npz') else path)
    kept = [k for k in kept if k != base] + [base]
    if max_to_keep and max_to_keep > 0:
        for old in kept[:-max_to_keep]:
            for f in _checkpoint_files(os.path.join(d, old)):
                try:
                    os.remove(f)
                except OSError:
                    pass
        kept = kept[-max_to_keep:]
    with open(os.path.join(d, 'checkpoint'), 'w') as fh:
        fh.write('model_checkpoint_path: "%s"\n' % base)
        for k in kept:
            fh.write('all_model_checkpoint_paths: "%s"\n' % k)


def save_checkpoint(save_file, optimizers=(), tf_bundle: bool = False, max_to_keep: int = MAX_TO_KEEP, extra=None):
    """saver.save(): one .npz of checkpoint_state(); tf_bundle=True writes TensorFlow's tensor-bundle container instead
    (`<save_file>.index` + `.data-00000-of-00001`, common/tf_checkpoint.py -- pure-Python checksums, slow for large
    models).  Like tf.train.Saver(max_to_keep=5) the directory keeps a `checkpoint` state file and only the last
    `max_to_keep` checkpoints (0 / None: keep everything).  extra: further entries, e.g. {'Variable': global_step}, the
    name TF gives the trainers' un-named global_step counter (ACGAN/train.py:124, Pix2Pix/train.py:520)."""
    state = checkpoint_state(optimizers)
    if extra:
        state.update({k: np.asarray(v) for k, v in extra.items()})
    path = os.fspath(save_file)
    kept = _kept_checkpoints(os.path.dirname(path) or '.')
    if tf_bundle:
        from .tf_checkpoint import write_checkpoint
        write_checkpoint(path, {k: np.asarray(v, dtype=np.float32) for k, v in state.items()})
    else:
        np.savez(path if path.endswith('.npz') else path + '.npz', **{k: np.asarray(v) for k, v in state.items()})
    _rotate_checkpoints(path, max_to_keep, kept)
    return sorted(state)


def restore_checkpoint(save_file, optimizers=()):
    """saver.restore() with optimistic_restore's name + shape rule, including the Adam slots and step counts of
    `optimizers` (same order as at save time).  Returns the restored names."""
    if isinstance(save_file, (str, os.PathLike)) and os.path.exists(os.fspath(save_file) + '.index'):
        from .tf_checkpoint import CheckpointReader
        state = CheckpointReader(os.fspath(save_file)).state_dict()      # a TensorFlow-1 tensor bundle
    elif isinstance(save_file, (str, os.PathLike)):
        path = os.fspath(save_file)
        with np.load(path if path.endswith('.npz') else path + '.npz') as z:
            state = {k: z[k] for k in z.files}
    else:
        state = dict(save_file)
    store = get_store()
    restored = store.load_state_dict({k: v for k, v in state.items() if k in store.vars}, strict=False)
    for i, opt in enumerate(optimizers):
        flat = opt.flat
        for v, off in zip(flat.variables, flat.offsets):
            n = v.data.numel()
            for slot, buf in (('/Adam', flat.m), ('/Adam_1', flat.v)):
                arr = state.get(v.key + slot)
                if arr is not None and tuple(arr.shape) == tuple(v.data.shape):
                    buf[off:off + n].copy_(torch.from_numpy(np.asarray(arr, dtype=np.float32)).reshape(-1))
                    restored.append(v.key + slot)
        suffix = '' if i == 0 else '_%d' % i
        b2p = state.get('beta2_power' + suffix)
        step = state.get('ganb200/adam_step' + suffix)
        if step is not None:
            opt.t = int(step)
            restored.append('ganb200/adam_step' + suffix)
        elif b2p is not None and 0.0 < opt.beta2 < 1.0:
            if 0.0 < float(b2p) < 1.0:
                opt.t = max(int(round(np.log(float(b2p)) / np.log(opt.beta2))) - 1, 0)
            else:
                # a TensorFlow checkpoint whose beta2_power has underflowed to 0 in fp32: the optimiser is fully warmed
                # up (sqrt(1 - beta2^t) = 1 to fp32 precision); any large t reproduces TF's learning-rate factor
                opt.t = 1 << 20
            restored.append('beta2_power' + suffix)
    return restored
