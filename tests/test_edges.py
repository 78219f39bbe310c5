"""The step either side of the hot path: common.misc (sample grid, get_z, get_loss,
optimistic_restore, TF-1 checkpoint names incl. the Adam slots) and common.plot.  CPU tests use the host-logic mode
(kernels recorded, not run); the GPU tests check the sample-grid kernel against a NumPy restatement of
generate_image + save_images (SNGAN/gan_cifar_resnet.py:536-539, common/misc.py:215-244)."""
import os
import pickle

import numpy as np
import pytest
import torch

from tests.test_host_logic import host  # noqa: F401


def _reference_grid(samples_pm1):
    """NumPy restatement of generate_image + save_images up to the imsave call: returns the float grid it passes on."""
    X = ((samples_pm1 + 1.) * (255. / 2)).astype('int32')                   # gan_cifar_resnet.py:538
    n_samples = X.shape[0]
    rows = int(np.sqrt(n_samples))                                         # misc.py:221-226
    while n_samples % rows != 0:
        rows -= 1
    nh, nw = rows, int(n_samples / rows)
    h, w = X[0].shape[:2]
    img = np.zeros((h * nh, w * nw, 3))                                    # misc.py:232-235
    for n, x in enumerate(X):                                              # misc.py:239-242
        j, i = int(n / nw), int(n % nw)
        img[j * h:j * h + h, i * w:i * w + w] = x
    return img


def _imsave_bytes(img):
    """scipy.misc.imsave -> toimage -> bytescale(cmin=None, cmax=None) of a non-uint8 array (scipy 1.0 source)."""
    cmin, cmax = img.min(), img.max()
    cscale = (cmax - cmin) or 1
    bytedata = (img - cmin) * (255.0 / cscale)
    return (bytedata.clip(0, 255) + 0.5).astype(np.uint8)


def test_grid_shape_and_bytescale():
    from gan_lib_tensorflow_b200.common import misc

    assert misc.grid_shape(100) == (10, 10) and misc.grid_shape(64) == (8, 8)
    assert misc.grid_shape(50) == (5, 10) and misc.grid_shape(7) == (1, 7) and misc.grid_shape(12) == (3, 4)
    rs = np.random.RandomState(0)
    img = rs.randint(3, 250, size=(8, 8, 3)).astype("float64")
    np.testing.assert_array_equal(misc._bytescale(img), _imsave_bytes(img))
    assert misc._bytescale(np.full((2, 2), 7.0)).max() == 0      # constant image: cscale falls back to 1


def test_save_images_host_path_matches_the_reference_loop(tmp_path):
    from PIL import Image

    from gan_lib_tensorflow_b200.common import misc

    rs = np.random.RandomState(1)
    samples = rs.uniform(-1, 1, size=(12, 8, 8, 3)).astype("float32")
    X = ((samples + 1.) * (255. / 2)).astype('int32')
    out = misc.save_images(X, tmp_path / "s.png")
    np.testing.assert_array_equal(out, _imsave_bytes(_reference_grid(samples)))
    np.testing.assert_array_equal(np.asarray(Image.open(tmp_path / "s.png")), out)
    # floats in [0, 1] are scaled by 255.99 (misc.py:217-218)
    f = rs.uniform(size=(4, 6, 6, 3))
    out_f = misc.save_images(f, tmp_path / "f.png", stretch=False)
    assert out_f.shape == (12, 12, 3) and out_f[0, 0, 0] == int(255.99 * f[0, 0, 0, 0])


def test_plot_tick_flush_log(tmp_path, capsys):
    from gan_lib_tensorflow_b200.common import plot

    plot.reset()
    plot.set_output_dir(str(tmp_path))
    for it in range(3):
        plot.plot('d_cost', 1.0 + it)
        plot.plot('g_cost', torch.tensor([0.5 * it]))       # device scalars are read back at flush time only
        plot.tick()
    plot.flush()
    printed = capsys.readouterr().out
    assert "iter 3" in printed and "d_cost: 2.0" in printed and "g_cost: 0.5" in printed
    with open(tmp_path / "log.pkl", "rb") as fh:
        log = pickle.load(fh)
    assert log["d_cost"] == {0: 1.0, 1: 2.0, 2: 3.0} and log["g_cost"][2] == 1.0
    plot.plot('d_cost', 10.0)
    plot.flush()
    with open(tmp_path / "log.pkl", "rb") as fh:
        assert pickle.load(fh)["d_cost"][3] == 10.0
    plot.reset()
    plot.set_output_dir('.')


def test_checkpoint_names_and_round_trip(host, tmp_path):  # noqa: F811
    """TF-1 names of a Saver checkpoint: variables, `<var>/Adam`, `<var>/Adam_1`, beta powers per optimiser
    (gen_opt first, SNGAN/gan_cifar_resnet.py:520-526); restore follows optimistic_restore's name + shape rule."""
    store, _ = host
    from gan_lib_tensorflow_b200.PGGAN import train as PT
    from gan_lib_tensorflow_b200.common import misc

    tr = PT.Trainer(block_count=1, trans=False, batch_size=2, seed=0)
    og, od = tr.players.opt["g"], tr.players.opt["d"]
    og.t, od.t = 3, 15
    og.flat.m.uniform_(-1, 1); og.flat.v.uniform_(0, 1); od.flat.m.uniform_(-1, 1); od.flat.v.uniform_(0, 1)
    names = misc.save_checkpoint(tmp_path / "model.ckpt-7", optimizers=(og, od))
    assert "g_net/G.Input/W" in names and "g_net/G.Input/W/Adam" in names and "g_net/G.Input/W/Adam_1" in names
    assert "d_net/D.Conv/filters/spectral_norm/u" in names and "d_net/D.Conv/filters/spectral_norm/u/Adam" not in names
    assert {"beta1_power", "beta2_power", "beta1_power_1", "beta2_power_1"} <= set(names)
    state = misc.checkpoint_state((og, od))
    assert abs(float(state["beta2_power"]) - 0.9 ** 4) < 1e-7 and abs(float(state["beta2_power_1"]) - 0.9 ** 16) < 1e-7
    assert float(state["beta1_power"]) == 0.0
    w = store.vars["g_net/G.Input/W"]
    off = og.flat.offsets[og.flat.variables.index(w)]
    np.testing.assert_array_equal(state["g_net/G.Input/W/Adam"].reshape(-1),
                                  og.flat.m[off:off + w.data.numel()].numpy())
    saved = {k: v.copy() for k, v in state.items()}
    # perturb everything, then restore
    for f in (og.flat, od.flat):
        f.params.add_(1.0); f.m.zero_(); f.v.zero_()
    og.t = od.t = 0
    restored = misc.restore_checkpoint(tmp_path / "model.ckpt-7", optimizers=(og, od))
    assert og.t == 3 and od.t == 15 and "g_net/G.Input/W/Adam_1" in restored
    now = misc.checkpoint_state((og, od))
    for k, v in saved.items():
        np.testing.assert_array_equal(now[k], v, err_msg=k)
    # optimistic rule: a shape mismatch and an unknown name are skipped silently
    bad = {"g_net/G.Input/W": np.zeros((3, 3), "float32"), "nope/W": np.zeros(2, "float32"),
           "g_net/G.Input/b": saved["g_net/G.Input/b"] + 1}
    got = misc.optimistic_restore(None, bad)
    assert got == ["g_net/G.Input/b"]


def test_checkpoint_rotation_and_step_counts(host, tmp_path):  # noqa: F811
    """tf.train.Saver(max_to_keep=5) behaviour of the savers in the reference scripts: the directory keeps the last five
    checkpoints and a `checkpoint` state file; the Adam step count survives a save / restore after beta2^(t+1) has
    underflowed in fp32, and a TensorFlow checkpoint with an underflowed beta2_power restores as a warmed-up optimiser."""
    store, _ = host
    from gan_lib_tensorflow_b200.PGGAN import train as PT
    from gan_lib_tensorflow_b200.common import misc, tf_checkpoint

    tr = PT.Trainer(block_count=0, trans=False, batch_size=2, seed=0)
    og, od = tr.players.opt["g"], tr.players.opt["d"]
    for step in range(8):
        og.t, od.t = 1000 * step, 5000 * step
        misc.save_checkpoint(tmp_path / ("model.ckpt-%d" % step), (og, od), extra={"Variable": np.int64(step)})
    files = sorted(f for f in os.listdir(tmp_path) if f.startswith("model.ckpt-"))
    assert files == ["model.ckpt-%d.npz" % i for i in range(3, 8)]
    lines = open(tmp_path / "checkpoint").read().splitlines()
    assert lines[0] == 'model_checkpoint_path: "model.ckpt-7"'
    assert lines[1:] == ['all_model_checkpoint_paths: "model.ckpt-%d"' % i for i in range(3, 8)]
    assert tf_checkpoint.latest_checkpoint(str(tmp_path)).endswith("model.ckpt-7")
    with np.load(tmp_path / "model.ckpt-7.npz") as z:
        assert int(z["Variable"]) == 7 and float(z["beta2_power_1"]) == 0.0          # 0.9^35001 underflows in fp32
    og.t = od.t = 0
    misc.restore_checkpoint(tmp_path / "model.ckpt-7", (og, od))
    assert og.t == 7000 and od.t == 35000
    # a TensorFlow-written state has no integer step: beta2_power == 0 means "fully warmed up"
    state = misc.checkpoint_state((og, od))
    state = {k: v for k, v in state.items() if not k.startswith("ganb200/")}
    og.t = od.t = 0
    misc.restore_checkpoint(state, (og, od))
    assert od.t >= 1 << 20 and og.t >= 1 << 20
    # max_to_keep=0 keeps everything
    misc.save_checkpoint(tmp_path / "model.ckpt-8", (og, od), max_to_keep=0)
    assert len([f for f in os.listdir(tmp_path) if f.startswith("model.ckpt-")]) == 6


def test_tf1_tensor_bundle_reader(host, tmp_path):  # noqa: F811
    """common/tf_checkpoint.py: the TensorFlow-1 checkpoint container (LevelDB-style table index + raw data shard) read
    without TensorFlow.  Known answers for the checksum, a hand-assembled table block, a multi-block round trip through
    the independent writer, corruption detection, and optimistic_restore / restore_checkpoint straight from a bundle."""
    import struct

    store, _ = host
    from gan_lib_tensorflow_b200.PGGAN import train as PT
    from gan_lib_tensorflow_b200.common import misc
    from gan_lib_tensorflow_b200.common import tf_checkpoint as T

    assert T.crc32c(b"123456789") == 0xE3069283                      # CRC-32C (Castagnoli) check value
    assert T.crc32c(b"") == 0 and T.masked_crc32c(b"") == 0xa282ead8
    assert T.crc32c(bytes(32)) == 0x8A9136AA and T.crc32c(b"\xff" * 32) == 0x62A8AB43   # RFC 3720 B.4 test vectors
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    # a block written by hand: keys "ab" -> "1", "abc" -> "22" (shares 2 bytes), one restart at 0
    blk = bytes([0, 2, 1]) + b"ab" + b"1" + bytes([2, 1, 2]) + b"c" + b"22" + struct.pack("<II", 0, 1)
    assert list(T._block_entries(blk)) == [(b"ab", b"1"), (b"abc", b"22")]
    assert T._put_varint(300) == bytes([0xAC, 0x02]) and T._varint(bytes([0xAC, 0x02]), 0) == (300, 2)
    e = T._parse_entry(T._msg([(1, 0, 1), (2, 2, T._msg([(2, 2, T._msg([(1, 0, 3)])), (2, 2, T._msg([(1, 0, 5)]))])),
                               (4, 0, 64), (5, 0, 60), (6, 5, 7)]))
    assert e["dtype"] == 1 and e["shape"] == [3, 5] and e["offset"] == 64 and e["size"] == 60 and e["crc32c"] == 7

    rs = np.random.RandomState(0)
    tensors = {"net/layer_%03d/W" % i: rs.standard_normal((3, i % 5 + 1)).astype("float32") for i in range(150)}
    tensors["global_step"] = np.array(1234, dtype=np.int64)
    tensors["beta1_power"] = np.float32(0.0).reshape(())
    prefix = str(tmp_path / "model.ckpt-1234")
    T.write_checkpoint(prefix, tensors, block_entries=16)            # 10 data blocks
    with open(prefix + ".index", "rb") as fh:
        assert struct.unpack("<Q", fh.read()[-8:])[0] == T.TABLE_MAGIC
    r = T.CheckpointReader(prefix, verify=True)
    assert r.get_variable_to_shape_map()["net/layer_007/W"] == [3, 3] and r.get_variable_to_shape_map()["global_step"] == []
    assert sorted(r.entries) == sorted(tensors) and r.has_tensor("beta1_power") and not r.has_tensor("nope")
    for k, v in tensors.items():
        got = r.get_tensor(k)
        assert got.dtype == v.dtype and got.shape == v.shape
        np.testing.assert_array_equal(got, v)
    assert T.latest_checkpoint(str(tmp_path)) == prefix
    with open(prefix + ".data-00000-of-00001", "r+b") as fh:          # flip one byte: the entry checksum catches it
        fh.seek(5)
        b = fh.read(1)
        fh.seek(5)
        fh.write(bytes([b[0] ^ 0xFF]))
    with pytest.raises(ValueError):
        T.CheckpointReader(prefix, verify=True).get_tensor("global_step")        # bytes 4..11 of the shard
    with pytest.raises(ValueError):
        T.read_index(prefix + ".data-00000-of-00001")

    # a trainer's state through the TF container and back (names incl. Adam slots, optimistic rule on the way in)
    tr = PT.Trainer(block_count=1, trans=False, batch_size=2, seed=0)
    og, od = tr.players.opt["g"], tr.players.opt["d"]
    og.t, od.t = 4, 9
    og.flat.m.uniform_(-1, 1)
    saved = misc.checkpoint_state((og, od))
    p2 = str(tmp_path / "pggan.ckpt-9")
    names = misc.save_checkpoint(p2, (og, od), tf_bundle=True)
    assert os.path.exists(p2 + ".index") and "g_net/G.Input/W/Adam" in names
    og.flat.params.add_(1.0)
    og.flat.m.zero_()
    og.t = od.t = 0
    restored = misc.restore_checkpoint(p2, (og, od))
    assert og.t == 4 and od.t == 9 and "g_net/G.Input/W/Adam" in restored
    now = misc.checkpoint_state((og, od))
    for k in ("g_net/G.Input/W", "g_net/G.Input/W/Adam", "d_net/D.Conv/Filters"):
        np.testing.assert_array_equal(now[k], saved[k], err_msg=k)
    og.flat.params.add_(1.0)
    got = misc.optimistic_restore(None, p2)
    assert "g_net/G.Input/W" in got and "g_net/G.Input/W/Adam" not in got      # variables of the store only
    np.testing.assert_array_equal(store.vars["g_net/G.Input/W"].data.numpy(), saved["g_net/G.Input/W"])


def test_cifar10_loader_interface(tmp_path, capsys):
    """common/data/cifar10.py: load() -> (train, dev) epoch factories over the pickled batches; partial batches are
    dropped; images and labels stay paired through the per-epoch reshuffle (shared RNG state)."""
    import gan_lib_tensorflow_b200.common as lib
    from gan_lib_tensorflow_b200.common.data import cifar10

    rs = np.random.RandomState(0)
    for name, n in [("data_batch_%d" % i, 10) for i in range(1, 6)] + [("test_batch", 7)]:
        labels = rs.randint(0, 10, size=n)
        data = np.tile(labels[:, None].astype("uint8"), (1, 3072))       # every pixel carries the label: pairing check
        with open(tmp_path / name, "wb") as fh:
            pickle.dump({b"data": data, b"labels": labels.tolist()}, fh)
    train_gen, dev_gen = cifar10.load(8, str(tmp_path))
    np.random.seed(1)
    epoch1 = list(train_gen())                                   # fresh gathers: safe to keep across epochs
    epoch2 = list(train_gen())
    assert len(epoch1) == 6 and len(list(dev_gen())) == 0 and epoch1[0][0].shape == (8, 3072)     # 50 // 8, 7 // 8
    for images, labels in epoch1 + epoch2:
        assert np.array_equal(images[:, 0], np.asarray(labels).astype("uint8"))
    assert not np.array_equal(epoch1[0][1], epoch2[0][1])                # reshuffled per epoch
    # same batch sequence and RNG consumption as two in-place shuffles under a saved / restored RNG state
    # (common/data/cifar10.py:30-33), epoch after epoch
    ref_labels = np.concatenate([np.asarray(pickle.load(open(tmp_path / ("data_batch_%d" % i), "rb"))[b"labels"])
                                 for i in range(1, 6)])
    np.random.seed(1)
    for epoch in (epoch1, epoch2):
        np.random.shuffle(ref_labels)
        assert np.array_equal(np.concatenate([l for _, l in epoch]), ref_labels[:48])
    assert np.random.randint(1 << 30) == (np.random.seed(1), [np.random.permutation(50) for _ in range(2)],
                                          np.random.randint(1 << 30))[2]
    lib.print_model_settings({"BATCH_SIZE": 64, "lower": 1, "T": 2})
    out = capsys.readouterr().out
    assert "BATCH_SIZE: 64" in out and "lower" not in out


def test_get_loss_needs_a_player_inside_a_tape(host):  # noqa: F811
    store, rec = host
    from gan_lib_tensorflow_b200.common import misc

    real, fake = torch.zeros(4), torch.zeros(4)
    d, g = misc.get_loss(real, fake, 'LSGAN')
    assert d is not None and g is not None and rec.names().count("ganb_gan_loss") == 2
    with store.gradient_tape():
        with pytest.raises(ValueError):
            misc.get_loss(real, fake, 'HINGE')
        d, g = misc.get_loss(real, fake, 'HINGE', player='d')
        assert d is not None and g is None
    z = misc.get_z(5, 7)
    assert z.shape == (5, 7) and z.dtype == np.float32


def test_reference_train_loop_call_sequence(host, tmp_path, capsys):  # noqa: F811
    """SNGAN.gan_cifar_resnet.train (the loop of gan_cifar_resnet.py:599-658) in host-logic mode: G-step skipped at
    iteration 0, N_CRITIC critic steps per iteration, the gen_cost fetch, dev cost + sample grid every `sample_every`,
    flush + checkpoint, tick."""
    store, rec = host
    from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P
    from gan_lib_tensorflow_b200.common import plot

    plot.reset()
    tr = P.train(iters=3, out_dir=str(tmp_path), batch_size=4, capture=False, fetch_gen_cost=True, sample_every=2)
    names = rec.names()
    assert names.count("ganb_adam") == 2 + 3 * P.N_CRITIC          # G-steps at iterations 1, 2; 5 D-steps each
    assert tr.gen_opt.t == 2 and tr.disc_opt.t == 15
    assert names.count("ganb_sample_grid") == 1                     # iteration 1 only (1 % 2 == 1)
    # gan_loss launches: 15 D-steps + 2 G-steps + 15 gen_cost fetches + 2 dev batches
    assert names.count("ganb_gan_loss") == 15 + 2 + 15 + 2
    assert os.path.exists(tmp_path / "samples_1.png") and os.path.exists(tmp_path / "log.pkl")
    # three checkpoints and the Saver's `checkpoint` state file (max_to_keep = 5 not reached)
    assert sorted(os.listdir(tmp_path / "checkpoint")) == ["checkpoint", "model.ckpt-0.npz", "model.ckpt-1.npz",
                                                           "model.ckpt-2.npz"]
    with open(tmp_path / "log.pkl", "rb") as fh:
        log = pickle.load(fh)
    assert set(log) == {"d_cost", "g_cost", "dev_cost"} and sorted(log["d_cost"]) == [0, 1, 2] and list(log["dev_cost"]) == [1]
    assert "iter 2" in capsys.readouterr().out
    with np.load(tmp_path / "checkpoint" / "model.ckpt-2.npz") as z:
        assert "Generator/G.Input/W/Adam_1" in z.files and "beta2_power_1" in z.files
        assert abs(float(z["beta2_power"]) - 0.9 ** 3) < 1e-6 and abs(float(z["beta2_power_1"]) - 0.9 ** 16) < 1e-6
    plot.reset()
    plot.set_output_dir('.')


def test_acgan_and_pggan_train_loops_call_sequence(host, tmp_path, capsys):  # noqa: F811
    """ACGAN/train.py:176-233 and PGGAN/train.py:168-226 through Trainer.train (training.reference_loop), host-logic
    mode: G step skipped at step 0 for ACGAN but not for PGGAN, n_dis critic steps, progress line, dev cost + sample
    grid + checkpoint every out_image_interval, resume."""
    store, rec = host
    from gan_lib_tensorflow_b200 import framework
    from gan_lib_tensorflow_b200.ACGAN import train as AT
    from gan_lib_tensorflow_b200.PGGAN import train as PT
    from gan_lib_tensorflow_b200.common import plot

    plot.reset()
    rs = np.random.RandomState(0)

    def cifar(n_batches, b=4):
        data = rs.randint(0, 256, size=(n_batches, b, 3072)).astype("int32")
        labels = rs.randint(0, 10, size=(n_batches, b)).astype("int32")
        return lambda: ((data[i], labels[i]) for i in range(n_batches))

    out_a = tmp_path / "acgan"
    os.makedirs(out_a)
    tr = AT.Trainer(batch_size=4, gradient_penalty=False, seed=0, max_iter=10)
    n0 = len(rec.calls)
    tr.train(3, cifar(5), cifar(2), n_dis=2, out_dir=str(out_a), display_interval=1, out_image_interval=2,
             capture_after=None)
    names = rec.names()[n0:]
    assert tr.players.opt["g"].t == 2 and tr.players.opt["d"].t == 6        # G skipped at step 0; 2 D-steps per step
    assert names.count("ganb_adam") == 8 and names.count("ganb_sample_grid") == 1
    assert sorted(os.listdir(out_a / "checkpoint")) == ["checkpoint", "model.ckpt-1.npz"] and os.path.exists(out_a / "samples_1.png")
    printed = capsys.readouterr().out
    assert "step: 0, d_loss_gan:" in printed and "step: 2, g_loss_gan:" in printed and "dev_cost" in printed
    with open(out_a / "log.pkl", "rb") as fh:
        assert list(pickle.load(fh)["dev_cost"]) == [1]
    # resume: a fresh trainer picks up parameters and optimiser step counts
    framework.reset_default_graph("cpu", u_seed=2)
    tr2 = AT.Trainer(batch_size=4, gradient_penalty=False, seed=1, max_iter=10)
    tr2.train(0, cifar(1), None, out_dir=str(out_a), restore=True)
    assert tr2.players.opt["g"].t == 1 and tr2.players.opt["d"].t == 4       # state of the checkpoint written at step 1
    assert "Restore model from: model.ckpt-1" in capsys.readouterr().out

    framework.reset_default_graph("cpu", u_seed=2)
    plot.reset()
    out_p = tmp_path / "pggan"
    os.makedirs(out_p)
    imgs = rs.uniform(-1, 1, size=(3, 2, 8, 8, 3)).astype("float32")
    trp = PT.Trainer(block_count=1, trans=True, batch_size=2, seed=0, model="resnet")
    trp.train(lambda: (imgs[i] for i in range(3)), lambda: ((imgs[i], None) for i in range(2)), max_iter=2, n_dis=2,
              out_dir=str(out_p), display_interval=1, out_image_interval=2, capture_after=None)
    assert trp.players.opt["g"].t == 2 and trp.players.opt["d"].t == 4      # PGGAN runs the G step at step 0 as well
    assert abs(trp.alpha(1) - 0.5) < 1e-12
    assert os.path.exists(out_p / "samples_1.png") and os.path.exists(out_p / "checkpoint" / "model.ckpt-1.npz")
    assert "step: 1, g_loss:" in capsys.readouterr().out
    plot.reset()
    plot.set_output_dir('.')


def test_imagenet_train_loop_call_sequence(host, tmp_path, monkeypatch):  # noqa: F811
    """gan_imagNet_resnet.train (gan_imagNet_resnet.py:546-705): the CIFAR loop around the ImageNet Trainer, 25 fixed
    samples of one class from SAMPLE_LABELS per grid (file name carries the label), no dev cost."""
    store, rec = host
    from gan_lib_tensorflow_b200.SNGAN import gan_imagNet_resnet as P
    from gan_lib_tensorflow_b200.common import plot

    monkeypatch.setattr(P, "DIM_G", 16)
    monkeypatch.setattr(P, "DIM_D", 16)
    plot.reset()
    np.random.seed(0)
    tr = P.train(iters=2, out_dir=str(tmp_path), batch_size=2, capture=False, sample_every=2, flush_until=0,
                 flush_every=2)
    names = rec.names()
    assert tr.output_dim == 49152 and tr.gen_opt.t == 1 and tr.disc_opt.t == 2 * 5
    assert names.count("ganb_sample_grid") == 1
    pngs = [f for f in os.listdir(tmp_path) if f.endswith(".png")]
    assert len(pngs) == 1 and pngs[0].startswith("samples_1_") and int(pngs[0][:-4].split("_")[2]) in P.SAMPLE_LABELS
    from PIL import Image
    assert np.asarray(Image.open(tmp_path / pngs[0])).shape == (5 * 128, 5 * 128, 3)
    with open(tmp_path / "log.pkl", "rb") as fh:
        assert set(pickle.load(fh)) == {"d_cost", "g_cost"}           # the dev cost is commented out in this script
    assert os.path.exists(tmp_path / "checkpoint" / "model.ckpt-1.npz")
    plot.reset()
    plot.set_output_dir('.')


def test_pix2pix_train_loop_call_sequence(host, tmp_path):  # noqa: F811
    """Pix2Pix/train.py:694-772 through Trainer.train, host-logic mode: n_dis critic steps then the generator step per
    batch, should(freq) firing on multiples and on the last step, display grid and validation images."""
    store, rec = host
    from gan_lib_tensorflow_b200.Pix2Pix import train as PT

    tr = PT.Trainer(ngf=8, ndf=8, size=256, max_steps=3)
    batches = [(torch.zeros(1, 256, 256, 3), torch.zeros(1, 256, 256, 3)) for _ in range(2)]
    lines = []
    n0 = len(rec.calls)
    tr.train(batches, n_dis=2, progress_freq=2, display_freq=2, save_freq=3, val_batches=batches[:1],
             out_dir=str(tmp_path), capture_after=None, log=lambda *a: lines.append(" ".join(str(x) for x in a)))
    names = rec.names()[n0:]
    assert tr.players.opt["d"].t == 6 and tr.players.opt["g"].t == 3 and tr.global_step == 3
    assert names.count("ganb_adam") == 9
    assert [l.split()[0] for l in lines].count("progress") == 2          # steps 1 (multiple of 2) and 2 (last)
    assert any(l.startswith("gen_loss_L1") for l in lines) and "evaluated image val_0000.png" in lines
    assert sorted(f for f in os.listdir(tmp_path)) == ["train_00000002.png", "train_00000003.png", "val_0000.png"]
    assert names.count("ganb_sample_grid") == 3


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w,dtype", [(100, 32, 32, torch.float32), (12, 8, 16, torch.float32),
                                         (7, 4, 4, torch.bfloat16), (64, 128, 128, torch.float32)])
def test_sample_grid_kernel_matches_generate_image(n, h, w, dtype, tmp_path):
    from PIL import Image

    from gan_lib_tensorflow_b200 import framework
    from gan_lib_tensorflow_b200.common import misc

    framework.reset_default_graph("cuda")
    try:
        rs = np.random.RandomState(2)
        s = np.tanh(rs.standard_normal((n, h, w, 3)) * 2).astype("float32")
        s[0, 0, 0, :] = [-1.0, 1.0, 0.0]
        t = torch.from_numpy(s).cuda().to(dtype)
        s = t.float().cpu().numpy()
        ref = _reference_grid(s)
        got = misc.sample_grid(t, stretch=False)
        assert got.dtype == np.uint8 and got.shape == ref.shape
        np.testing.assert_array_equal(got, ref.astype(np.uint8))            # bit-exact integer work
        out = misc.save_images(t, tmp_path / "g.png")
        np.testing.assert_array_equal(out, _imsave_bytes(ref))
        np.testing.assert_array_equal(np.asarray(Image.open(tmp_path / "g.png")), out)
        if h == w:   # the [n, h*w*3] layout the Generator returns
            np.testing.assert_array_equal(misc.sample_grid(t.reshape(n, -1), stretch=False), got)
    finally:
        framework.set_store(None)


@pytest.mark.gpu
@pytest.mark.parametrize("loss_type", ["HINGE", "WGAN", "WGAN-GP", "LSGAN", "CGAN", "Modified_MiniMax", "MiniMax"])
def test_get_loss_pair_matches_the_oracle(loss_type):
    from gan_lib_tensorflow_b200 import framework
    from gan_lib_tensorflow_b200 import functional as F
    from gan_lib_tensorflow_b200.common import misc
    from oracle import acgan as OA

    store = framework.reset_default_graph("cuda")
    try:
        rs = np.random.RandomState(3)
        real = rs.standard_normal(24).astype("float32")
        fake = rs.standard_normal(24).astype("float32")
        d, g = misc.get_loss(torch.from_numpy(real).cuda(), torch.from_numpy(fake).cuda(), loss_type)
        rt = torch.from_numpy(real).double().requires_grad_(True)
        ft = torch.from_numpy(fake).double().requires_grad_(True)
        od, og = OA.get_loss(rt, ft, loss_type)
        assert abs(float(d.data.item()) - od.item()) < 1e-5 and abs(float(g.data.item()) - og.item()) < 1e-5
        for player, loss in (("d", od), ("g", og)):
            rv = F.Var(torch.from_numpy(real).cuda(), requires_grad=True)
            fv = F.Var(torch.from_numpy(fake).cuda(), requires_grad=True)
            with store.gradient_tape() as tape:
                pair = misc.get_loss(rv, fv, loss_type, player=player)
                tape.backward(pair[0] if player == "d" else pair[1])
            gr, gf = torch.autograd.grad(loss, [rt, ft], allow_unused=True, retain_graph=True)
            if gr is not None and player == "d":
                np.testing.assert_allclose(rv.grad.cpu().numpy(), gr.numpy(), atol=1e-6)
            np.testing.assert_allclose(fv.grad.cpu().numpy(), gf.numpy(), atol=1e-6)
    finally:
        framework.set_store(None)


@pytest.mark.gpu
def test_reference_train_loop_runs_and_resumes(tmp_path):
    """gan_cifar_resnet.train for a few iterations on synthetic batches with CUDA-graph capture: finite losses, a
    sample grid, log.pkl and TF-1-named checkpoints; a second call with restore=True resumes from the last checkpoint
    (parameters, Adam slots and step counts)."""
    from gan_lib_tensorflow_b200 import framework
    from gan_lib_tensorflow_b200.SNGAN import gan_cifar_resnet as P
    from gan_lib_tensorflow_b200.common import misc, plot

    plot.reset()
    framework.reset_default_graph("cuda", u_seed=2)
    try:
        tr = P.train(iters=4, out_dir=str(tmp_path), batch_size=16, capture=True, sample_every=2)
        torch.cuda.synchronize()
        assert tr._graphs and np.isfinite(float(tr.d_loss.item())) and np.isfinite(float(tr.g_loss.item()))
        with open(tmp_path / "log.pkl", "rb") as fh:
            log = pickle.load(fh)
        assert sorted(log["d_cost"]) == [0, 1, 2, 3] and sorted(log["dev_cost"]) == [1, 3]
        assert all(np.isfinite(v) for v in log["d_cost"].values()) and 0.0 <= log["dev_cost"][3] < 10.0
        from PIL import Image
        img = np.asarray(Image.open(tmp_path / "samples_3.png"))
        assert img.shape == (320, 320, 3) and img.std() > 1.0
        state = misc.checkpoint_state((tr.gen_opt, tr.disc_opt))
        with np.load(tmp_path / "checkpoint" / "model.ckpt-3.npz") as z:
            for k in ("Generator/G.Block.1.Conv1/Filters", "Discriminator/D.Block.2.Conv1/filters/spectral_norm/u",
                      "Discriminator/D.Output/W/Adam_1"):
                np.testing.assert_array_equal(z[k], state[k], err_msg=k)
        # resume in a fresh graph
        plot.reset()
        framework.reset_default_graph("cuda", u_seed=3)
        tr2 = P.Trainer(batch_size=16, seed=1)
        before = tr2.store.vars["Generator/G.Input/W"].data.clone()
        P.train(iters=0, out_dir=str(tmp_path), batch_size=16, restore=True, trainer=tr2)
        assert tr2.gen_opt.t == 3 and tr2.disc_opt.t == 20
        assert not torch.equal(before, tr2.store.vars["Generator/G.Input/W"].data)
        np.testing.assert_array_equal(tr2.store.vars["Generator/G.Input/W"].data.cpu().numpy(),
                                      state["Generator/G.Input/W"])
    finally:
        plot.reset()
        plot.set_output_dir('.')
        framework.set_store(None)


@pytest.mark.gpu
def test_acgan_and_pggan_train_loops_run_with_graphs(tmp_path):
    """Trainer.train of ACGAN (with the gradient penalty) and PGGAN (ResNet variant, fade-in) for a few steps on the GPU:
    CUDA graphs captured after step 1, eager dev-cost evaluations and fixed-noise samples between replays."""
    from PIL import Image

    from gan_lib_tensorflow_b200 import framework
    from gan_lib_tensorflow_b200.ACGAN import train as AT
    from gan_lib_tensorflow_b200.PGGAN import train as PT
    from gan_lib_tensorflow_b200.common import plot

    rs = np.random.RandomState(0)
    data = rs.randint(0, 256, size=(4, 8, 3072)).astype("int32")
    labels = rs.randint(0, 10, size=(4, 8)).astype("int32")
    cifar = lambda: ((data[i], labels[i]) for i in range(4))  # noqa: E731
    try:
        plot.reset()
        framework.reset_default_graph("cuda", u_seed=2)
        out_a = tmp_path / "acgan"
        os.makedirs(out_a)
        tr = AT.Trainer(batch_size=8, gradient_penalty=True, seed=0, max_iter=10)
        tr.train(4, cifar, cifar, n_dis=2, out_dir=str(out_a), display_interval=2, out_image_interval=2)
        torch.cuda.synchronize()
        assert tr.players.captured("d") and tr.players.captured("g")
        assert tr.players.opt["g"].t == 3 and tr.players.opt["d"].t == 8
        with open(out_a / "log.pkl", "rb") as fh:
            dev = pickle.load(fh)["dev_cost"]
        assert sorted(dev) == [1, 3] and all(np.isfinite(v) for v in dev.values())
        assert np.asarray(Image.open(out_a / "samples_3.png")).shape == (320, 320, 3)

        plot.reset()
        framework.reset_default_graph("cuda", u_seed=2)
        out_p = tmp_path / "pggan"
        os.makedirs(out_p)
        imgs = rs.uniform(-1, 1, size=(4, 4, 8, 8, 3)).astype("float32")
        trp = PT.Trainer(block_count=1, trans=True, inputs_norm=True, batch_size=4, seed=0, model="resnet")
        trp.train(lambda: (imgs[i] for i in range(4)), lambda: (imgs[i] for i in range(2)), max_iter=4, n_dis=2,
                  out_dir=str(out_p), display_interval=2, out_image_interval=2)
        torch.cuda.synchronize()
        assert trp.players.captured("d") and trp.players.opt["g"].t == 4 and trp.players.opt["d"].t == 8
        with open(out_p / "log.pkl", "rb") as fh:
            dev = pickle.load(fh)["dev_cost"]
        assert sorted(dev) == [1, 3] and all(np.isfinite(v) for v in dev.values())
        assert np.asarray(Image.open(out_p / "samples_3.png")).shape == (80, 80, 3)
    finally:
        plot.reset()
        plot.set_output_dir('.')
        framework.set_store(None)


@pytest.mark.gpu
def test_pix2pix_train_loop_runs_with_graphs(tmp_path):
    """Pix2Pix Trainer.train on the GPU: graphs captured after the first step, display / validation grids written from
    eager generator passes between replays."""
    from PIL import Image

    from gan_lib_tensorflow_b200 import framework
    from gan_lib_tensorflow_b200.Pix2Pix import train as PT

    framework.reset_default_graph("cuda", u_seed=2)
    try:
        g = torch.Generator(device="cuda").manual_seed(0)
        batches = [(torch.rand(1, 256, 256, 3, device="cuda", generator=g) * 2 - 1,
                    torch.rand(1, 256, 256, 3, device="cuda", generator=g) * 2 - 1) for _ in range(2)]
        tr = PT.Trainer(ngf=8, ndf=8, size=256, max_steps=3, seed=0)
        lines = []
        tr.train(batches, n_dis=2, progress_freq=1, display_freq=3, save_freq=3, val_batches=batches[:1],
                 out_dir=str(tmp_path), log=lambda *a: lines.append(" ".join(str(x) for x in a)))
        torch.cuda.synchronize()
        assert tr.players.captured("d") and tr.players.captured("g") and tr.global_step == 3
        vals = [float(l.split()[1]) for l in lines if l.startswith(("discrim_loss", "gen_loss_GAN", "gen_loss_L1"))]
        assert len(vals) == 9 and all(np.isfinite(v) for v in vals)
        assert np.asarray(Image.open(tmp_path / "train_00000003.png")).shape == (256, 768, 3)
        assert os.path.exists(tmp_path / "val_0000.png")
    finally:
        framework.set_store(None)
