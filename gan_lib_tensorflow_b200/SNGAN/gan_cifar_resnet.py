"""SNGAN CIFAR-10 ResNet (conditional) on the B200 layer ops: same Generator / Discriminator / losses / optimiser
as the reference script SNGAN/gan_cifar_resnet.py, restated as an eager step that can be captured into CUDA graphs.

Reference structure kept:
  * constants (:38-68), Generator (:237-263), Discriminator (:266-313), hinge losses (:376-378, :492),
    LR decay (:454-457), two Adam(beta1=0, beta2=0.9) optimisers (:521-526), loop order (:599-620):
    G-step (skipped at iteration 0) then N_CRITIC D-steps;
  * the two per-device towers (DEVICES always has two entries, :70-75) are batched into one call whose batch
    statistics are computed per tower (`stat_towers(2)`), which is what the reference's per-tower tf.nn.moments do.
"""
from __future__ import annotations

import contextlib
import math
import os

import numpy as np
import torch

from .. import functional as F
from .. import kernels as K
from ..common import resnet_block as rb
from ..common.ops import conv2d as conv2d_ops
from ..common.ops import embedding as embedding_ops
from ..common.ops import linear as linear_ops
from ..training import AdamState  # noqa: F401  (tf.train.AdamOptimizer over a network's flat buffers)
from ..framework import Var, get_store
from ..framework import aux_stream as framework_aux_stream
from ..framework import _side_stream as framework_side_stream

BATCH_SIZE = 64  # Critic batch size
GEN_BS_MULTIPLE = 2  # Generator batch size, as a multiple of BATCH_SIZE
ITERS = 100000
DIM_G = 128
DIM_D = 128
NORMALIZATION_G = True
NORMALIZATION_D = False
OUTPUT_DIM = 3072
LR = 0.0002
DECAY = True
N_CRITIC = 5
CONDITIONAL = True
ACGAN = False
VOCAB_SIZE = 10
EMBEDDING_DIM = 300
N_TOWERS = 2  # len(DEVICES) is always 2 in the reference (:73-75)
# Generator step: the frozen critic's pass over the fake batch as two half-batch chains on two streams (Trainer._g_rest,
# framework.Tape.branch).  Correct (tests pass, gradients bit-identical) but measured SLOWER: 3.30 vs 3.15 ms per pair on
# one box -- the half-batch kernels lose the CTA-pair route at 16x16 and the tensor-core kernels of the two chains
# serialise anyway.  Opt-in (GANB_SPLIT_CRITIC=1) for further experiments.
SPLIT_CRITIC_PASS = os.environ.get("GANB_SPLIT_CRITIC", "0") == "1"
# where the generator step's G forward forks off in the pair schedule: "early" = next to the critic step's own generator
# pass, "late" = behind it, next to the critic's forward + backward (A/B: profiles/r02_pair_fork_ab.txt)
PAIR_FORK = os.environ.get("GANB_PAIR_FORK", "early")
# SM budgets (data-gradient chain, filter-gradient side stream) of the critic's backward pass; "0:0" = every kernel may take
# the whole GPU (framework.Tape.sm_split)
CRITIC_BWD_SPLIT = tuple(int(v) for v in os.environ.get("GANB_CRITIC_BWD_SPLIT", "0:0").split(":"))
if not any(CRITIC_BWD_SPLIT):
    CRITIC_BWD_SPLIT = None
# critic power iteration on a side stream next to the generator pass instead of heading the critic chain: no measurable
# change (3.07 vs 3.07 ms per pair), opt-in
EMBED_BRANCH = os.environ.get("GANB_EMBED_BRANCH", "1") != "0"   # critic label branch as a Tape branch (A/B switch)
PREFETCH_SN = os.environ.get("GANB_PREFETCH_SN", "0") == "1"
# late fork only: SMs for the tensor-core kernels of (generator forward on the aux stream, critic step), "0:0" = no limits
SM_SPLIT = tuple(int(v) for v in os.environ.get("GANB_SM_SPLIT", "0:0").split(":"))

BF16 = torch.bfloat16


def _normalize_kind(name, labels):
    """The script's own Normalize dispatch (gan_cifar_resnet.py:88-109)."""
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
    return rb.ResidualBlock(inputs, input_dim, output_dim, filter_size, name, labels=labels,
                            normalize_kind=lambda nm: _normalize_kind(nm, labels), **kw)


def Generator(n_samples_, labels, noise=None, reuse=False):
    """gan_cifar_resnet.py:237-263. Returns Var [n, 3072] (NHWC-flattened images in (-1, 1))."""
    store = get_store()
    with store.variable_scope("Generator", reuse=reuse):
        if noise is None:
            noise = torch.randn(n_samples_, 128, device=store.device)
        noise = F.as_var(noise)
        # every input of a (conditional) batch norm is stored in bf16 (DESIGN.md "Data layout"): G's residual stream
        output = linear_ops.Linear(noise, 128, 4 * 4 * DIM_G * 8, 'G.Input', out_dtype=rb._adt())
        output = F.reshape(output, (-1, 4, 4, DIM_G * 8))
        output = _block(output, DIM_G * 8, DIM_G * 2, 3, 'G.Block.1', resample='up', labels=labels, biases=True,
                        out_dtype=rb._adt(), out_bn_stats=NORMALIZATION_G)
        output = _block(output, DIM_G * 2, DIM_G * 2, 3, 'G.Block.2', resample='up', labels=labels, biases=True,
                        out_dtype=rb._adt(), out_bn_stats=NORMALIZATION_G)
        output = _block(output, DIM_G * 2, DIM_G * 2, 3, 'G.Block.3', resample='up', labels=labels, biases=True,
                        out_dtype=rb._adt(), out_bn_stats=NORMALIZATION_G)
        output, _ = rb._norm_act('G.OutputNorm', output, labels, _normalize_kind('G.OutputNorm', labels), 'relu')
        output = conv2d_ops.Conv2D(output, DIM_G * 2, 3, 3, 1, 'G.Output', he_init=False)
        output = F.activation(output, 'tanh')
        return F.reshape(output, (-1, OUTPUT_DIM))


def Discriminator(inputs, labels, update_collection=None, reuse=False):
    """gan_cifar_resnet.py:266-313. Returns (output_wgan Var [n], None)."""
    store = get_store()
    with store.variable_scope("Discriminator", reuse=reuse):
        output = F.reshape(F.as_var(inputs), (-1, 32, 32, 3))
        # The label branch (embedding + linear layer) only meets the image branch at the concat below.  On a recording tape
        # it is a Tape branch on a side stream: in the backward pass its three small launches (~50 us of CUDA-core GEMM
        # and scatter) then run NEXT TO D.Block.1's backward instead of in front of it on the data-gradient chain.
        tape = store.tape if (EMBED_BRANCH and reuse and not K.host_logic_only()) else None
        marker = tape.fork() if tape is not None else None
        output = rb.OptimizedResBlockDisc1(output, DIM_D=DIM_D, spectral_normed=True,
                                           update_collection=update_collection, biases=True,
                                           name_prefix='D.Block.1')
        # embedding labels, and concatenate to 'output'.
        # (a stream of its own: the aux stream carries the generator pass of the pair schedule at this point)
        with (tape.branch(marker, framework_side_stream(store.device, 3)) if tape is not None else contextlib.nullcontext()):
            embedding_y = embedding_ops.embed_y(labels, VOCAB_SIZE, EMBEDDING_DIM)
            embedding_y = linear_ops.Linear(embedding_y, EMBEDDING_DIM, DIM_D, 'D.Embedding_y', spectral_normed=True,
                                            update_collection=update_collection, biases=True)  # (N, DIM_D)
        if tape is not None:
            tape.join_forward(marker)
        pre = F.concat_label_map(output, embedding_y, act='relu')
        output = _block(None, DIM_D * 2, DIM_D, 3, 'D.Block.2', spectral_normed=True,
                        update_collection=update_collection, resample='down', labels=labels, biases=True,
                        pre_activated=pre)
        output = _block(output, DIM_D, DIM_D, 3, 'D.Block.3', spectral_normed=True,
                        update_collection=update_collection, resample=None, labels=labels, biases=True)
        output = _block(output, DIM_D, DIM_D, 3, 'D.Block.4', spectral_normed=True,
                        update_collection=update_collection, resample=None, labels=labels, biases=True)
        output = F.act_mean_hw(output, 'relu')
        output_wgan = linear_ops.Linear(output, DIM_D, 1, 'D.Output', spectral_normed=True,
                                        update_collection=update_collection)
        output_wgan = F.reshape(output_wgan, (-1,))
        return output_wgan, None


def lr_decay(iteration: int) -> float:
    """gan_cifar_resnet.py:454-457"""
    if not DECAY:
        return 1.0
    return max(0.0, 1.0 - iteration / 100000.0) if iteration < 50000 else 0.5


class Trainer:
    """Builds the variables in the reference's graph-construction order and runs D / G training steps.

    The model-specific pieces are class attributes so that the ImageNet script (same graph structure,
    SNGAN/gan_imagNet_resnet.py:336-500) reuses the step logic: see gan_imagNet_resnet.Trainer."""

    generator = staticmethod(lambda *a, **k: Generator(*a, **k))
    discriminator = staticmethod(lambda *a, **k: Discriminator(*a, **k))
    output_dim = OUTPUT_DIM          # flattened image size (CHW order on the input side)
    image_hw = 1024                  # pixels per image
    n_classes = 10                   # fake labels are uniform in [0, n_classes)
    n_towers = N_TOWERS
    gen_bs_multiple = GEN_BS_MULTIPLE
    base_lr = LR
    lr_schedule = staticmethod(lambda it: lr_decay(it))

    def __init__(self, batch_size: int = BATCH_SIZE, seed: int | None = 0, store=None, world_size: int = 1,
                 grad_allreduce=None, bn_sync: bool = False, peer=None, grad_wire: str = "fp32"):
        """bn_sync: reduce the (conditional) batch-norm statistics of G over all ranks as well (every statistic tower
        then spans the ranks' shares: N GPUs x batch/N reproduce one GPU at the global batch).  Default: per-rank
        statistics, the reference's per-tower semantics.
        peer: a peer.PeerComm -- the statistic exchanges then are single peer-memory kernels inside the captured graphs
        (csrc/peer.cu).  Without it they go through the all-reduce callable in the middle of the passes, and the mode
        runs eagerly (capture() is a no-op)."""
        if grad_wire not in ("fp32", "bf16"):
            raise ValueError("grad_wire must be 'fp32' or 'bf16'")
        # grad_wire='bf16': the flat gradient buffers cross NVLink as bf16 (half the bytes of the all-reduce; the sum of
        # the ranks' gradients is rounded to 8 mantissa bits, below the noise of their bf16-operand computation) and come
        # back into the fp32 buffers Adam reads.  The two casts are part of the captured compute / update graphs.
        self.grad_wire = grad_wire if (world_size > 1 and grad_allreduce is not None) else "fp32"
        self._wire = {}
        self.store = store or get_store()
        self.bn_sync = bool(bn_sync) and world_size > 1
        self.bn_sync_in_graph = self.bn_sync and peer is not None
        if self.bn_sync:
            if peer is not None:
                self.store.bn_sync = peer
            elif grad_allreduce is None:
                raise ValueError("bn_sync needs a peer.PeerComm or the all-reduce callable (grad_allreduce)")
            else:
                self.store.bn_sync = (grad_allreduce, world_size)
        self.batch = batch_size
        self.gen_batch = self.gen_bs_multiple * batch_size
        self.world_size = world_size
        self.grad_allreduce = grad_allreduce  # callable(flat_grads) -> None, sums over ranks (NCCL)
        dev = self.store.device
        if seed is not None:
            np.random.seed(seed)
        self._build()
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        # static step inputs (also the CUDA-graph placeholders)
        self.real_int = torch.zeros(batch_size, self.output_dim, **i32)
        self.real_labels = torch.zeros(batch_size, **i32)
        self.deq_noise = torch.zeros(batch_size, self.output_dim, **f32)
        self.z_d = torch.zeros(batch_size, 128, **f32)
        self.z_g = torch.zeros(self.gen_batch, 128, **f32)
        self.fake_labels = torch.zeros(self.gen_batch, **i32)
        self.d_in = torch.zeros(2 * batch_size, self.output_dim, **f32)
        self.d_labels = torch.zeros(2 * batch_size, **i32)
        self.d_loss = torch.zeros(1, **f32)
        self.g_loss = torch.zeros(1, **f32)
        self.gen_opt = AdamState(self.store.flat['Generator'])
        self.disc_opt = AdamState(self.store.flat['Discriminator'])
        self._graphs = {}
        self._pair_state = None

    # ------------------------------------------------------------------------------------------ build
    def _build(self):
        """Variable creation in reference order: G tower 0, G tower 1 (reuse), D, then the
        G-step towers (reuse) -- every reuse call draws and discards its NumPy initial values."""
        st = self.store
        dev = st.device
        z = torch.zeros(2, 128, device=dev)
        lab = torch.zeros(2, dtype=torch.int32, device=dev)
        with st.building():
            fake = self.generator(2, lab, noise=z)
            self.generator(2, lab, noise=z, reuse=True)
            self.discriminator(fake, lab, update_collection="NO_OPS")
            for _ in range(self.n_towers):
                self.discriminator(self.generator(2, lab, noise=z, reuse=True), lab, update_collection="NO_OPS",
                                   reuse=True)
        st.finalize()

    # ------------------------------------------------------------------------------------------ inputs
    def set_real_batch(self, data_int, labels):
        """data_int: int32 [B, 3072] CHW-flattened pixels (host or device), labels int32 [B]."""
        self.real_int.copy_(torch.as_tensor(data_int, dtype=torch.int32), non_blocking=True)
        self.real_labels.copy_(torch.as_tensor(labels, dtype=torch.int32), non_blocking=True)

    def sample_noise(self):
        """tf.random_normal / tf.random_uniform stand-ins (gan_cifar_resnet.py:240, 335, 467)."""
        self.z_d.normal_()
        self.z_g.normal_()
        self.deq_noise.uniform_(0.0, 1.0 / 128)
        self.fake_labels.copy_((torch.rand(self.gen_batch, device=self.store.device) * self.n_classes).to(torch.int32))

    # ------------------------------------------------------------------------------------------ steps
    def _d_compute(self, after_fake=None):
        """Forward + backward of the critic step (gan_cifar_resnet.py:322-381, 524): gradients of disc_cost.
        `after_fake` (pair schedule) is called once the fake batch of this step has been issued."""
        st = self.store
        b = self.batch
        st.zero_grad('Discriminator')
        # the critic's power iteration only reads its weights and u: it runs on the side stream next to the generator pass
        # below instead of heading the critic's own chain of small kernels (~35 us of sn_fwd_* per step)
        sn_side = None
        group = st.sn_groups.get('Discriminator')
        if PREFETCH_SN and group is not None and not K.host_logic_only():
            main = torch.cuda.current_stream()
            sn_side = framework_side_stream(main.device, 2)
            sn_side.wait_stream(main)
            with torch.cuda.stream(sn_side):
                if not group.prefetch(True):
                    sn_side = None
        with st.stat_towers(self.n_towers):
            fake = self.generator(b, self.real_labels, noise=self.z_d, reuse=True)  # no tape: var_list = disc_params
        if after_fake is not None:
            after_fake()
        if sn_side is not None:
            torch.cuda.current_stream().wait_stream(sn_side)
        real = self._preprocess_real(b)
        self.d_in[:b].copy_(real)
        self.d_in[b:].copy_(fake.data)
        self.d_labels[:b].copy_(self.real_labels)
        self.d_labels[b:].copy_(self.real_labels)
        with st.gradient_tape() as tape, st.frozen_scopes('Generator'):
            disc_all, _ = self.discriminator(Var(self.d_in), self.d_labels, update_collection=None, reuse=True)
            loss = F.gan_loss(disc_all, 'hinge_d', n_real=b)
            tape.sm_split = CRITIC_BWD_SPLIT
            tape.backward(loss)
        self.d_loss.copy_(loss.data)
        self._wire_out('Discriminator')

    def _preprocess_real(self, b):
        """int pixels -> [-1, 1) + dequantisation noise, CHW -> NHWC (gan_cifar_resnet.py:334-337)."""
        return K.preprocess_real(self.real_int, self.deq_noise, b, self.image_hw)

    def _repack(self, root):
        """bf16 operand copies of the updated network, refreshed right behind its Adam step (so the compute graphs
        of later steps never re-pack a network that did not change, e.g. G during the N_CRITIC critic steps)."""
        group = self.store.pack_groups.get(root)
        if group is not None and group.entries:
            group.refresh()

    # ---- gradient exchange
    def _wire_out(self, root):
        """fp32 gradients -> the buffer that is all-reduced (end of a compute graph)."""
        if self.grad_wire == "bf16":
            g = self.store.flat[root].grads
            buf = self._wire.get(root)
            if buf is None:
                buf = self._wire[root] = torch.empty(g.numel(), dtype=torch.bfloat16, device=g.device)
            K.cast_into(g, buf)

    def _wire_in(self, root):
        """all-reduced buffer -> the fp32 gradients Adam reads (start of an update graph)."""
        if self.grad_wire == "bf16":
            K.cast_into(self._wire[root], self.store.flat[root].grads)

    def _exchange(self, root):
        if self.grad_allreduce is not None:
            self.grad_allreduce(self._wire[root] if self.grad_wire == "bf16" else self.store.flat[root].grads)

    def _d_update(self):
        self._wire_in('Discriminator')
        self.disc_opt.apply(1.0 / self.world_size)
        self.store.bump('Discriminator')
        self._repack('Discriminator')

    def _g_forward(self, tape):
        """First half of the generator step: G's forward pass on `tape`.  It reads G's parameters only, so it does not
        depend on the critic step in front of it (see _pair_body)."""
        st = self.store
        with st.resume_tape(tape), st.frozen_scopes('Discriminator'), st.stat_towers(self.n_towers):
            return self.generator(self.gen_batch, self.fake_labels, noise=self.z_g, reuse=True)

    def _g_rest(self, tape, fake):
        """Second half: D on the fake batch, gen_cost and the backward pass through D and G."""
        st = self.store
        with st.resume_tape(tape), st.frozen_scopes('Discriminator'):
            if SPLIT_CRITIC_PASS and self.gen_batch % 2 == 0 and not K.host_logic_only():
                # The critic is frozen here and has no batch statistics: its pass over the fake batch is two independent
                # half-batch chains.  They run on two streams (forward and backward), so the small, latency-bound
                # kernels of D's 16x16 / 8x8 layers overlap instead of queueing behind each other; gen_cost =
                # -mean(D(fake)) = the two half means averaged, gradients are bit-identical to the unsplit pass.
                h = self.gen_batch // 2
                halves = F.split_rows(fake, [h, h])     # recorded in front of the marker: its backward joins the branches
                marker = tape.fork()
                losses = []
                for k, blk in enumerate(halves):
                    stream = None if k == 0 else framework_aux_stream(st.device)
                    with tape.branch(marker, stream):
                        logits, _ = self.discriminator(blk, self.fake_labels[k * h:(k + 1) * h],
                                                       update_collection="NO_OPS", reuse=True)
                        losses.append(F.gan_loss(logits, 'gen', scale=0.5))
                tape.join_forward(marker)
                loss = F.add_scalars(losses[0], losses[1])
            else:
                disc_fake, _ = self.discriminator(fake, self.fake_labels, update_collection="NO_OPS", reuse=True)
                loss = F.gan_loss(disc_fake, 'gen')
            tape.backward(loss)
        self.g_loss.copy_(loss.data)
        self._wire_out('Generator')

    def _g_compute(self):
        """Forward + backward of the generator step (gan_cifar_resnet.py:462-498, 523): gradients of gen_cost."""
        st = self.store
        st.zero_grad('Generator')
        with st.gradient_tape() as tape:
            pass
        self._g_rest(tape, self._g_forward(tape))

    def _g_update(self):
        self._wire_in('Generator')
        self.gen_opt.apply(1.0 / self.world_size)
        self.store.bump('Generator')
        self._repack('Generator')

    def _d_body(self):
        self._d_compute()
        self._exchange('Discriminator')
        self._d_update()

    def _g_body(self):
        self._g_compute()
        self._exchange('Generator')
        self._g_update()

    # ------------------------------------------------------------------------------------------ D+G pair
    # One critic step followed by one generator step.  The generator step starts with G's forward pass on fresh noise,
    # which reads nothing the critic step writes: it is issued on a third stream next to the critic step, so the
    # tensor-core kernels of one pass run while the other pass is in its bandwidth-bound normalisation kernels
    # (the same idea as the filter-gradient side stream of framework.Tape, one level up).  Results are identical to
    # d_step(); g_step(): only the launch order of independent work changes.
    def _pair_fork(self, join: bool = False):
        """[aux stream] G forward of the generator step  ||  [main stream] critic forward + backward."""
        st = self.store
        with st.gradient_tape() as tape:
            pass
        if K.host_logic_only():          # CPU call-sequence tests: no streams
            st.zero_grad('Generator')
            self._pair_state = (tape, self._g_forward(tape), None)
            self._d_compute()
            return
        main = torch.cuda.current_stream()
        aux = framework_aux_stream(main.device)
        out = []

        def g_forward_on_aux():
            aux.wait_stream(main)
            with torch.cuda.stream(aux), K.sm_limit(SM_SPLIT[0] if PAIR_FORK == "late" else 0):
                st.zero_grad('Generator')
                out.append(self._g_forward(tape))
            if PAIR_FORK == "late":
                limit[0] = K.L().ganb_set_sm_limit(SM_SPLIT[1])     # the critic's kernels from here on
        limit = [None]
        if PAIR_FORK == "late":
            # the aux pass starts behind the (no-gradient) generator pass of the critic step: G's tensor-bound forward
            # then runs next to the critic's forward + backward, a chain of small latency-bound kernels
            try:
                self._d_compute(after_fake=g_forward_on_aux)
            finally:
                if limit[0] is not None:
                    K.L().ganb_set_sm_limit(limit[0])
        else:
            g_forward_on_aux()
            self._d_compute()
        fake = out[0]
        if join:          # this half is a CUDA graph of its own (a collective follows): its streams meet at its end
            main.wait_stream(aux)
            aux = None
        self._pair_state = (tape, fake, aux)

    def _pair_join(self):
        """Critic update, then the rest of the generator step behind the join with the aux stream."""
        tape, fake, aux = self._pair_state
        self._pair_state = None
        self._d_update()
        if aux is not None:
            torch.cuda.current_stream().wait_stream(aux)
        self._g_rest(tape, fake)

    def _pair_body(self):
        # a collective issued between the halves needs the streams joined, unless it is captured with them
        self._pair_fork(join=self.grad_allreduce is not None and not getattr(self, "capture_collectives", False))
        self._exchange('Discriminator')
        self._pair_join()
        self._exchange('Generator')
        self._g_update()

    def pair_step(self, iteration: int):
        """d_step(iteration) followed by g_step(iteration) as one schedule (one CUDA graph on a single GPU)."""
        self.disc_opt.set_lr(self.base_lr * self.lr_schedule(iteration))
        self.gen_opt.set_lr(self.base_lr * self.lr_schedule(iteration))
        gs = self._graphs
        if "pair_full" in gs:
            gs["pair_full"].replay()
        elif "pair_fork" in gs:
            gs["pair_fork"].replay()
            self._exchange('Discriminator')
            gs["pair_join"].replay()
            self._exchange('Generator')
            gs["g_update"].replay()
        else:
            self._pair_body()
        return self.d_loss, self.g_loss

    def _invalidate_caches(self, packs=True):
        for g in self.store.sn_groups.values():
            g.valid_for = None
            g.fresh_for = None
        if packs:
            for g in self.store.pack_groups.values():
                g.valid_for = None

    def capture(self, pair: bool = True):
        """Captures the training ops into CUDA graphs (static shapes; launch latency is first-order at batch 64).
        Must be called after at least one eager D-step and G-step (all workspaces / descriptor tables exist).
        With a gradient collective the compute and update halves are captured separately and the all-reduce
        runs between them on the same stream."""
        self.graph_launches = {}
        if self.bn_sync and not self.bn_sync_in_graph:
            return   # library collectives inside the passes: eager mode
        if self.grad_allreduce is None:
            # single GPU: nothing sits between the backward pass and the update, so each step is ONE graph (every graph
            # boundary costs ~20-30 us of idle GPU at this step size)
            parts = (("d_full", self._d_body), ("g_full", self._g_body))
        else:
            parts = (("d_compute", self._d_compute), ("d_update", self._d_update),
                     ("g_compute", self._g_compute), ("g_update", self._g_update))
        for root in ('Generator', 'Discriminator'):   # operand copies are current before any compute graph runs
            self._repack(root)
        for name, body in parts:
            # spectral-norm state is re-evaluated inside every compute graph; weight packs only behind an update
            self._invalidate_caches(packs=False)
            before = K.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            self._graphs[name] = g
            self.graph_launches[name] = K.launch_count() - before
        if pair:
            # the D+G pair schedule (pair_step): one graph on a single GPU; with a gradient collective the graphs
            # fork / join share one memory pool, because G's saved activations live across the all-reduce between them
            pool = torch.cuda.graph_pool_handle()
            one_graph = self.grad_allreduce is None or getattr(self, "capture_collectives", False)
            pparts = (("pair_full", self._pair_body),) if one_graph else (
                ("pair_fork", lambda: self._pair_fork(join=True)), ("pair_join", self._pair_join))
            self.pair_launches = 0
            for name, body in pparts:
                self._invalidate_caches(packs=False)
                before = K.launch_count()
                g = torch.cuda.CUDAGraph()
                # GANB_MAIN_PRIORITY=high: the chain of the capturing stream (critic step, backward data gradients) is
                # scheduled ahead of the side streams' kernels (kernel nodes keep their stream's priority)
                cap = torch.cuda.Stream(priority=-1) if os.environ.get("GANB_MAIN_PRIORITY") == "high" else None
                with torch.cuda.graph(g, pool=pool, stream=cap):
                    body()
                self._graphs[name] = g
                self.pair_launches += K.launch_count() - before
            if not one_graph:
                self.pair_launches += self.graph_launches["g_update"]
        self._invalidate_caches(packs=False)

    def _run(self, which):
        if (which + "_full") in self._graphs:
            self._graphs[which + "_full"].replay()
        elif (which + "_compute") in self._graphs:
            self._graphs[which + "_compute"].replay()
            self._exchange('Discriminator' if which == 'd' else 'Generator')
            self._graphs[which + "_update"].replay()
        elif which == 'd':
            self._d_body()
        else:
            self._g_body()

    def d_step(self, iteration: int):
        self.disc_opt.set_lr(self.base_lr * self.lr_schedule(iteration))
        self._run('d')
        return self.d_loss

    def g_step(self, iteration: int):
        self.gen_opt.set_lr(self.base_lr * self.lr_schedule(iteration))
        self._run('g')
        return self.g_loss

    def launches_per_pair(self) -> int:
        """libganb200 kernels in one D-step + one G-step (valid after capture(); 0 in eager mode, where the caller
        counts ganb_launch_count() around its own steps -- running extra steps here would issue collectives on one
        rank only)."""
        if getattr(self, "pair_launches", 0):
            return self.pair_launches        # the pair schedule launches the same kernels as d_full + g_full
        return sum(self.graph_launches.values())

    # ------------------------------------------------------------------------------------------ evaluation fetches
    def disc_cost_eval(self, data_int, labels):
        """session.run([disc_cost], feed_dict={all_real_data_int, all_real_labels}) -- the dev-set cost of
        gan_cifar_resnet.py:640-647: fresh noise, G forward, D forward on real + fake, no gradients.  The D call of this
        graph has update_collection=None, so -- like the reference -- every evaluation also assigns u."""
        st = self.store
        b = self.batch
        self._invalidate_caches(packs=False)    # graph replays move the weights behind the Python-side version counters
        self.set_real_batch(data_int, labels)
        self.sample_noise()
        with st.stat_towers(self.n_towers):
            fake = self.generator(b, self.real_labels, noise=self.z_d, reuse=True)
        real = self._preprocess_real(b)
        d_in = torch.cat([real.reshape(b, -1), fake.data.reshape(b, -1)], dim=0)      # tensor plumbing (no arithmetic)
        d_labels = torch.cat([self.real_labels, self.real_labels], dim=0)
        disc_all, _ = self.discriminator(Var(d_in), d_labels, update_collection=None, reuse=True)
        loss = F.gan_loss(disc_all, 'hinge_d', n_real=b).data
        self._invalidate_caches(packs=False)
        return loss

    def gen_cost_eval(self):
        """The `gen_cost` fetch that the reference adds to every critic step's session.run (:611-615): a forward pass of
        the generator-step graph on fresh noise / labels (D with NO_OPS), nothing updated."""
        st = self.store
        self._invalidate_caches(packs=False)
        self.sample_noise()
        with st.stat_towers(self.n_towers):
            fake = self.generator(self.gen_batch, self.fake_labels, noise=self.z_g, reuse=True)
        disc_fake, _ = self.discriminator(fake, self.fake_labels, update_collection="NO_OPS", reuse=True)
        loss = F.gan_loss(disc_fake, 'gen').data
        self._invalidate_caches(packs=False)
        return loss

    def fixed_samples(self, fixed_noise, fixed_labels):
        """fixed_noise_samples of :530-533: Generator(100, fixed_labels, noise=fixed_noise, reuse=True) as ONE tower."""
        out = self.generator(fixed_noise.shape[0], fixed_labels, noise=fixed_noise, reuse=True)
        return out.data

    def train_iteration(self, iteration: int, batches):
        """One reference iteration (gan_cifar_resnet.py:599-620): G-step if iteration > 0, then N_CRITIC D-steps.
        `batches` yields (data_int, labels)."""
        if iteration > 0:
            self.sample_noise()
            self.g_step(iteration)
        for _ in range(N_CRITIC):
            data, labels = next(batches)
            self.set_real_batch(data, labels)
            self.sample_noise()
            self.d_step(iteration)
        return self.d_loss, self.g_loss


def synthetic_batches(batch_size: int = BATCH_SIZE, seed: int = 0, n_batches: int = 8, output_dim: int = OUTPUT_DIM,
                      n_classes: int = 10):
    """Stand-in for common.data.cifar10.load (:557): an epoch generator factory over CHW-flattened int pixels in
    [0, 255] and int labels, shaped like the loader's batches (there is no dataset in this environment)."""
    rs = np.random.RandomState(seed)
    data = rs.randint(0, 256, size=(n_batches, batch_size, output_dim)).astype('int32')
    labels = rs.randint(0, n_classes, size=(n_batches, batch_size)).astype('int32')

    def get_epoch():
        for i in range(n_batches):
            yield data[i], labels[i]
    return get_epoch


def train(iters: int = ITERS, train_gen=None, dev_gen=None, out_dir: str = '.', checkpoint_dir: str | None = None,
          batch_size: int = BATCH_SIZE, seed: int | None = 0, capture: bool = True, fetch_gen_cost: bool = False,
          restore: bool = False, sample_every: int = 100, flush_until: int = 500, flush_every: int = 1000,
          trainer: "Trainer | None" = None, n_fixed: int = 100, fixed_labels_fn=None):
    """The training loop of the reference script (gan_cifar_resnet.py:528-660) on the B200 ops: per iteration one
    generator step (skipped at iteration 0) and N_CRITIC critic steps, `lib.plot` scalars d_cost / g_cost, every
    `sample_every` iterations the dev-set cost and a 10 x 10 sample grid from fixed noise (labels 0..9 repeated), flush +
    checkpoint while iteration < flush_until and every flush_every-th iteration.

    train_gen / dev_gen: epoch generator factories as returned by common.data.cifar10.load (default: synthetic).
    fetch_gen_cost: evaluate gen_cost inside every critic step like the reference's session.run does (:611-615; a
    redundant generator-step forward pass) instead of reporting the loss of the last generator step.
    n_fixed / fixed_labels_fn() -> (int32 labels [n_fixed], file-name tag): the fixed-noise sample batch (default: 100
    samples, labels 0..9 repeated; the ImageNet script uses 25 samples of one randomly chosen class, see
    gan_imagNet_resnet.train).  dev_gen=False skips the dev cost (commented out in the ImageNet script).
    The Inception score hook (:634-637) needs the external Inception graph and is not run.  Returns the Trainer."""
    import os

    from ..common import misc as lib_misc
    from ..common import plot as lib_plot

    tr = trainer or Trainer(batch_size=batch_size, seed=seed)
    st = tr.store
    # :530-533 -- drawn from the global NumPy stream right after graph construction
    fixed_noise = torch.from_numpy(np.random.normal(size=(n_fixed, 128)).astype('float32')).to(st.device)
    if fixed_labels_fn is None:
        fixed_labels_fn = lambda: (np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9] * 10, dtype='int32'), None)  # noqa: E731
    side = int(round((tr.output_dim // 3) ** 0.5))
    train_gen = train_gen or synthetic_batches(batch_size, seed=0, output_dim=tr.output_dim, n_classes=tr.n_classes)
    if dev_gen is None:
        dev_gen = synthetic_batches(batch_size, seed=1, n_batches=2, output_dim=tr.output_dim, n_classes=tr.n_classes)
    checkpoint_dir = checkpoint_dir or os.path.join(out_dir, 'checkpoint')
    lib_plot.set_output_dir(out_dir)
    if restore:                                                           # :590-594
        ckpts = sorted(f for f in os.listdir(checkpoint_dir) if f.startswith('model.ckpt-')) \
            if os.path.isdir(checkpoint_dir) else []
        if ckpts:
            latest = max(ckpts, key=lambda f: int(f.split('-')[1].split('.')[0]))
            print('restore model from: {}...'.format(latest))
            prefix = latest.split('.npz')[0].split('.index')[0].split('.data-')[0]     # .npz or a TF tensor bundle
            lib_misc.restore_checkpoint(os.path.join(checkpoint_dir, prefix), (tr.gen_opt, tr.disc_opt))

    def inf_train_gen():                                                  # :560-566
        while True:
            for images_, labels_ in train_gen():
                yield images_, labels_

    gen = inf_train_gen()
    captured = False
    for iteration in range(iters):
        if 0 < iteration:                                                 # :602-603
            tr.sample_noise()
            tr.g_step(iteration)
        gen_cost = tr.g_loss
        for _ in range(N_CRITIC):                                         # :605-620
            _data, _labels = next(gen)
            tr.set_real_batch(_data, _labels)
            tr.sample_noise()
            tr.d_step(iteration)
            if fetch_gen_cost:
                gen_cost = tr.gen_cost_eval()
        if capture and not captured and iteration >= 1:                   # one eager G-step and D-step have run
            tr.capture()
            captured = True
        lib_plot.plot('d_cost', tr.d_loss)                                # :625-626
        lib_plot.plot('g_cost', gen_cost)
        if iteration % sample_every == sample_every - 1:                  # :639-649
            if dev_gen:
                dev_disc_costs = [tr.disc_cost_eval(images, _labels).clone() for images, _labels in dev_gen()]
                lib_plot.plot('dev_cost', torch.stack(dev_disc_costs).mean())
            labels_np, tag = fixed_labels_fn()
            fixed_labels = torch.from_numpy(np.asarray(labels_np, dtype='int32')).to(st.device)
            samples = tr.fixed_samples(fixed_noise, fixed_labels)         # generate_image, :536-539
            name = 'samples_{}.png'.format(iteration) if tag is None else 'samples_{}_{}.png'.format(iteration, tag)
            lib_misc.save_images(samples.reshape(n_fixed, side, side, 3), os.path.join(out_dir, name))
        if (iteration < flush_until) or (iteration % flush_every == flush_every - 1):   # :651-656
            lib_plot.flush()
            if not os.path.exists(checkpoint_dir):
                os.mkdir(checkpoint_dir)
            lib_misc.save_checkpoint(os.path.join(checkpoint_dir, 'model.ckpt-{}'.format(iteration)),
                                     (tr.gen_opt, tr.disc_opt))
        lib_plot.tick()
    return tr
