"""Host-side state that the reference keeps inside TensorFlow: scope-named variables, reuse, the gradient tape,
spectral-norm `u` state and flat optimiser buffers.

Reference behaviour mirrored here
  * tf.variable_scope(name, reuse=) / tf.get_variable(name, ...) naming: "Discriminator/D.Block.2.Conv1/Filters"
    (common/ops/conv2d.py:59,142,213; linear.py:45,140,177; sn.py:28,32; normalization.py:43,49,51;
    embedding.py:28,40) and tf.trainable_variables() filtered by substring (SNGAN/gan_cifar_resnet.py:507,512);
  * NumPy initial values are drawn on EVERY layer call while the graph is being built, also when the variable
    already exists (conv2d.py:124-144) -- `building()` turns that on so the RNG stream matches the reference;
  * tf.gradients(...) -> Tape.backward(); there is no torch.autograd on this path.

torch supplies device memory and streams; all arithmetic runs in libganb200 kernels (kernels.py).
"""
from __future__ import annotations

import contextlib
import os
from collections import OrderedDict

import numpy as np
import torch

from . import kernels as K

ACT_GRAD_BF16 = True  # gradients of bf16 activations are stored in bf16
SMALL_K = 32  # narrowest im2col width (bf16 columns) of the <=8-channel tensor-core route


def small_k(taps: int, c: int):
    """im2col width for a <=8-channel operand: taps*c columns padded to 32 / 64 / 128 (None: does not fit)."""
    for k in (32, 64, 128):
        if taps * c <= k:
            return k
    return None
_ALIGN = 64  # floats: every variable starts on a 256-byte boundary inside the flat buffers


class Var:
    """A value on the tape: `data` is a device tensor, `grad` is filled by Tape.backward()."""

    __slots__ = ("data", "grad", "requires_grad", "grad_dtype", "quad", "stats", "fused_act")

    def __init__(self, data: torch.Tensor, requires_grad: bool = False, grad_dtype=None):
        self.data = data
        self.quad = False   # storage is the quad layout of functional.upconv2d (value AND gradient)
        self.stats = None   # kernels.FusedStats left by the producing convolution's epilogue (functional.conv2d)
        # activation that the producing convolution applied in its epilogue (functional.conv2d(act=...)): `grad` is then
        # the gradient wrt the PRE-activation, which only a consumer that applies act' itself may deliver (accum(gated=True))
        self.fused_act = None
        self.grad = None
        self.requires_grad = requires_grad
        self.grad_dtype = grad_dtype  # dtype producers should use for this value's gradient (None: fp32)

    @property
    def gdtype(self):
        # A gradient has the dtype of its value unless a producer/consumer pair asked otherwise: bf16 tensor-core
        # operands get bf16 gradients (half the traffic of the data-gradient write and of both backward reads;
        # measured to be invisible next to the ReLU-mask sensitivity of a bf16 forward, DESIGN.md "Parity").
        if self.grad_dtype is not None:
            return self.grad_dtype
        return self.data.dtype if ACT_GRAD_BF16 else torch.float32

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def dtype(self):
        return self.data.dtype

    def accum(self, g: torch.Tensor, gated: bool = False) -> None:
        """Adds a gradient contribution (takes ownership of `g` when it is the first one)."""
        if (self.fused_act is not None) != gated:
            raise RuntimeError("gradient of a value with an epilogue-fused activation must come from a consumer that "
                               "applies the activation's derivative (functional.conv2d), and only from such a consumer")
        if self.grad is None:
            self.grad = g
        else:
            if g.dtype != torch.float32 or self.grad.dtype != torch.float32:
                a, b = K.cast(self.grad, torch.float32), K.cast(g, torch.float32)
                K.axpby(b, a, 1.0, 1.0)
                self.grad = K.cast(a, self.grad.dtype) if self.grad.dtype != torch.float32 else a
            else:
                K.axpby(g, self.grad, 1.0, 1.0)

    def numpy(self):
        return self.data.float().cpu().numpy()


class Variable(Var):
    """A named, persistent tensor (tf.Variable)."""

    __slots__ = ("key", "trainable", "root", "store")

    def __init__(self, key, data, trainable, store):
        super().__init__(data, requires_grad=trainable)
        self.key = key
        self.trainable = trainable
        self.root = key.split("/")[0]
        self.store = store

    @property
    def name(self):
        return self.key + ":0"

    @property
    def needs_grad(self):
        return self.trainable and self.root not in self.store.frozen

    def add_grad(self, writer) -> None:
        """`writer(dst, beta)` must compute dst = beta*dst + contribution. Gradients of variables accumulate."""
        if self.grad is None:
            self.grad = torch.zeros_like(self.data)
        writer(self.grad, 1.0)


class DerivedWeight(Variable):
    """Filters after weight-norm and / or a constant mask: W_eff = W * (g / ||W||) * mask (common/ops/conv2d.py:153-167,
    linear.py:143-155, deconv2d.py:87-96).  Layers (and the spectral-norm / operand-pack groups) see it as the weight;
    `data` is recomputed when the network's version moves, `grad` collects dL/dW_eff during a backward pass and
    Tape.backward maps it onto W.grad / g.grad at the end (after the spectral-norm backward, which also adds into
    `grad`).  geom = (a, c, b) of include/ganb200.h."""

    __slots__ = ("src", "gvar", "mask", "geom", "norms", "valid_for", "tape_token")

    def __init__(self, src: Variable, gvar, mask, geom):
        super().__init__(src.key + ":effective", torch.empty_like(src.data), False, src.store)
        self.root = src.root
        self.src, self.gvar, self.mask, self.geom = src, gvar, mask, geom
        self.norms = torch.empty(geom[1], dtype=torch.float32, device=src.data.device) if gvar is not None else None
        self.grad = torch.zeros_like(self.data)
        self.valid_for = None
        self.tape_token = None

    @property
    def needs_grad(self):
        return self.src.needs_grad

    def refresh(self) -> None:
        ver = self.store.version(self.root)
        if self.valid_for == ver:
            return
        K.weight_transform_fwd(self.src.data, None if self.gvar is None else self.gvar.data, self.mask, self.data,
                               self.norms, *self.geom)
        self.valid_for = ver

    def attach(self) -> None:
        """Called by the layer on every use: on a recording tape the first use clears `grad` and books the backward map."""
        st = self.store
        if st.tape is None or not self.needs_grad or self.tape_token == st.tape_token:
            return
        self.tape_token = st.tape_token
        self.grad.zero_()
        st.tape.pending_derived.append(self)

    def backward(self) -> None:
        for v in (self.src, self.gvar):
            if v is not None and v.grad is None:
                v.grad = torch.zeros_like(v.data)
        K.weight_transform_bwd(self.src.data, self.grad, None if self.gvar is None else self.gvar.data, self.mask,
                               self.norms, self.src.grad, None if self.gvar is None else self.gvar.grad, *self.geom)


SIDE_STREAM = os.environ.get("GANB_SIDE_STREAM", "1") != "0"
# bias gradients on a side stream of their own (lane 1): measured SLOWER (3.14 vs 3.06-3.11 ms per pair: the extra
# concurrency takes bandwidth from the data-gradient chain), so they stay queued behind the filter gradients; opt-in
BIAS_LANE = os.environ.get("GANB_BIAS_LANE", "0") == "1"
_side_streams = {}


def _side_stream(device, lane: int = 0) -> "torch.cuda.Stream":
    key = (torch.device(device).index or 0, lane)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


_aux_streams = {}


def aux_stream(device) -> "torch.cuda.Stream":
    """Third stream: a whole independent pass (the generator forward of the G-step, which does not depend on the critic
    update) runs here next to the critic step of the main stream; see SNGAN.gan_cifar_resnet.Trainer._pair_body."""
    key = torch.device(device).index or 0
    if key not in _aux_streams:
        _aux_streams[key] = torch.cuda.Stream(device=device)
    return _aux_streams[key]


class Tape:
    """Records backward closures in execution order; backward() replays them in reverse.

    Work that is off the critical chain of the backward pass (filter gradients, bias gradients, their split
    reductions) is issued on a second stream so that these launches overlap the data-gradient / normalisation chain:
    a tensor-core kernel of one stream runs next to the bandwidth-bound kernels of the other.  Under CUDA-graph
    capture the fork/join events become graph edges."""

    def __init__(self, store):
        self.nodes = []
        self.store = store
        self.pending_sn = OrderedDict()  # root -> list of SN entries whose G buffer has been written
        self.pending_derived = []        # DerivedWeight instances used on this tape (weight-norm / masks)
        self.keep = []       # temporaries read by side-stream launches: kept alive until the join
        self._forked = set()   # side-stream lanes with work in flight
        # (main, side) SM budgets of the tensor-core kernels during backward(): with small layers (the critic) a data-
        # gradient kernel and a filter-gradient kernel that each take every SM run one after the other; two half-GPU
        # kernels run side by side and lose little, because such launches are dominated by fixed latencies.  None: off
        self.sm_split = None
        self.token = 0       # identity of this tape for the spectral-norm evaluation bookkeeping (VariableStore)
        self.sn_gen = {}     # root -> index of the spectral-norm state set in use on this tape
        self.node_stream = {}   # node index -> stream of the branch it was recorded in
        self.branches = []      # (fork marker, stream)
        self._branch = None

    def record(self, fn) -> None:
        self.nodes.append(fn)
        if self._branch is not None:
            self.node_stream[len(self.nodes) - 1] = self._branch

    # ---- branches: independent sub-graphs of ONE pass issued on their own streams (forward and backward)
    def fork(self) -> int:
        """Marks the point where branches start; returns the marker that branch() / join_forward() take."""
        return len(self.nodes)

    @contextlib.contextmanager
    def branch(self, marker: int, stream: "torch.cuda.Stream | None"):
        """Ops recorded inside run on `stream` (None: the current stream, i.e. a branch that stays on the main chain), in
        the forward pass and -- their backward closures -- in the backward pass.  The branch must only consume values
        produced before fork() and must not share gradient accumulators with other branches; Tape.backward makes the
        stream wait for the upstream gradients and makes the nodes in front of the marker wait for the branch."""
        if stream is None or K.host_logic_only():
            yield
            return
        main = torch.cuda.current_stream()
        stream.wait_stream(main)
        self._branch = stream
        self.branches.append((marker, stream))
        try:
            with torch.cuda.stream(stream):
                yield
        finally:
            self._branch = None

    def join_forward(self, marker: int) -> None:
        """The current stream waits for the forward work of every branch forked at `marker`."""
        main = torch.cuda.current_stream()
        for m, s in self.branches:
            if m == marker:
                main.wait_stream(s)

    @contextlib.contextmanager
    def offchain(self, lane: int = 0):
        """Launches issued inside run on a side stream, ordered after everything queued on the current stream.
        lane 0: filter gradients (tensor-core kernels + their split reductions); lane 1: bias gradients (small
        bandwidth-bound column sums, which would otherwise queue between the filter-gradient kernels of lane 0)."""
        if not SIDE_STREAM or K.host_logic_only():
            yield
            return
        if lane and not BIAS_LANE:
            lane = 0
        main = torch.cuda.current_stream()
        side = _side_stream(main.device, lane)
        side.wait_stream(main)
        self._forked.add(lane)
        with torch.cuda.stream(side), K.sm_limit(self.sm_split[1] if self.sm_split else 0):
            yield

    def join(self) -> None:
        if self._forked:
            main = torch.cuda.current_stream()
            for lane in sorted(self._forked):
                main.wait_stream(_side_stream(main.device, lane))
            self._forked.clear()
        self.keep.clear()

    def backward(self, loss: Var, grad: torch.Tensor | None = None) -> None:
        if grad is not None:
            loss.accum(grad)
        if not self.node_stream:
            with K.sm_limit(self.sm_split[0] if self.sm_split else 0):
                for fn in reversed(self.nodes):
                    fn()
        else:
            main = torch.cuda.current_stream()
            entered, pending = set(), []          # branch streams already running backward work / not yet joined
            for i in range(len(self.nodes) - 1, -1, -1):
                s = self.node_stream.get(i)
                if pending:                       # nodes in front of a fork marker consume what its branches produced
                    for m, bs in list(pending):
                        if i < m:
                            main.wait_stream(bs)
                            pending.remove((m, bs))
                if s is None:
                    self.nodes[i]()
                    continue
                if s not in entered:              # the branch needs the gradients the main chain has produced so far
                    s.wait_stream(main)
                    entered.add(s)
                    pending.extend((m, bs) for m, bs in self.branches if bs is s)
                with torch.cuda.stream(s):
                    self.nodes[i]()
            for _m, bs in pending:
                main.wait_stream(bs)
            self.node_stream.clear()
            self.branches.clear()
        self.join()
        self.nodes.clear()
        for root, entries in self.pending_sn.items():
            # one launch sequence per evaluation of the network: a weight evaluated twice on this tape (D(real) and
            # D(fake) as separate calls) has one state per evaluation, and both add into the same dW
            by_group = OrderedDict()
            for e in entries:
                by_group.setdefault(id(e.group), (e.group, []))[1].append(e)
            for group, es in by_group.values():
                group.backward(es)
        self.pending_sn.clear()
        for d in self.pending_derived:   # dL/dW_eff (incl. the spectral-norm part) -> dL/dW, dL/dg
            d.backward()
            d.tape_token = None
        self.pending_derived.clear()


class FlatGroup:
    """All trainable variables of one network in flat params / grads / Adam-slot buffers."""

    def __init__(self, variables, device):
        self.variables = variables
        off = 0
        self.offsets = []
        for v in variables:
            self.offsets.append(off)
            off += -(-v.data.numel() // _ALIGN) * _ALIGN
        self.size = off
        self.params = torch.zeros(off, dtype=torch.float32, device=device)
        self.grads = torch.zeros(off, dtype=torch.float32, device=device)
        self.m = torch.zeros(off, dtype=torch.float32, device=device)
        self.v = torch.zeros(off, dtype=torch.float32, device=device)
        for v, o in zip(variables, self.offsets):
            n = v.data.numel()
            self.params[o:o + n].copy_(v.data.reshape(-1))
            v.data = self.params[o:o + n].view(v.data.shape)
            v.grad = self.grads[o:o + n].view(v.data.shape)

    def zero_grad(self):
        self.grads.zero_()


class SNEntry:
    """Per-weight spectral-norm state: u (persistent), the vectors of the last evaluation and the G buffer."""

    def __init__(self, w: Variable, u: Variable):
        dev = w.data.device
        self.w, self.u = w, u
        self.c = w.data.shape[-1]
        self.k = w.data.numel() // self.c
        f32 = dict(dtype=torch.float32, device=dev)
        self.u_out = torch.zeros(self.c, **f32)
        self.u_used = torch.zeros(self.c, **f32)
        self.v = torch.zeros(self.k, **f32)
        self.b = torch.zeros(self.c, **f32)
        self.scal = torch.zeros(8, **f32)            # sigma, 1/sigma, |Wu|, |b|, bwd scratch
        self.nblk = -(-self.k // K.SN_ROWS)
        self.t = torch.zeros(self.k, **f32)
        self.work = torch.zeros(self.nblk * (self.c + 4), **f32)
        self.g = torch.zeros(self.k, self.c, **f32)  # dL/d(W/sigma), written by the layer's wgrad
        self.g_written = False
        self.fresh = False
        self.group = None   # the SNGroup that owns this evaluation state

    @property
    def inv_sigma(self):
        return self.scal[1:2]

    @property
    def sigma(self):
        return self.scal[0:1]


class SNGroup:
    """All spectrally-normalised weights of one network: evaluated by ONE grouped launch per weight version."""

    def __init__(self, store, root):
        self.store, self.root = store, root
        self.entries: "OrderedDict[str, SNEntry]" = OrderedDict()
        self.table = None
        self.table_ptrs = None
        self.valid_for = None
        self.fresh_for = None
        self.used_token = None   # tape on which the current evaluation has consumers (see VariableStore.sn_group)

    def entry(self, w: Variable, u: Variable) -> SNEntry:
        e = self.entries.get(w.key)
        if e is None:
            e = SNEntry(w, u)
            e.group = self
            self.entries[w.key] = e
            self.table = None
            self.valid_for = None
            self.fresh_for = None
        return e

    def _build_table(self, entries):
        items, blocks = [], 0
        for e in entries:
            if e.w.grad is None:
                e.w.grad = torch.zeros_like(e.w.data)
            items.append(K.SnLayerStruct(e.w.data.data_ptr(), e.u.data.data_ptr(), e.u_out.data_ptr(),
                                         e.u_used.data_ptr(), e.v.data_ptr(), e.b.data_ptr(), e.scal.data_ptr(),
                                         e.g.data_ptr(), e.w.grad.data_ptr(), e.t.data_ptr(), e.work.data_ptr(),
                                         e.k, e.c, blocks, 0))
            blocks += e.nblk
        dev = entries[0].w.data.device
        return K.struct_array_to_device(items, dev), blocks, max(e.c for e in entries)

    def _ptrs(self):
        return tuple((e.w.data.data_ptr(), e.u.data.data_ptr(), 0 if e.w.grad is None else e.w.grad.data_ptr())
                     for e in self.entries.values())

    def _run(self, assign: bool) -> None:
        entries = list(self.entries.values())
        if self.table is None or self.table_ptrs != self._ptrs():
            self.table, self.blocks, self.max_c = self._build_table(entries)
            self.table_ptrs = self._ptrs()
        K.sn_power_iter(self.table, len(entries), self.blocks, self.max_c, assign)

    def needs_new_evaluation(self, e: SNEntry, assign: bool) -> bool:
        """True when acquire(e, assign) would run the power iteration again (and overwrite sigma / v / u_used)."""
        ver = self.store.version(self.root)
        if assign:
            return not (e.fresh and self.fresh_for == ver)
        return self.valid_for != (ver, self.store.u_version(self.root))

    def acquire(self, e: SNEntry, assign: bool) -> None:
        """Makes e.scal / e.v / e.u_used current for this evaluation of the network.

        assign=True  (update_collection=None): ONE power iteration per forward pass of the network, u <- u'.
                     The first layer that finds its entry already consumed starts a new pass for all layers.
        assign=False (NO_OPS): a fresh iteration from the stored u, cached until weights or u change."""
        ver = self.store.version(self.root)
        if assign:
            if not (e.fresh and self.fresh_for == ver):
                self._run(True)
                self.store.bump_u(self.root)
                self.valid_for = None
                self.fresh_for = ver
                for x in self.entries.values():
                    x.fresh = True
            e.fresh = False
        else:
            key = (ver, self.store.u_version(self.root))
            if self.valid_for != key:
                self._run(False)
                self.valid_for = key
                for x in self.entries.values():
                    x.fresh = False

    def prefetch(self, assign: bool) -> bool:
        """Runs the evaluation that the first acquire() of the coming pass would start, ahead of that pass (on whatever
        stream is current: the power iteration reads the weights and u only, so it can sit next to unrelated work).
        The layers' acquire() calls then find their state current.  Returns False when nothing had to run."""
        if not self.entries:
            return False
        ver = self.store.version(self.root)
        if assign:
            if self.fresh_for == ver and all(e.fresh for e in self.entries.values()):
                return False
            self._run(True)
            self.store.bump_u(self.root)
            self.valid_for = None
            self.fresh_for = ver
            for x in self.entries.values():
                x.fresh = True
            return True
        key = (ver, self.store.u_version(self.root))
        if self.valid_for == key:
            return False
        self._run(False)
        self.valid_for = key
        for x in self.entries.values():
            x.fresh = False
        return True

    def backward(self, entries) -> None:
        all_entries = list(self.entries.values())
        if len(entries) == len(all_entries) and self.table is not None and self.table_ptrs == self._ptrs():
            K.sn_bwd(self.table, len(all_entries), self.blocks, self.max_c)
        else:
            table, blocks, mc = self._build_table(entries)
            K.sn_bwd(table, len(entries), blocks, mc)
        for e in entries:
            e.g_written = False


class PackEntry:
    def __init__(self, w: Variable):
        shape = w.data.shape
        self.w = w
        self.co = shape[-1]
        self.ci = shape[-2]
        self.taps = w.data.numel() // (self.ci * self.co)
        dev = w.data.device
        # input channels of the operand copies: ragged counts above 8 (513 behind minibatch_std) are zero-padded to
        # the next multiple of 8 so that every row meets the 16-byte granularity of the TMA path
        self.ci_pad = self.ci if (self.ci < 8 or self.ci % 8 == 0) else (self.ci + 7) // 8 * 8
        alloc = torch.zeros if self.ci_pad != self.ci else torch.empty
        self.wn = alloc(self.taps, self.ci_pad, self.co, dtype=torch.bfloat16, device=dev)  # [tap][ci][co]
        self.wt = alloc(self.taps, self.co, self.ci_pad, dtype=torch.bfloat16, device=dev)  # [tap][co][ci]
        # <=8-channel side: [large channel][tap*cs + c] padded to SMALL_K columns (operand of the im2col route)
        self.small = None
        self.ws = None
        self.kpad = None
        if self.ci < 8 and small_k(self.taps, self.ci) is not None:
            self.small = "ci"
            self.kpad = small_k(self.taps, self.ci)
            self.ws = torch.empty(self.co, self.kpad, dtype=torch.bfloat16, device=dev)
        elif self.co < 8 and small_k(self.taps, self.co) is not None:
            self.small = "co"
            self.kpad = small_k(self.taps, self.co)
            self.ws = torch.empty(self.ci, self.kpad, dtype=torch.bfloat16, device=dev)
        # effective 2x2 filters of the sub-pixel UpsampleConv (functional.upconv2d), allocated on first use
        self.we_t = None   # [16][co][ci]
        self.we_n = None   # [16][ci][co]

    def enable_upconv(self) -> bool:
        """Allocates the sub-pixel operand copies; returns True when they were just created (the group must re-pack)."""
        if self.we_t is not None:
            return False
        dev = self.w.data.device
        self.we_t = torch.empty(16, self.co, self.ci, dtype=torch.bfloat16, device=dev)
        self.we_n = torch.empty(16, self.ci, self.co, dtype=torch.bfloat16, device=dev)
        return True


class PackGroup:
    """bf16 operand copies of every tensor-core weight of one network, refreshed by one grouped launch."""

    def __init__(self, store, root):
        self.store, self.root = store, root
        self.entries: "OrderedDict[str, PackEntry]" = OrderedDict()
        self.table = None
        self.table_ptrs = None
        self.valid_for = None

    def entry(self, w: Variable) -> PackEntry:
        e = self.entries.get(w.key)
        if e is None:
            e = PackEntry(w)
            self.entries[w.key] = e
            self.table = None
            self.valid_for = None
        return e

    def _ptrs(self):
        return tuple(e.w.data.data_ptr() for e in self.entries.values())

    def refresh(self) -> None:
        ver = self.store.version(self.root)
        if self.valid_for == ver:
            return
        entries = list(self.entries.values())
        for e in entries:
            if isinstance(e.w, DerivedWeight):
                e.w.refresh()        # the operand copies are taken from the effective filters
        if self.table is None or self.table_ptrs != self._ptrs():
            items, tiles = [], 0
            for e in entries:
                items.append(K.PackLayerStruct(e.w.data.data_ptr(), e.wn.data_ptr(), e.wt.data_ptr(), e.taps, e.ci,
                                               e.co, tiles, e.ci_pad, 0))
                tiles += e.taps * (-(-e.ci // 64)) * (-(-e.co // 64))     # ganb_pack_layer.tile_begin
            self.table = K.struct_array_to_device(items, entries[0].w.data.device)
            self.total_tiles = tiles
            self.table_ptrs = self._ptrs()
        K.pack_weights(self.table, len(entries), self.total_tiles)
        for e in entries:
            if e.small is not None:
                K.pack_small(e.w.data, e.ws, e.taps, e.ci, e.co, e.small == "ci", e.kpad)
            if e.we_t is not None:
                K.upconv_pack(e.w.data, e.we_t, e.we_n, e.ci, e.co)
        self.valid_for = ver


class VariableStore:
    def __init__(self, device="cuda", u_seed: int = 2):
        self.device = torch.device(device)
        self.vars: "OrderedDict[str, Variable]" = OrderedDict()
        self._scopes: list[tuple[str, bool | None]] = []
        self.draw_on_reuse = False
        self.frozen: set[str] = set()
        self._versions: dict[str, int] = {}
        self._u_versions: dict[str, int] = {}
        self.sn_groups: dict[str, SNGroup] = {}
        # further evaluations of the same network on ONE tape (PGGAN / Pix2Pix call D on real and fake images
        # separately, with different u): every evaluation keeps its own sigma / v / u' until the backward pass
        self.sn_shadow: dict[str, list] = {}
        self._sn_gen: dict[str, int] = {}
        self.tape_token = 0
        self._token_seq = 0
        self.pack_groups: dict[str, PackGroup] = {}
        self.derived: dict[str, DerivedWeight] = {}
        self.flat: dict[str, FlatGroup] = {}
        self.u_rng = np.random.RandomState(u_seed)
        self.tape: Tape | None = None
        self.stat_groups = 1
        self._consts: dict[float, torch.Tensor] = {}
        # cross-GPU batch statistics: (allreduce_sum(tensor) -> None, world size) or None = per-rank statistics, the
        # reference's per-tower semantics (common/ops/normalization.py:47)
        self.bn_sync = None
        # operand precision of the tensor-core layers per network (root scope): 'bf16' (default) or 'tf32' -- fp32
        # activations and filters read by kind::tf32 MMAs (north star: "BF16 or TF32 inputs"; <= 1e-3 per layer)
        self.precision: dict[str, str] = {}
        self._tf32_packs: dict[str, tuple] = {}

    # -- scopes ---------------------------------------------------------------------------------
    @contextlib.contextmanager
    def variable_scope(self, name, reuse=None):
        self._scopes.append((name, reuse))
        try:
            yield
        finally:
            self._scopes.pop()

    def scope_name(self) -> str:
        return "/".join(s for s, _ in self._scopes if s)

    def _reuse(self) -> bool:
        return any(r for _, r in self._scopes)

    def root(self) -> str:
        for s, _ in self._scopes:
            if s:
                return s
        return ""

    # -- variables ------------------------------------------------------------------------------
    def get_variable(self, name, shape=None, initializer=None, trainable=True) -> Variable:
        key = "/".join([s for s, _ in self._scopes if s] + [name])
        v = self.vars.get(key)
        if v is not None:
            if self.draw_on_reuse and callable(initializer):
                initializer(shape)  # the reference draws and discards on every reuse call
            return v
        if self._reuse():
            raise ValueError(f"Variable {key} does not exist, but reuse=True was requested")
        if initializer is None:
            raise ValueError(f"Variable {key} needs an initializer")
        value = initializer(shape) if callable(initializer) else initializer
        arr = np.ascontiguousarray(np.asarray(value, dtype=np.float32))
        if shape is not None and tuple(arr.shape) != tuple(shape):
            arr = np.broadcast_to(arr, shape).copy()
        data = torch.from_numpy(arr).to(self.device)
        v = Variable(key, data, bool(trainable), self)
        self.vars[key] = v
        return v

    def trainable_variables(self, substring: str = ""):
        return [v for v in self.vars.values() if v.trainable and substring in v.name]

    def global_variables(self):
        return list(self.vars.values())

    # -- build / finalize -----------------------------------------------------------------------
    @contextlib.contextmanager
    def building(self):
        """Graph-construction phase: reuse calls draw (and discard) their NumPy initial values."""
        self.draw_on_reuse = True
        try:
            yield
        finally:
            self.draw_on_reuse = False

    def finalize(self) -> None:
        """Moves the trainable variables of every root scope into flat buffers (params / grads / Adam slots)."""
        roots = OrderedDict()
        for v in self.vars.values():
            if v.trainable and v.root not in self.flat:
                roots.setdefault(v.root, []).append(v)
        for root, variables in roots.items():
            self.flat[root] = FlatGroup(variables, self.device)
            self.bump(root)

    # -- versions -------------------------------------------------------------------------------
    def version(self, root):
        return self._versions.get(root, 0)

    def bump(self, root):
        self._versions[root] = self._versions.get(root, 0) + 1

    def u_version(self, root):
        return self._u_versions.get(root, 0)

    def bump_u(self, root):
        self._u_versions[root] = self._u_versions.get(root, 0) + 1

    # -- operand precision ----------------------------------------------------------------------
    def set_precision(self, root: str, precision: str) -> None:
        if precision not in ("bf16", "tf32"):
            raise ValueError("precision must be 'bf16' or 'tf32'")
        self.precision[root] = precision

    def is_tf32(self, root: str | None = None) -> bool:
        return self.precision.get(self.root() if root is None else root, "bf16") == "tf32"

    def tf32_operands(self, w: "Variable"):
        """(fprop copy [taps, cout, cin], dgrad copy = the HWIO filter [taps, cin, cout]) of `w`, rounded to TF32;
        refreshed when the network's version moves."""
        if isinstance(w, DerivedWeight):
            w.refresh()
        ver = self.version(w.root)
        hit = self._tf32_packs.get(w.key)
        if hit is not None and hit[0] == ver and hit[3] == w.data.data_ptr():
            return hit[1], hit[2]
        co, ci = w.data.shape[-1], w.data.shape[-2]
        taps = w.data.numel() // (ci * co)
        flat = w.data.reshape(taps, ci, co)
        wt = K.transpose_tf32(flat, taps, ci, co)
        wr = K.round_tf32(flat)
        self._tf32_packs[w.key] = (ver, wt, wr, w.data.data_ptr())
        return wt, wr

    # -- helpers --------------------------------------------------------------------------------
    def const(self, value: float) -> torch.Tensor:
        """Device-resident fp32 scalar (the `alpha` of a GEMM epilogue that is not a 1/sigma), cached by value."""
        value = float(value)
        t = self._consts.get(value)
        if t is None:
            t = torch.full((1,), value, dtype=torch.float32, device=self.device)
            self._consts[value] = t
        return t

    def sn_group(self, root, gen: int = 0) -> SNGroup:
        g = self.sn_groups.get(root)
        if g is None:
            g = self.sn_groups[root] = SNGroup(self, root)
        if gen == 0:
            return g
        shadows = self.sn_shadow.setdefault(root, [])
        while len(shadows) < gen:
            sh = SNGroup(self, root)
            for e in g.entries.values():     # same layers, separate evaluation state
                sh.entry(e.w, e.u)
            shadows.append(sh)
        sh = shadows[gen - 1]
        for e in g.entries.values():
            if e.w.key not in sh.entries:
                sh.entry(e.w, e.u)
        return sh

    def sn_acquire(self, w: Variable, u: Variable, assign: bool) -> SNEntry:
        """Evaluation state of W / sigma for the layer call being built.  While a tape is recording, a network whose
        current state already has consumers on that tape and would be re-evaluated (changed u, or another
        update_collection=None pass) moves on to a fresh state set instead of overwriting the one in use."""
        root = w.root
        gen = self._sn_gen.get(root, 0) if self.tape is not None else 0
        group = self.sn_group(root, gen)
        entry = group.entry(w, u)
        if (self.tape is not None and group.used_token == self.tape_token
                and group.needs_new_evaluation(entry, assign)):
            gen += 1
            self._sn_gen[root] = gen
            group = self.sn_group(root, gen)
            entry = group.entry(w, u)
        group.acquire(entry, assign)
        if self.tape is not None:
            group.used_token = self.tape_token
        return entry

    def effective_weight(self, w: Variable, gvar=None, mask=None, geom=None) -> Variable:
        """The weight a layer multiplies with: `w` itself, or its DerivedWeight when weight-norm (gvar) and / or a
        constant mask apply.  Current on return; registered with the recording tape."""
        if gvar is None and mask is None:
            return w
        d = self.derived.get(w.key)
        if d is None:
            if callable(mask):
                mask = mask()
            if mask is not None and not torch.is_tensor(mask):
                mask = torch.from_numpy(np.ascontiguousarray(mask, dtype=np.float32)).to(self.device)
            d = self.derived[w.key] = DerivedWeight(w, gvar, mask, geom)
        d.refresh()
        d.attach()
        return d

    def pack_group(self, root) -> PackGroup:
        g = self.pack_groups.get(root)
        if g is None:
            g = self.pack_groups[root] = PackGroup(self, root)
        return g

    @contextlib.contextmanager
    def gradient_tape(self):
        self._token_seq += 1
        tape = Tape(self)
        tape.token = self._token_seq
        with self.resume_tape(tape):
            yield tape

    @contextlib.contextmanager
    def resume_tape(self, tape: Tape):
        """Makes `tape` the recording tape again: a forward pass may be recorded in pieces with other tapes (another
        network's whole step) in between -- the generator forward of the G-step is issued next to the critic step."""
        prev = (self.tape, self.tape_token, self._sn_gen)
        self.tape, self.tape_token, self._sn_gen = tape, tape.token, tape.sn_gen
        try:
            yield tape
        finally:
            self.tape, self.tape_token, self._sn_gen = prev

    @contextlib.contextmanager
    def frozen_scopes(self, *roots):
        """Variables under these root scopes receive no gradient (var_list of the other optimiser)."""
        prev = set(self.frozen)
        self.frozen |= set(roots)
        try:
            yield
        finally:
            self.frozen = prev

    @contextlib.contextmanager
    def stat_towers(self, groups: int):
        """Batch-statistic groups: the reference builds `groups` towers that each normalise their own
        slice of the batch (SNGAN/gan_cifar_resnet.py:326-332, 464-482)."""
        prev = self.stat_groups
        self.stat_groups = groups
        try:
            yield
        finally:
            self.stat_groups = prev

    def zero_grad(self, root=None):
        for r, f in self.flat.items():
            if root is None or r == root:
                f.zero_grad()

    def state_dict(self):
        return {k: v.data.detach().cpu().numpy().copy() for k, v in self.vars.items()}

    def load_state_dict(self, state, strict=False):
        """optimistic_restore semantics (common/misc.py:275-307): only name AND shape matches are restored."""
        restored = []
        for k, arr in state.items():
            v = self.vars.get(k)
            if v is not None and tuple(v.data.shape) == tuple(arr.shape):
                v.data.copy_(torch.from_numpy(np.asarray(arr, dtype=np.float32)).to(self.device))
                restored.append(k)
            elif strict:
                raise KeyError(k)
        for root in {k.split("/")[0] for k in restored}:
            self.bump(root)
            self.bump_u(root)
        return restored


def truncated_normal(shape, rng: np.random.RandomState):
    """tf.truncated_normal_initializer(): standard normal, values beyond 2 sigma re-drawn (sn.py:32)."""
    out = rng.standard_normal(size=shape)
    bad = np.abs(out) > 2
    while bad.any():
        out[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(out) > 2
    return out.astype("float32")


_default_store: VariableStore | None = None


def get_store() -> VariableStore:
    global _default_store
    if _default_store is None:
        if K.host_logic_only():
            _default_store = VariableStore("cpu")
        elif not torch.cuda.is_available():
            raise RuntimeError("gan_lib_tensorflow_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        else:
            _default_store = VariableStore("cuda")
    return _default_store


def set_store(store: VariableStore | None) -> None:
    global _default_store
    _default_store = store


def reset_default_graph(device="cuda", u_seed: int = 2) -> VariableStore:
    """tf.reset_default_graph(): drops every variable."""
    set_store(VariableStore(device, u_seed))
    return get_store()
