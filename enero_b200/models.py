"""Drop-in ``myModel`` classes of actorPPOmiddR.py / criticPPO.py whose ``call`` runs the
hand-written sm_100a message-passing kernels through the C ABI (enero_gnn_forward).

Same constructor, ``build()`` and call signatures, same hyper-parameters
(``link_state_dim``, ``readout_units``, ``T``), same 11 variables in Keras order."""
from __future__ import annotations

import numpy as np

from . import _lib
from .runtime import WEIGHT_SHAPES, default_context

VARIABLE_NAMES = ("FirstLayer/kernel:0", "FirstLayer/bias:0", "gru_cell/kernel:0", "gru_cell/recurrent_kernel:0",
                  "gru_cell/bias:0", "Readout1/kernel:0", "Readout1/bias:0", "Readout2/kernel:0", "Readout2/bias:0",
                  "Readout3/kernel:0", "Readout3/bias:0")


class Variable:
    """Minimal stand-in for tf.Variable: .numpy(), .assign(), .shape, .name."""

    def __init__(self, owner, index):
        self._owner, self._index = owner, index
        self.name = VARIABLE_NAMES[index]

    def numpy(self):
        return self._owner.get_weights()[self._index]

    def assign(self, value):
        w = self._owner.get_weights()
        w[self._index] = np.asarray(value, dtype=np.float32).reshape(WEIGHT_SHAPES[self._index])
        self._owner.set_weights(w)
        return self

    @property
    def shape(self):
        return WEIGHT_SHAPES[self._index]

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a


def orthogonal(shape, gain, rng):
    """Orthogonal initialiser (same recipe as Keras: QR of a normal matrix, sign-fixed).  The
    reference seeds TensorFlow's RNG (train_Enero_3top_script.py:76-79), which cannot be reproduced
    without TensorFlow; trained weights come from checkpoints."""
    rows, cols = shape[0], int(np.prod(shape[1:]))
    a = rng.normal(0.0, 1.0, (max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return (gain * q[:rows, :cols]).reshape(shape).astype(np.float32)


def glorot_uniform(shape, rng):
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, shape).astype(np.float32)


class _GNNModel:
    WHICH = _lib.ACTOR

    def __init__(self, hparams, hidden_init=None, kernel_init=None, context=None):
        if hparams["link_state_dim"] != 20 or hparams["readout_units"] != 20:
            raise ValueError("the sm_100a kernels are specialised for link_state_dim = readout_units = 20")
        self.hparams = hparams
        self._hidden_gain = getattr(hidden_init, "gain", np.sqrt(2.0)) if hidden_init is not None else np.sqrt(2.0)
        self._kernel_gain = getattr(kernel_init, "gain", None)
        self._ctx = context
        self._weights = None
        self._device_dirty = False
        self._version = 0            # bumped whenever the weights change (value caches key on it)
        self.built = False

    # ---- Keras-shaped surface ------------------------------------------------------------
    def build(self, input_shape=None):
        if self._weights is None:
            rng = np.random.RandomState(9)
            kg = self._kernel_gain if self._kernel_gain is not None else (np.sqrt(0.01) if self.WHICH == _lib.ACTOR else 1.0)
            w = [orthogonal((40, 20), self._hidden_gain, rng), np.zeros(20, np.float32),
                 glorot_uniform((20, 60), rng), orthogonal((20, 60), 1.0, rng), np.zeros((2, 60), np.float32),
                 orthogonal((20, 20), self._hidden_gain, rng), np.zeros(20, np.float32),
                 orthogonal((20, 20), self._hidden_gain, rng), np.zeros(20, np.float32),
                 orthogonal((20, 1), kg, rng), np.zeros(1, np.float32)]
            self.set_weights(w)
        self.built = True

    @property
    def context(self):
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    # ---- device residency ------------------------------------------------------------------
    # A context holds ONE actor and ONE critic weight buffer (plus their Adam slots).  Several model objects may
    # share a context (an eval agent next to a training agent, two TrainingRuns in one process): the context records
    # which object's weights each buffer currently holds, and bind() -- called by everything that is about to launch
    # with this model -- swaps the buffer over, first pulling a trained-on-device owner's weights back into its host
    # mirror so that nothing is lost or silently overwritten.
    def bind(self):
        ctx = self.context
        owners = ctx.__dict__.setdefault("_weight_owner", {})
        cur = owners.get(self.WHICH)
        if cur is self:
            return
        if self._weights is None:
            raise RuntimeError("model has no weights yet: call build() or set_weights() first")
        if cur is not None and cur._device_dirty:
            cur.sync_from_device()
        self._upload()

    def _upload(self):
        self.context.upload_weights(self.WHICH, self._weights)          # (clears the slot's owner)
        self.context.__dict__.setdefault("_weight_owner", {})[self.WHICH] = self

    def _owns_device(self):
        return self.context.__dict__.get("_weight_owner", {}).get(self.WHICH) is self

    def mark_device_dirty(self):
        """the device copy was updated by a GPU optimiser step: the host mirror is stale until the next read"""
        self._device_dirty = True
        self._version += 1

    def set_weights(self, tensors):
        w = [np.array(t, dtype=np.float32).reshape(s) for t, s in zip(tensors, WEIGHT_SHAPES)]
        if len(w) != 11:
            raise ValueError("expected 11 tensors in Keras variable order")
        owned = self._owns_device()
        self._weights = w
        self._device_dirty = False
        self._version += 1
        if owned:
            self._upload()

    def get_weights(self):
        if self._device_dirty and self._owns_device():
            self.sync_from_device()
        return [w.copy() for w in self._weights]

    def sync_from_device(self):
        """after a GPU optimiser step: refresh the host mirror"""
        if self._owns_device():
            self._weights = self.context.download_weights(self.WHICH)
        self._device_dirty = False

    @property
    def variables(self):
        return [Variable(self, i) for i in range(11)]

    trainable_weights = variables
    trainable_variables = variables


class ActorModel(_GNNModel):
    """actorPPOmiddR.myModel: call(link_state, states_graph_ids, states_first, states_second,
    sates_num_edges, training=False) -> [G, 1] logits (actorPPOmiddR.py:42-69)."""
    WHICH = _lib.ACTOR

    def call(self, link_state, states_graph_ids, states_first, states_second, sates_num_edges, training=False):
        ls = np.asarray(link_state, dtype=np.float32)
        gid = np.asarray(states_graph_ids, dtype=np.int32).reshape(-1)
        n = int(np.asarray(sates_num_edges))
        if ls.shape[0] != n:
            raise ValueError("num_edges (%d) must equal the number of link_state rows (%d)" % (n, ls.shape[0]))
        G = int(gid.max()) + 1 if gid.size else 0
        self.bind()
        out = self.context.gnn_forward(self.WHICH, ls, gid, states_first, states_second, G, T=int(self.hparams["T"]))
        return out.reshape(G, 1)

    __call__ = call


class CriticModel(_GNNModel):
    """criticPPO.myModel: call(link_state, first_critic, second_critic, num_edges_critic,
    training=False) -> [1, 1] value (criticPPO.py:38-64)."""
    WHICH = _lib.CRITIC

    def call(self, link_state, first_critic, second_critic, num_edges_critic, training=False):
        ls = np.asarray(link_state, dtype=np.float32)
        n = int(np.asarray(num_edges_critic))
        if ls.shape[0] != n:
            raise ValueError("num_edges (%d) must equal the number of link_state rows (%d)" % (n, ls.shape[0]))
        self.bind()
        out = self.context.gnn_forward(self.WHICH, ls, None, first_critic, second_critic, 1, T=int(self.hparams["T"]))
        return out.reshape(1, 1)

    __call__ = call
