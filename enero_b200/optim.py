"""Host-side mirror of ``tf.keras.optimizers.Adam(learning_rate, beta_1=0.9, epsilon=1e-05)``
as the reference builds it (train_Enero_3top_script.py:118): hyper-parameters, the step counter
and the m/v slots per model.  The update itself runs on the GPU (csrc/train_kernels.cuh)."""
from __future__ import annotations

import numpy as np

from .runtime import WEIGHT_SHAPES


class Adam:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-07, decay=0.0):
        self.learning_rate = np.float32(learning_rate)
        self.beta_1 = np.float32(beta_1)
        self.beta_2 = np.float32(beta_2)
        self.epsilon = float(epsilon)
        self.decay = np.float32(decay)
        self.iterations = 0
        self._slots = {}

    def _zero(self):
        return [np.zeros(s, dtype=np.float32) for s in WEIGHT_SHAPES]

    def get_slots(self, model):
        if id(model) not in self._slots:
            self._slots[id(model)] = (self._zero(), self._zero())
        return self._slots[id(model)]

    def set_slots(self, model, m, v):
        self._slots[id(model)] = ([np.array(x, dtype=np.float32) for x in m],
                                  [np.array(x, dtype=np.float32) for x in v])
