"""TensorBundle reader/writer against the shipped ckpt_ACT-1835 (CRC-protected golden file)."""
import filecmp
import os

import numpy as np

from conftest import GOLDEN
from enero_b200 import checkpoint as ck
from enero_b200.optim import Adam

CKPT = os.path.join(GOLDEN, "ckpt_ACT-1835")


class _Model:
    def __init__(self):
        self.w = None

    def set_weights(self, w):
        self.w = [np.array(x) for x in w]

    def get_weights(self):
        return self.w


def test_reader_known_answers():
    b = ck.Bundle.load(CKPT)                         # verifies masked crc32c of every tensor
    sc = b.optimizer_scalars()
    assert sc["iter"] == 234880 == 1835 * 8 * 16
    assert sc["save_counter"] == 1835
    assert abs(sc["learning_rate"] - 2e-4) < 1e-9 and abs(sc["beta_1"] - 0.9) < 1e-7 and abs(sc["beta_2"] - 0.999) < 1e-7
    w = b.model_weights()
    assert [x.shape for x in w] == list(ck.MODEL_SHAPES)
    assert sum(x.size for x in w) == 4201
    assert len(b.tensors) == 40
    z = np.load(os.path.join(GOLDEN, "ckpt_ACT-1835_weights.npz"))
    for x, k in zip(w, ("msg_kernel", "msg_bias", "gru_kernel", "gru_recurrent_kernel", "gru_bias", "r1_kernel",
                        "r1_bias", "r2_kernel", "r2_bias", "r3_kernel", "r3_bias")):
        assert np.array_equal(x, z[k])


def test_crc_detects_corruption(tmp_path):
    data = bytearray(open(CKPT + ".data-00000-of-00001", "rb").read())
    data[5000] ^= 0x40
    p = str(tmp_path / "bad")
    open(p + ".data-00000-of-00001", "wb").write(bytes(data))
    open(p + ".index", "wb").write(open(CKPT + ".index", "rb").read())
    try:
        ck.Bundle.load(p)
    except ValueError as exc:
        assert "crc" in str(exc)
    else:
        raise AssertionError("corruption not detected")


def test_checkpoint_save_reproduces_shipped_files_byte_for_byte(tmp_path):
    model, opt = _Model(), Adam(learning_rate=2e-4, beta_1=0.9, epsilon=1e-5)
    c = ck.Checkpoint(model=model, optimizer=opt).restore(CKPT)
    assert opt.iterations == 234880 and c.save_counter == 1835
    c.save_counter = 1834                            # .save() increments, like tf.train.Checkpoint
    path = c.save(str(tmp_path / "ckpt_ACT"))
    assert path.endswith("ckpt_ACT-1835")
    for ext in (".index", ".data-00000-of-00001"):
        assert filecmp.cmp(path + ext, CKPT + ext, shallow=False), ext
    assert 'model_checkpoint_path: "ckpt_ACT-1835"' in open(str(tmp_path / "checkpoint")).read()
