"""Stand-in for the three TensorFlow names environment16.py touches
(convert_to_tensor/float32/int32 at environment16.py:523-524,574-575)."""
import numpy as np

float32 = np.float32
int32 = np.int32


def convert_to_tensor(value, dtype=None):
    return np.asarray(value, dtype=dtype)
