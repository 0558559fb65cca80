import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASE_NAMES = ("tri9", "tiny12", "eli20", "zoo30", "hur24")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_case(name):
    with open(os.path.join(GOLDEN, "case_%s.json" % name)) as fp:
        return json.load(fp)


def case_graph_file(case, tmp_path):
    from enero_b200 import datasets as ds
    p = os.path.join(str(tmp_path), case["name"] + ".graph")
    with open(p, "w") as fp:
        fp.write(case["graph_text"])
    return ds.parse_graph_file(p)


def golden_weights():
    z = np.load(os.path.join(GOLDEN, "ckpt_ACT-1835_weights.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def weights():
    return golden_weights()
