"""Independent anchors for the GNN oracle (the part the reference cannot pin: TensorFlow is not
installable).  These do not replace TF golden vectors -- parity against TF stays UNPINNED -- but they
tie the restated semantics to a second, independently written implementation of the same published
definitions:

  * the GRU step of oracle.message_passing equals torch.nn.GRUCell (the cuDNN-compatible cell that Keras'
    GRUCell(reset_after=True) documents itself to be) after re-ordering Keras' gates z,r,h to torch's r,z,n;
  * selu equals torch.nn.functional.selu (same published alpha / scale);
  * the two restatements (numpy oracle, torch PPO oracle) agree with each other;
  * unsorted_segment_sum / segment_sum are plain sums per destination (numpy add.at == torch index_add).
"""
import numpy as np
import torch

from conftest import case_graph_file, load_case
from oracle import enero_oracle as orc
from oracle import ppo_torch as ppo_o


def test_selu_matches_torch():
    x = np.linspace(-12, 12, 4001, dtype=np.float32)
    np.testing.assert_allclose(orc.selu(x), torch.nn.functional.selu(torch.tensor(x)).numpy(), rtol=2e-6, atol=3e-7)   # exp(x)-1 vs expm1 near 0-


def test_gru_step_matches_torch_grucell(weights):
    rng = np.random.default_rng(0)
    n = 64
    agg = rng.normal(0, 1, (n, 20)).astype(np.float32)
    h = rng.normal(0, 1, (n, 20)).astype(np.float32)
    kx, kh, b = weights["gru_kernel"], weights["gru_recurrent_kernel"], weights["gru_bias"]
    # oracle's step (message_passing body, lines following the segment sum)
    x = agg @ kx + b[0]
    y = h @ kh + b[1]
    z = orc.sigmoid(x[:, :20] + y[:, :20])
    r = orc.sigmoid(x[:, 20:40] + y[:, 20:40])
    hh = np.tanh(x[:, 40:] + r * y[:, 40:]).astype(np.float32)
    want = z * h + (1 - z) * hh
    # torch.nn.GRUCell: gates ordered r, z, n; weights stored transposed
    cell = torch.nn.GRUCell(20, 20)
    perm = np.concatenate([np.arange(20, 40), np.arange(0, 20), np.arange(40, 60)])      # keras (z,r,h) -> torch (r,z,n)
    with torch.no_grad():
        cell.weight_ih.copy_(torch.tensor(kx[:, perm].T.copy()))
        cell.weight_hh.copy_(torch.tensor(kh[:, perm].T.copy()))
        cell.bias_ih.copy_(torch.tensor(b[0][perm].copy()))
        cell.bias_hh.copy_(torch.tensor(b[1][perm].copy()))
        got = cell(torch.tensor(agg), torch.tensor(h)).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6)


def test_numpy_and_torch_restatements_agree(weights, tmp_path):
    case = load_case("tiny12")
    ot = orc.Topology(case_graph_file(case, tmp_path), K=100)
    env = orc.Env16(ot)
    _, s, d = env.reset(np.array(case["tms"]["0"]))
    probs, batch, logits = orc.pred_action_distrib(weights, env, s, d)
    w = ppo_o.to_params(weights)
    tl = ppo_o.actor_logits(w, batch).detach().numpy()
    np.testing.assert_allclose(tl, logits, rtol=1e-5, atol=1e-5 * np.abs(logits).max())
    v_np = orc.critic_forward(weights, env.critic_features(), ot.first, ot.second, ot.E).reshape(())
    v_t = ppo_o.critic_value(w, env.critic_features(), ot.first, ot.second, ot.E).detach().numpy()
    np.testing.assert_allclose(v_t, v_np, rtol=1e-5)


def test_segment_sums_are_order_preserving_sums():
    rng = np.random.default_rng(1)
    m = rng.normal(0, 1, (500, 20)).astype(np.float32)
    seg = rng.integers(0, 37, 500)
    a = np.zeros((37, 20), np.float32)
    np.add.at(a, seg, m)                      # sequential in pair order, like TF-CPU's unsorted_segment_sum
    b = np.zeros((37, 20), np.float32)
    for i in range(500):
        b[seg[i]] += m[i]
    assert np.array_equal(a, b)
    c = torch.zeros(37, 20).index_add(0, torch.tensor(seg), torch.tensor(m)).numpy()
    np.testing.assert_allclose(c, a, rtol=1e-5, atol=1e-5)
