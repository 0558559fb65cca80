"""Device-resident sampled rollouts (enero_batch_rollout, train_Enero_3top_script.py:483-625): the records of a
whole batch of episodes against the same episodes driven decision by decision through policy() / value() / step()."""
import numpy as np
import pytest

from conftest import case_graph_file, load_case
from oracle import enero_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(weights):
    from enero_b200.runtime import Context
    c = Context(0)
    flat = [weights[k] for k in orc.WEIGHT_NAMES]
    c.upload_weights(0, flat)
    rng = np.random.default_rng(7)
    c.upload_weights(1, [w + rng.normal(0, 0.05, w.shape).astype(np.float32) for w in flat])
    yield c
    c.close()


def _batch(ctx, name, n, tmp_path):
    from enero_b200 import datasets as ds
    from enero_b200.engine import EpisodeBatch
    from enero_b200.topology import Topology
    gf = case_graph_file(load_case(name), tmp_path)
    topo = Topology(gf, K=100)
    sigma = ds.sigma_for_mean_utilisation(gf)
    tms = np.stack([ds.uniform_tm(gf.num_nodes, sigma, 4000 + j) for j in range(n)])
    return EpisodeBatch(ctx, topo, n), tms, topo


def test_rollout_records_replay_step_by_step(ctx, tmp_path):
    eb, tms, topo = _batch(ctx, "tiny12", 6, tmp_path)
    eb.reset(tms)
    r = eb.rollout(seed=1234)
    assert (r["steps"] > 0).all()
    # determinism and independence of the batch composition: same seed, same records; episode b alone gives row b
    eb.reset(tms)
    r2 = eb.rollout(seed=1234)
    for k in r:
        assert np.array_equal(r[k], r2[k]), k
    eb.reset(tms)
    r3 = eb.rollout(seed=99)
    assert not np.array_equal(r["action"], r3["action"])
    # replay: the recorded actions through the stepwise API reproduce states, probabilities, values and rewards
    eb.reset(tms)
    for s in range(int(r["steps"].max())):
        live = s < r["steps"]
        st = eb.state()
        probs, nK = eb.policy()
        vals = eb.value()
        for b in np.nonzero(live)[0]:
            assert np.array_equal(st["loads"][b], r["loads"][b, s])
            assert (st["cur"][b] == r["sd"][b, s]).all() and st["bw"][b] == r["bw"][b, s]
            a = int(r["action"][b, s])
            assert 0 <= a < nK[b]
            assert probs[b, a] == r["prob"][b, s] and vals[b] == r["value"][b, s]
        acts = np.where(live, r["action"][:, min(s, r["action"].shape[1] - 1)], 0).astype(np.int32)
        reward, done = eb.step(acts)
        for b in np.nonzero(live)[0]:
            assert reward[b] == r["reward"][b, s]
            assert bool(done[b]) == (s == r["steps"][b] - 1)
    assert np.array_equal(eb.value(), r["last_value"])
    eb.close()


def test_sampled_actions_follow_the_policy(ctx, tmp_path):
    """first decision of 4096 copies of one traffic matrix: the empirical action frequencies match the actor's
    probabilities (chi-square with a generous bound; a broken inverse CDF or a correlated generator fails by orders)"""
    from enero_b200.engine import EpisodeBatch
    eb1, tms, topo = _batch(ctx, "tiny12", 1, tmp_path)
    n = 4096
    eb = EpisodeBatch(ctx, topo, n)
    eb.reset(np.repeat(tms[:1], n, axis=0))
    probs, nK = eb.policy()
    k = int(nK[0])
    r = eb.rollout(seed=7)
    counts = np.bincount(r["action"][:, 0], minlength=k)[:k].astype(np.float64)
    expect = probs[0, :k].astype(np.float64) * n
    chi2 = float(((counts - expect) ** 2 / np.maximum(expect, 1e-9)).sum())
    assert chi2 < 5.0 * k + 30.0, (chi2, counts, expect)
    eb.close()
    eb1.close()
