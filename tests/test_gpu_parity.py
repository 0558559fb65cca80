"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle and the
reference-recorded golden fixtures.

Tolerances: integer / fp64 quantities (indices, loads, utilisations, rewards, action and
move sequences) bit-exact.  fp32 logits and values ELEMENT-WISE within 1e-5 relative, as
BASELINE.json's north_star states, with an absolute floor:
        |got - want| <= 1e-5 * |want| + 1e-5        for every element.
Why the floor is 1e-5: a logit is a sum over ~100 link states and 5 iterations of O(1) terms, each with its
own fp32 rounding, so its error does not shrink when the sum happens to land near zero.  The fp64 arbiter
(oracle.actor_forward64) measures that rounding floor on 17 812 decisions (tools/parity_sweep.py,
profiles/r02_parity_sweep.json): the numpy fp32 oracle itself needs a floor of 1.02e-5 to stay within
rtol 1e-5 of the arbiter, the CUDA path 8.3e-6 (fast math) / 6.6e-6 (accurate math) -- no two fp32
implementations of this network (TensorFlow's included) can be pinned to each other more tightly.  It is also
torch.testing's default absolute tolerance for fp32.
Probabilities are soft-max outputs: an absolute logit error d becomes a relative probability error
of up to 2d, so they are held, purely relatively, to the propagated bound 2 * (1e-5 * max|logit| + ATOL)
(north_star lists logits, values, utilisations and rewards, not probabilities, under the 1e-5 rule).
"""
import numpy as np
import pytest

from conftest import CASE_NAMES, case_graph_file, load_case
from oracle import enero_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ATOL = 1e-5


def assert_close(got, want, rtol=RTOL, atol=ATOL):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    want = np.asarray(want, dtype=np.float64).reshape(-1)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


def assert_probs_close(got, want, logits):
    """soft-max propagation of the logit tolerance: |dp/p| <= 2 max|dlogit|, purely relative"""
    assert_close(got, want, rtol=2 * (RTOL * float(np.abs(logits).max()) + ATOL), atol=1e-9)


@pytest.fixture(scope="module")
def ctx(weights):
    from enero_b200.runtime import Context
    c = Context(0)
    flat = [weights[k] for k in orc.WEIGHT_NAMES]
    c.upload_weights(0, flat)
    # no critic checkpoint is shipped: a deterministic perturbation of the actor stands in
    rng = np.random.default_rng(7)
    c.upload_weights(1, [w + rng.normal(0, 0.05, w.shape).astype(np.float32) for w in flat])
    yield c
    c.close()


@pytest.fixture(scope="module")
def critic_weights(ctx):
    return dict(zip(orc.WEIGHT_NAMES, ctx.download_weights(1)))


def _topos(name, tmp_path_factory):
    from enero_b200.topology import Topology
    case = load_case(name)
    gf = case_graph_file(case, tmp_path_factory.mktemp(name))
    return case, Topology(gf, K=100), orc.Topology(gf, K=100)


@pytest.fixture(scope="module", params=CASE_NAMES)
def trio(request, tmp_path_factory):
    return _topos(request.param, tmp_path_factory)


def test_weights_roundtrip(ctx, weights):
    got = ctx.download_weights(0)
    for g, k in zip(got, orc.WEIGHT_NAMES):
        assert np.array_equal(g, weights[k])


def test_gnn_forward_signature_path(ctx, trio, weights, critic_weights):
    """actor.myModel.call / critic.myModel.call with explicit first/second (quirk Q1 included:
    tiny12 and zoo30 have max(second)+1 < E)."""
    case, topo, ot = trio
    ep = case["episodes"][0]
    env = orc.Env16(ot)
    _, s, d = env.reset(np.array(case["tms"][str(ep["tm_id"])]))
    for step in range(3):
        feats = env.candidate_features(s, d)
        b = orc.batch_candidates(feats, ot.first, ot.second)
        want = orc.actor_forward(weights, b["link_state"], b["graph_id"], b["first"], b["second"], b["num_edges"]).reshape(-1)
        got = ctx.gnn_forward(0, b["link_state"], b["graph_id"], b["first"], b["second"], len(feats))
        assert_close(got, want)
        cf = env.critic_features()
        wantv = orc.critic_forward(critic_weights, cf, ot.first, ot.second, ot.E).reshape(-1)
        gotv = ctx.gnn_forward(1, cf, None, ot.first, ot.second, 1)
        assert_close(gotv, wantv)
        _, _, _, _, s, d, _, _, _ = env.step(ep["steps"][step]["action"])


def test_gnn_forward_arbitrary_indices(ctx, weights):
    """unsorted, repeated and missing destinations; empty graphs; shuffled pair order."""
    rng = np.random.default_rng(3)
    n, p, G = 257, 900, 7
    ls = rng.normal(0, 0.5, (n, 20)).astype(np.float32)
    first = rng.integers(0, n, p).astype(np.int32)
    second = rng.integers(0, n - 40, p).astype(np.int32)       # rows >= n-40 receive nothing
    gid = np.sort(rng.choice([0, 1, 2, 4, 6], n)).astype(np.int32)   # graphs 3 and 5 are empty
    want = orc.actor_forward(weights, ls, gid, first, second, n).reshape(-1)
    full = np.zeros(G, dtype=np.float32)
    full[:] = orc.readout(weights, np.zeros((1, 20), dtype=np.float32)).reshape(-1)[0]
    full[:want.size] = want
    for g in (3, 5):
        full[g] = orc.readout(weights, np.zeros((1, 20), dtype=np.float32)).reshape(-1)[0]
    got = ctx.gnn_forward(0, ls, gid, first, second, G)
    assert_close(got, full)


def test_gnn_forward_rejects_bad_indices(ctx):
    from enero_b200._lib import EneroError
    ls = np.zeros((8, 20), dtype=np.float32)
    with pytest.raises(EneroError):
        ctx.gnn_forward(0, ls, np.zeros(8, np.int32), np.array([0, 9], np.int32), np.array([1, 2], np.int32), 1)
    with pytest.raises(EneroError):
        ctx.gnn_forward(0, ls, np.array([1, 0, 0, 0, 0, 0, 0, 0], np.int32), np.array([0], np.int32), np.array([1], np.int32), 2)


def test_decide_batch_matches_oracle_and_golden(ctx, trio, weights):
    case, topo, ot = trio
    for ep in case["episodes"]:
        tm = np.array(case["tms"][str(ep["tm_id"])])
        decs = ep["decisions"]
        B = len(decs)
        loads = np.array([dc["loads_before"] for dc in decs])
        sd = np.array([[dc["s"], dc["d"]] for dc in decs], dtype=np.int32)
        bw = np.array([tm[dc["s"], dc["d"]] for dc in decs])
        logits, probs, nK, action = ctx.decide_batch(topo, loads, sd, bw)
        for b, dc in enumerate(decs):
            k = len(dc["logits"])
            assert nK[b] == k
            assert_close(logits[b, :k], dc["logits"])
            assert_probs_close(probs[b, :k], dc["probs"], dc["logits"])
            assert not logits[b, k:].any() and not probs[b, k:].any()
            assert action[b] == ep["steps"][b]["action"]


def test_value_batch_matches_oracle(ctx, trio, critic_weights):
    case, topo, ot = trio
    ep = case["episodes"][0]
    loads = np.array([dc["loads_before"] for dc in ep["decisions"][:8]])
    got = ctx.value_batch(topo, loads)
    for b in range(loads.shape[0]):
        ls = np.zeros((ot.E, 20), dtype=np.float32)
        ls[:, 0] = (loads[b] / ot.cap).astype(np.float32)
        ls[:, 1] = ot.link_capacity_feature
        want = orc.critic_forward(critic_weights, ls, ot.first, ot.second, ot.E).reshape(-1)[0]
        assert_close(got[b], want)


def test_greedy_episodes_and_local_search_bit_exact(ctx, trio):
    """Whole Enero pipeline on device against the reference-recorded trace: greedy action
    sequence, rewards, max utilisation per step, best routing snapshot, LS move sequence and
    final max link utilisation -- all bit-exact."""
    from enero_b200.engine import EpisodeBatch
    case, topo, ot = trio
    eps = case["episodes"]
    tms = np.array([case["tms"][str(ep["tm_id"])] for ep in eps])
    eb = EpisodeBatch(ctx, topo, len(eps))
    bw, cur = eb.reset(tms)
    elig, n_elig = eb.eligible()
    st = eb.state()
    for b, ep in enumerate(eps):
        r = ep["reset"]
        assert (bw[b], cur[b, 0], cur[b, 1]) == (r["demand"], r["s"], r["d"])
        assert [[int(s), int(d)] for s, d in elig[b, :n_elig[b]]] == [[e[0], e[1]] for e in r["elig"]]
        assert st["max_uti"][b] == r["edge_max"][2]
        assert st["loads"][b].tolist() == ep["decisions"][0]["loads_before"]
    eb.run_greedy()
    act, rew, mx = eb.history()
    st = eb.state()
    best, ospf, routing = eb.best()
    for b, ep in enumerate(eps):
        n = len(ep["steps"])
        assert st["steps"][b] == n and st["done"][b]
        assert act[b, :n].tolist() == [s["action"] for s in ep["steps"]]
        assert rew[b, :n].tolist() == [s["reward"] for s in ep["steps"]]
        assert mx[b, :n].tolist() == [s["edge_max"][2] for s in ep["steps"]]
        assert st["loads"][b].tolist() == ep["steps"][-1]["loads_after"]
        last = ep["steps"][-1]["edge_max"]
        assert st["max_edge"][b] == ot.eid[(last[0], last[1])]
        assert (best[b], ospf[b]) == (ep["best"], ep["ospf"])
        assert [list(r) for r in routing[b]] == ep["best_routing"]
    ls = eb.local_search(None, mode=0)
    le, ln = eb.ls_eligible()
    for b, ep in enumerate(eps):
        g = ep["ls"]
        assert ls["start"][b] == -g["reset_val"]
        assert [[int(s), int(d)] for s, d in le[b, :ln[b]]] == [[e[0], e[1]] for e in g["elig"]]
        n = ls["n_moves"][b]
        got = [[int(ls["moves"][b, k, 0]), int(ls["moves"][b, k, 1]), int(ls["moves"][b, k, 2]), float(ls["move_val"][b, k])]
               for k in range(n)]
        assert got == g["moves"]
        assert ls["final"][b] == g["final"]
    # the same Local Search from an explicitly supplied routing
    ls2 = eb.local_search([ [tuple(r) for r in ep["best_routing"]] for ep in eps ], mode=0)
    assert ls2["final"].tolist() == [ep["ls"]["final"] for ep in eps]
    eb.close()


def test_stepwise_env_matches_trace(ctx, trio):
    """policy() / step(actions) one decision at a time (the training-rollout usage)."""
    from enero_b200.engine import EpisodeBatch
    case, topo, ot = trio
    ep = case["episodes"][0]
    eb = EpisodeBatch(ctx, topo, 1)
    eb.reset(np.array([case["tms"][str(ep["tm_id"])]]))
    for k, (stp, dc) in enumerate(zip(ep["steps"][:6], ep["decisions"][:6])):
        probs, nK = eb.policy()
        assert_probs_close(probs[0, :nK[0]], dc["probs"], dc["logits"])
        reward, done = eb.step([stp["action"]])
        assert reward[0] == stp["reward"] and bool(done[0]) == stp["done"]
        st = eb.state()
        assert st["loads"][0].tolist() == stp["loads_after"]
        assert [st["bw"][0], st["cur"][0, 0], st["cur"][0, 1]] == stp["next"]
    eb.close()


@pytest.mark.parametrize("name", ["tri9", "tiny12"])
def test_plain_hill_climbing(ctx, name, tmp_path_factory):
    from enero_b200.engine import EpisodeBatch
    case, topo, ot = _topos(name, tmp_path_factory)
    eps = case["episodes"]
    eb = EpisodeBatch(ctx, topo, len(eps))
    eb.reset(np.array([case["tms"][str(ep["tm_id"])] for ep in eps]))
    ls = eb.local_search(None, mode=1, want_moves=False)
    assert ls["final"].tolist() == [ep["plain_hill_final"] for ep in eps]
    eb.close()


def test_full_size_properties(ctx, tmp_path):
    """BASELINE config 2 size (100 links): size-independent properties on 48 synthetic TMs."""
    from enero_b200 import datasets as ds
    from enero_b200.engine import EpisodeBatch
    from enero_b200.topology import Topology
    links, caps = ds.zoo_like(30, 50, 4242)
    p = str(tmp_path / "g.graph")
    ds.write_graph_file(p, 30, links, caps)
    gf = ds.parse_graph_file(p)
    topo = Topology(gf, K=100)
    sigma = ds.sigma_for_mean_utilisation(gf)
    base = [ds.uniform_tm(30, sigma, 9000 + j) for j in range(24)]
    tms = np.array(base + base)                     # every TM twice: determinism across batch slots
    eb = EpisodeBatch(ctx, topo, len(tms))
    eb.reset(tms)
    st0 = eb.state()
    # conservation: total load == sum over demands of bw * hops, with the first demand removed
    hops = np.array([[len(topo.sp_edges(s, d)) if s != d else 0 for d in range(30)] for s in range(30)])
    for b in range(len(tms)):
        s, d = st0["cur"][b]
        assert st0["loads"][b].sum() == (tms[b] * hops).sum() - tms[b][s, d] * hops[s, d]
    eb.run_greedy()
    best, ospf, routing = eb.best()
    act, rew, mx = eb.history()
    st = eb.state()
    assert (best <= ospf).all() and st["done"].all()
    assert np.array_equal(act[:24], act[24:]) and np.array_equal(mx[:24], mx[24:])
    # telescoping rewards: sum of un-rounded rewards == 10 * (ospf - final max)
    for b in range(len(tms)):
        n = st["steps"][b]
        assert abs(rew[b, :n].sum() - 10 * (ospf[b] - mx[b, n - 1])) < 0.005 * n + 1e-9
    ls = eb.local_search(None, mode=0)
    assert (ls["final"] <= ls["start"] + 1e-12).all()
    assert np.array_equal(ls["final"][:24], ls["final"][24:])
    # Local Search restarts from the DRL best routing: same max utilisation, unless the best was the
    # final step (whose snapshot lacks the dummy demand (1,2), environment16.py:622-635)
    for b in range(len(tms)):
        if mx[b, st["steps"][b] - 1] > best[b]:
            assert ls["start"][b] == best[b]
    # idempotence: a second LS from the LS result makes no move -- replay moves to build that routing
    final_routing = []
    for b in range(len(tms)):
        cur = {(s, d): m for s, d, m in routing[b]}
        for k in range(ls["n_moves"][b]):
            a, s, d = (int(v) for v in ls["moves"][b, k])
            m = topo.middlepoints(s, d)[a]
            cur.pop((s, d), None)
            if m != d:
                cur[(s, d)] = m
        final_routing.append([(s, d, m) for (s, d), m in cur.items()])
    if max(len(r) for r in final_routing) <= eb.Dmax:
        ls2 = eb.local_search(final_routing, mode=0)
        assert np.array_equal(ls2["start"], ls["final"])
    eb.close()


@pytest.mark.parametrize("name,n_tms", [("tiny12", 24), ("eli20", 6)])
def test_greedy_actions_many_tms(ctx, weights, name, n_tms, tmp_path_factory):
    """north_star: 'the greedy action sequence and final max link utilisation must be identical on the
    shipped checkpoint' -- checked on fresh synthetic TMs (not only the recorded golden ones), whole
    batch at once on the GPU against the oracle one episode at a time, Local Search included."""
    from enero_b200 import datasets as ds
    from enero_b200.engine import EpisodeBatch
    case, topo, ot = _topos(name, tmp_path_factory)
    gf = case_graph_file(case, tmp_path_factory.mktemp(name + "many"))
    sigma = ds.sigma_for_mean_utilisation(gf)
    tms = np.stack([ds.uniform_tm(gf.num_nodes, sigma, 7000 + j) for j in range(n_tms)])
    eb = EpisodeBatch(ctx, topo, n_tms)
    eb.reset(tms)
    eb.run_greedy()
    best, ospf, routing = eb.best()
    act, rew, hmax = eb.history()
    steps = eb.state()["steps"]
    ls = eb.local_search(None, mode=0)
    for i in range(n_tms):
        obest, oospf, orouting, oactions = orc.play_middrout_greedy(weights, orc.Env16(ot), tms[i])
        assert act[i, :steps[i]].tolist() == oactions, "TM %d: action sequences differ" % i
        assert (best[i], ospf[i]) == (obest, oospf)
        e15 = orc.Env15(ot)
        cur = e15.reset_DRL_hill_sp(tms[i], orouting)
        ofinal, _ = orc.local_search(e15, cur)
        assert ls["final"][i] == ofinal
    eb.close()


def test_math_modes_against_fp64_arbiter(ctx, trio, weights):
    """Both math modes (fast = default, accurate = opt-in) against the fp64 arbiter (oracle.actor_forward64) on the
    same states, at the same element-wise tolerance as against the fp32 oracle; both must pick the arbiter's
    arg-max.  tools/parity_sweep.py runs this comparison over 692 episodes (profiles/r02_parity_sweep.json)."""
    case, topo, ot = trio
    ep = case["episodes"][0]
    tm = np.array(case["tms"][str(ep["tm_id"])])
    decs = ep["decisions"][:12]
    loads = np.array([dc["loads_before"] for dc in decs])
    sd = np.array([[dc["s"], dc["d"]] for dc in decs], dtype=np.int32)
    bw = np.array([tm[dc["s"], dc["d"]] for dc in decs])
    truth = []
    env = orc.Env16(ot)
    env.tm = tm
    for b, dc in enumerate(decs):
        env.loads = loads[b].copy()
        bt = orc.batch_candidates(env.candidate_features(dc["s"], dc["d"]), ot.first, ot.second)
        truth.append(orc.actor_forward64(weights, bt["link_state"], bt["graph_id"], bt["first"], bt["second"], bt["num_edges"]).reshape(-1))
    before = ctx.math
    try:
        for mode in ("accurate", "fast"):
            ctx.math = mode
            logits, _, nK, action = ctx.decide_batch(topo, loads, sd, bw)
            for b, t in enumerate(truth):
                np.testing.assert_allclose(logits[b, :nK[b]], t, rtol=RTOL, atol=ATOL)
                top2 = np.sort(t)[-2:]
                if top2[1] - top2[0] > 4 * ATOL:           # the arbiter's arg-max, unless its two best are within the floor
                    assert action[b] == int(np.argmax(t))
    finally:
        ctx.math = before


def test_accurate_math_greedy_episode_identical(ctx, weights, tmp_path_factory):
    """the opt-in accurate math mode reproduces the oracle's greedy action sequences too (eli20-sized graph)"""
    from enero_b200.engine import EpisodeBatch
    case, topo, ot = _topos("tiny12", tmp_path_factory)
    tms = np.array([case["tms"][str(ep["tm_id"])] for ep in case["episodes"]])
    before = ctx.math
    try:
        ctx.math = "accurate"
        eb = EpisodeBatch(ctx, topo, len(tms))
        eb.reset(tms)
        eb.run_greedy()
        act, _, _ = eb.history()
        for b, ep in enumerate(case["episodes"]):
            assert act[b, :len(ep["steps"])].tolist() == [s["action"] for s in ep["steps"]]
        eb.close()
    finally:
        ctx.math = before


def test_caller_stream_and_engines_agree(ctx, weights, tmp_path_factory):
    """enero_ctx_set_stream (the stream-argument form SURVEY 8b asks for, cuBLAS style): the same decisions on a
    caller-owned torch stream; and the two message-passing engines (k_mp_tc tensor cores, k_mp_iter FP32 pipe) agree
    within the element-wise tolerance and on every greedy action."""
    import torch
    case, topo, ot = _topos("eli20", tmp_path_factory)
    ep = case["episodes"][0]
    tm = np.array(case["tms"][str(ep["tm_id"])])
    decs = ep["decisions"][:10]
    loads = np.array([dc["loads_before"] for dc in decs])
    sd = np.array([[dc["s"], dc["d"]] for dc in decs], dtype=np.int32)
    bw = np.array([tm[dc["s"], dc["d"]] for dc in decs])
    base = ctx.decide_batch(topo, loads, sd, bw)
    s = torch.cuda.Stream(device=ctx.device)
    ctx.set_stream(s)
    try:
        assert ctx.stream == s.cuda_stream
        mine = ctx.decide_batch(topo, loads, sd, bw)
    finally:
        ctx.set_stream(None)
    for a, b in zip(base, mine):
        assert np.array_equal(a, b)
    before = ctx.engine
    try:
        out = {}
        for eng in ("tc", "ffma"):
            ctx.engine = eng
            out[eng] = ctx.decide_batch(topo, loads, sd, bw)
        np.testing.assert_allclose(out["tc"][0], out["ffma"][0], rtol=RTOL, atol=ATOL)
        assert np.array_equal(out["tc"][3], out["ffma"][3]) and np.array_equal(out["tc"][2], out["ffma"][2])
    finally:
        ctx.engine = before


def test_decision_tail_variants_identical(weights, tmp_path_factory, monkeypatch):
    """The tail of a decision runs either as one launch (k_decision_finish: read-out + soft-max + arg-max + Env16.step,
    picked below 512 decisions in flight) or as k_readout + k_policy + k_env_step (picked above): same arithmetic, so
    logits, probabilities, greedy histories and best routings must be bit-identical, whichever the batch size picks."""
    from enero_b200.runtime import Context
    from enero_b200.engine import EpisodeBatch
    from enero_b200 import datasets as ds
    case, topo, _ = _topos("tiny12", tmp_path_factory)
    gf = case_graph_file(case, tmp_path_factory.mktemp("tail"))
    sigma = ds.sigma_for_mean_utilisation(gf)
    B = 24
    tms = np.stack([ds.uniform_tm(gf.num_nodes, sigma, 300 + j) for j in range(B)])
    got = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("ENERO_FUSED_TAIL", mode)
        c = Context(0)
        try:
            c.upload_weights(0, [weights[k] for k in orc.WEIGHT_NAMES])
            eb = EpisodeBatch(c, topo, B)
            eb.reset(tms)
            st = eb.state()
            dec = c.decide_batch(topo, st["loads"], st["cur"], st["bw"])
            eb.run_greedy()
            got[mode] = (dec, eb.history(), eb.best_arrays(), eb.state())
            eb.close()
        finally:
            c.close()
    monkeypatch.delenv("ENERO_FUSED_TAIL")
    d0, h0, b0, s0 = got["0"]
    d1, h1, b1, s1 = got["1"]
    for a, b in zip(d0, d1):
        assert np.array_equal(a, b)
    for a, b in zip(h0 + b0, h1 + b1):
        assert np.array_equal(a, b)
    for k in s0:
        assert np.array_equal(s0[k], s1[k]), k
    assert int(s0["steps"].sum()) > B
