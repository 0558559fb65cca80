"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of ENERO's hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module, and only as the checker or the timed CPU
baseline.  The product (enero_b200/) never imports it.

Parity status
-------------
* Topology / indexing / environment / Local Search: PINNED.  oracle/gen_golden.py
  runs the reference's own, unmodified environment16.py / environment15.py /
  script_eval_on_single_topology.py (under the stubs in oracle/stubs) on seeded
  synthetic datasets and tests/test_oracle_golden.py checks this restatement
  against those recorded outputs bit for bit.
* GNN actor/critic and PPO arithmetic: PARITY UNPINNED.  TensorFlow 2.4.1 /
  Keras 2.4.3 (requirements.txt:67,33) cannot be installed here and the
  reference ships no golden vectors; the functions below restate the published
  Keras/TF semantics (Dense, selu, unsorted_segment_sum, GRUCell reset_after=True,
  segment_sum, softmax) following actorPPOmiddR.py:42-69 / criticPPO.py:38-64
  with the shipped ckpt_ACT-1835 weights.

Every function cites the reference lines it follows (relative to /root/reference).
"""
from __future__ import annotations

import math

import numpy as np

SELU_ALPHA = np.float32(1.6732632423543772)
SELU_SCALE = np.float32(1.0507009873554805)

# Keras variable order, train_Enero_3top_script.py:238-248
WEIGHT_NAMES = ("msg_kernel", "msg_bias", "gru_kernel", "gru_recurrent_kernel", "gru_bias",
                "r1_kernel", "r1_bias", "r2_kernel", "r2_bias", "r3_kernel", "r3_bias")
WEIGHT_SHAPES = ((40, 20), (20,), (20, 60), (20, 60), (2, 60), (20, 20), (20,), (20, 20),
                 (20,), (20, 1), (1,))


# ==========================================================================
# topology (environment16.py:526-580, 507-524, 478-505, 395-476)
# ==========================================================================
class Topology:
    def __init__(self, graph_file, K=100, paths=None):
        """graph_file: enero_b200.datasets.GraphFile (a faithful parse of the
        .graph format).  paths: optional {"s:d": [nodes]} (shortest_paths.json)."""
        self.N = graph_file.num_nodes
        self.node_order = list(graph_file.node_order)
        self.adj = {u: list(v) for u, v in graph_file.adj.items()}
        # edge ids: environment16.py:557-569
        self.edges = []
        self.eid = {}
        for i in self.node_order:
            for j in self.adj[i]:
                self.eid[(i, j)] = len(self.edges)
                self.edges.append((i, j))
        self.E = len(self.edges)
        self.cap = np.array([graph_file.links_bw[e] for e in self.edges], dtype=np.float64)
        self.max_cap = float(self.cap.max())
        # environment16.py:574
        self.link_capacity_feature = (self.cap / self.max_cap).astype(np.float32)
        self.K = min(K, self.N)                         # environment16.py:549-551
        # pairs: environment16.py:507-524
        first, second = [], []
        for i in self.node_order:
            for j in self.adj[i]:
                for n in self.adj[j]:                   # graph.edges(j) = out-edges of j
                    m = j
                    if (i != m or j != n) and (i != n or j != m):
                        first.append(self.eid[(i, j)])
                        second.append(self.eid[(m, n)])
        self.first = np.asarray(first, dtype=np.int32)
        self.second = np.asarray(second, dtype=np.int32)
        self.P = len(first)
        # shortest paths: environment16.py:478-505 (closed form: SURVEY Q3)
        if paths is None:
            paths = _bfs_lexmin_paths(self.N, self.node_order, self.adj)
        self.paths = {}
        self.sp_edges = {}
        for s in range(self.N):
            for d in range(self.N):
                if s != d:
                    p = list(paths["%d:%d" % (s, d)])
                    self.paths[(s, d)] = p
                    self.sp_edges[(s, d)] = [self.eid[(a, b)] for a, b in zip(p[:-1], p[1:])]
        self._middlepoints()

    def _flags(self, s, m, d):
        """mark_action_to_edges, environment16.py:381-393 (visit counts)."""
        f = np.zeros(self.E)
        for e in self.sp_edges[(s, m)]:
            f[e] += 1.0
        if m != d:
            for e in self.sp_edges[(m, d)]:
                f[e] += 1.0
        return f

    def _middlepoints(self):
        """compute_middlepoint_set_remove_rep_actions_no_loop, environment16.py:395-476."""
        self.mids = {}
        for n1 in range(self.N):
            for n2 in range(self.N):
                if n1 == n2:
                    continue
                kept, acc = [], []
                for m in range(self.K):
                    if m == n1:
                        continue
                    fl = self._flags(n1, m, n2)
                    if any(np.array_equal(fl, prev) for prev in acc):
                        continue
                    if m != n2:
                        p1 = self.paths[(n1, m)]
                        p2 = self.paths[(m, n2)]
                        cur = p1[:-1] + p2
                        if sum(1 for v in cur if v == n1 or v == n2) != 2:
                            continue
                    kept.append(m)
                    acc.append(fl)
                self.mids[(n1, n2)] = kept

    def route_edges(self, s, d, mid):
        """Edge ids (with multiplicity) of s->mid->d; mid in (None, d) = direct SP."""
        if mid is None or mid == d:
            return list(self.sp_edges[(s, d)])
        return list(self.sp_edges[(s, mid)]) + list(self.sp_edges[(mid, d)])

    def max_decisions(self, percentage=0.15):
        return int(math.ceil(self.N * (self.N - 1) * percentage))   # environment16.py:199


def _bfs_lexmin_paths(N, node_order, adj):
    INF = 1 << 30
    radj = {v: [] for v in range(N)}
    for u in node_order:
        for v in adj[u]:
            radj[v].append(u)
    dist = np.full((N, N), INF, dtype=np.int64)
    for d in range(N):
        dist[d, d] = 0
        frontier = [d]
        while frontier:
            nxt = []
            for v in frontier:
                for u in radj[v]:
                    if dist[u, d] == INF:
                        dist[u, d] = dist[v, d] + 1
                        nxt.append(u)
            frontier = nxt
    paths = {}
    for s in range(N):
        for d in range(N):
            if s == d:
                continue
            p, cur = [s], s
            while cur != d:
                cur = min(v for v in adj[cur] if dist[v, d] == dist[cur, d] - 1)
                p.append(cur)
            paths["%d:%d" % (s, d)] = p
    return paths


def _scan_max(loads, cap, edges):
    """Strict '>' scan from (0, 0, 0): environment16.py:604-612, 656-670."""
    best = (0, 0, 0)
    for e in range(len(loads)):
        u = loads[e] / cap[e]
        if u > best[2]:
            best = (edges[e][0], edges[e][1], u)
    return best


def _critical_demands(topo, tm, loads, crossing, percentage, num_critical=5):
    """environment16.py:672-674 + 188-200 (== environment15.py:537-538, 471-483).
    crossing[e] = ordered list of (s, d) whose current route crosses link e."""
    order = sorted(range(topo.E), key=lambda e: loads[e], reverse=True)[:num_critical]
    lst = []
    for e in order:
        for (s, d) in crossing[e]:
            t = (s, d, tm[s, d])
            if t not in lst:
                lst.append(t)
    lst = sorted(lst, key=lambda t: t[2], reverse=True)
    lim = int(math.ceil(topo.N * (topo.N - 1) * percentage))
    if len(lst) > lim:
        lst = lst[:lim]
    return lst


# ==========================================================================
# Env16: DRL middlepoint environment (environment16.py:582-725)
# ==========================================================================
class Env16:
    def __init__(self, topo: Topology, percentage=0.15):
        self.t = topo
        self.percentage = percentage
        self.edge_state = np.zeros((topo.E, 3))
        self.edge_state[:, 1] = topo.cap

    def reset(self, tm):
        """environment16.py:642-690."""
        t = self.t
        self.tm = np.asarray(tm, dtype=np.float64)
        self.sp_middlepoints = {}                  # ordered dict (s, d) -> mid
        loads = np.zeros(t.E)
        crossing = [[] for _ in range(t.E)]
        for s in range(t.N):                       # compute_link_utilization_reset :240-245
            for d in range(t.N):
                if s != d:
                    for e in t.sp_edges[(s, d)]:
                        loads[e] += self.tm[s, d]
                        crossing[e].append((s, d))
        self.loads = loads
        self.edge_max = _scan_max(loads, t.cap, t.edges)
        self.elig = _critical_demands(t, self.tm, loads, crossing, self.percentage)
        self.it = 0
        s, d, bw = self.elig[0]                    # _obtain_demand :266-271
        self.it = 1
        self.cur = (s, d)
        for e in t.sp_edges[(s, d)]:               # decrease_links_utilization_sp :685
            self.loads[e] -= self.tm[s, d]
        self.edge_state[:, 0] = self.loads
        self.edge_state[:, 2] = 0
        return self.tm[s, d], s, d

    def candidate_features(self, s, d):
        """mark_action_sp (environment16.py:711-725) + get_graph_features
        (script_eval_on_single_topology.py:131-160): list of [E,20] fp32, one per candidate."""
        t = self.t
        out = []
        util = (self.loads / t.cap).astype(np.float32)
        for m in t.mids[(s, d)]:
            mark = np.zeros(t.E)
            for e in t.route_edges(s, d, m):
                mark[e] = self.tm[s, d] / t.cap[e]          # assignment, not accumulation (Q9)
            ls = np.zeros((t.E, 20), dtype=np.float32)
            ls[:, 0] = util
            ls[:, 1] = t.link_capacity_feature
            ls[:, 2] = mark.astype(np.float32)
            out.append(ls)
        return out

    def critic_features(self):
        """critic_get_graph_features, train_Enero_3top_script.py:204-230."""
        t = self.t
        ls = np.zeros((t.E, 20), dtype=np.float32)
        ls[:, 0] = (self.loads / t.cap).astype(np.float32)
        ls[:, 1] = t.link_capacity_feature
        return ls

    def step(self, action):
        """environment16.py:582-640; returns the reference's 9-tuple."""
        t = self.t
        s, d = self.cur
        bw = self.tm[s, d]
        mid = t.mids[(s, d)][action]
        for e in t.sp_edges[(s, mid)]:
            self.loads[e] += bw
        if mid != d:
            for e in t.sp_edges[(mid, d)]:
                self.loads[e] += bw
            self.sp_middlepoints[(s, d)] = mid
        old = self.edge_max[2]
        self.edge_max = _scan_max(self.loads, t.cap, t.edges)
        reward = np.around((old - self.edge_max[2]) * 10, 2)
        done = False
        if self.it < len(self.elig):
            ns, nd, _ = self.elig[self.it]
            self.it += 1
        else:
            ns, nd = 1, 2
            done = True
        if (ns, nd) in self.sp_middlepoints:
            m2 = self.sp_middlepoints.pop((ns, nd))
            route = t.route_edges(ns, nd, m2)
        else:
            route = t.sp_edges[(ns, nd)]
        for e in route:
            self.loads[e] -= self.tm[ns, nd]
        self.cur = (ns, nd)
        self.edge_state[:, 0] = self.loads
        self.edge_state[:, 2] = 0
        return (reward, done, 0.0, self.tm[ns, nd], ns, nd, self.edge_max, 0.0,
                np.std(self.loads))


# ==========================================================================
# batching (Q1) + actor / critic forward (fp32)
# ==========================================================================
def batch_candidates(features, first, second):
    """pred_action_node_distrib_sp, script_eval_on_single_topology.py:105-120 with
    old_cummax (:58-64): graph k is offset by k*(max(index)+1), NOT by k*E."""
    k = len(features)
    E = features[0].shape[0]
    off_f = int(first.max()) + 1 if len(first) else 0
    off_s = int(second.max()) + 1 if len(second) else 0
    return {
        "link_state": np.concatenate(features, axis=0),
        "graph_id": np.repeat(np.arange(k, dtype=np.int32), E),
        "first": np.concatenate([first + i * off_f for i in range(k)]).astype(np.int32),
        "second": np.concatenate([second + i * off_s for i in range(k)]).astype(np.int32),
        "num_edges": k * E,
    }


def selu(x):
    """tf.nn.selu: x<0 ? scale*alpha*(exp(x)-1) : scale*x (fp32)."""
    x = x.astype(np.float32, copy=False)
    neg = (SELU_SCALE * SELU_ALPHA) * (np.exp(np.minimum(x, np.float32(0))) - np.float32(1))
    return np.where(x < 0, neg, SELU_SCALE * x).astype(np.float32)


def sigmoid(x):
    return (np.float32(1) / (np.float32(1) + np.exp(-x))).astype(np.float32)


def message_passing(w, link_state, first, second, num_edges, T=5):
    """Body of actorPPOmiddR.py:45-63 / criticPPO.py:41-58, fp32, op for op."""
    h = np.asarray(link_state, dtype=np.float32)
    wm, bm, kx, kh, b = (w["msg_kernel"], w["msg_bias"], w["gru_kernel"],
                         w["gru_recurrent_kernel"], w["gru_bias"])
    for _ in range(T):
        cat = np.concatenate([h[first], h[second]], axis=1)            # gather x2 + concat
        m = selu(cat @ wm + bm)                                         # Dense(20, selu)
        agg = np.zeros((num_edges, h.shape[1]), dtype=np.float32)
        np.add.at(agg, second, m)                                       # unsorted_segment_sum (in order)
        x = agg @ kx + b[0]                                             # GRUCell, reset_after=True
        y = h @ kh + b[1]
        z = sigmoid(x[:, :20] + y[:, :20])
        r = sigmoid(x[:, 20:40] + y[:, 20:40])
        hh = np.tanh(x[:, 40:] + r * y[:, 40:]).astype(np.float32)
        h = (z * h + (np.float32(1) - z) * hh).astype(np.float32)
    return h


def readout(w, g):
    a = selu(g @ w["r1_kernel"] + w["r1_bias"])
    a = selu(a @ w["r2_kernel"] + w["r2_bias"])
    return (a @ w["r3_kernel"] + w["r3_bias"]).astype(np.float32)


def actor_forward(w, link_state, graph_id, first, second, num_edges, T=5):
    """actorPPOmiddR.py:42-69 -> [G,1] logits."""
    h = message_passing(w, link_state, first, second, num_edges, T)
    G = int(graph_id.max()) + 1
    g = np.zeros((G, h.shape[1]), dtype=np.float32)
    np.add.at(g, graph_id, h)                                           # segment_sum
    return readout(w, g)


def critic_forward(w, link_state, first, second, num_edges, T=5):
    """criticPPO.py:38-64 -> [1,1] value."""
    h = message_passing(w, link_state, first, second, num_edges, T)
    g = h.sum(axis=0, keepdims=True, dtype=np.float32)
    return readout(w, g)


# --------------------------------------------------------------------------
# fp64 arbiter: the same formulas (actorPPOmiddR.py:45-69) evaluated in double precision on the fp32
# inputs and fp32 weights (both exactly representable).  Its rounding error (~1e-15) is far below any
# fp32 implementation's, so when two fp32 implementations (TensorFlow's Eigen kernels, this file's numpy
# restatement, the CUDA kernels) disagree in the last bits, the distance to this value says which is closer
# to what the reference's formula means.  Used by tests/ and tools/parity_report.py only.
# --------------------------------------------------------------------------
def _selu64(x):
    return np.where(x < 0, (float(SELU_SCALE) * float(SELU_ALPHA)) * np.expm1(np.minimum(x, 0.0)), float(SELU_SCALE) * x)


def message_passing64(w, link_state, first, second, num_edges, T=5):
    h = np.asarray(link_state, dtype=np.float64)
    wm, bm, kx, kh, b = (np.asarray(w[k], dtype=np.float64) for k in
                         ("msg_kernel", "msg_bias", "gru_kernel", "gru_recurrent_kernel", "gru_bias"))
    for _ in range(T):
        m = _selu64(np.concatenate([h[first], h[second]], axis=1) @ wm + bm)
        agg = np.zeros((num_edges, h.shape[1]))
        np.add.at(agg, second, m)
        x = agg @ kx + b[0]
        y = h @ kh + b[1]
        z = 1.0 / (1.0 + np.exp(-(x[:, :20] + y[:, :20])))
        r = 1.0 / (1.0 + np.exp(-(x[:, 20:40] + y[:, 20:40])))
        hh = np.tanh(x[:, 40:] + r * y[:, 40:])
        h = z * h + (1.0 - z) * hh
    return h


def readout64(w, g):
    a = _selu64(g @ np.asarray(w["r1_kernel"], np.float64) + np.asarray(w["r1_bias"], np.float64))
    a = _selu64(a @ np.asarray(w["r2_kernel"], np.float64) + np.asarray(w["r2_bias"], np.float64))
    return a @ np.asarray(w["r3_kernel"], np.float64) + np.asarray(w["r3_bias"], np.float64)


def actor_forward64(w, link_state, graph_id, first, second, num_edges, T=5):
    h = message_passing64(w, link_state, first, second, num_edges, T)
    G = int(graph_id.max()) + 1
    g = np.zeros((G, h.shape[1]))
    np.add.at(g, graph_id, h)
    return readout64(w, g)


def critic_forward64(w, link_state, first, second, num_edges, T=5):
    h = message_passing64(w, link_state, first, second, num_edges, T)
    return readout64(w, h.sum(axis=0, keepdims=True))


def softmax64(logits):
    x = np.asarray(logits, dtype=np.float64).reshape(-1)
    e = np.exp(x - x.max())
    return e / e.sum()


def softmax(logits):
    x = np.asarray(logits, dtype=np.float32).reshape(-1)
    e = np.exp(x - x.max())
    return (e / e.sum()).astype(np.float32)


def pred_action_distrib(w, env: Env16, s, d):
    t = env.t
    feats = env.candidate_features(s, d)
    b = batch_candidates(feats, t.first, t.second)
    logits = actor_forward(w, b["link_state"], b["graph_id"], b["first"], b["second"], b["num_edges"])
    return softmax(logits), b, logits.reshape(-1)


def play_middrout_greedy(w, env: Env16, tm, record=None):
    """play_middRout_games_sp, script_eval_on_single_topology.py:162-188."""
    _, s, d = env.reset(tm)
    best = env.edge_max[2]
    ospf = best
    best_routing = dict(env.sp_middlepoints)
    actions = []
    while True:
        probs, _, logits = pred_action_distrib(w, env, s, d)
        a = int(np.argmax(probs))
        actions.append(a)
        if record is not None:
            record.append({"s": s, "d": d, "probs": probs, "logits": logits})
        reward, done, _, _, s, d, mx, _, _ = env.step(a)
        if mx[2] < best:
            best = mx[2]
            best_routing = dict(env.sp_middlepoints)
        if done:
            break
    return best, ospf, best_routing, actions


# ==========================================================================
# Env15 + Local Search (environment15.py:485-544, 378-404;
# script_eval_on_single_topology.py:317-351, 395-435, 473-509)
# ==========================================================================
class Env15:
    def __init__(self, topo: Topology, percentage=0.15):
        self.t = topo
        self.percentage = percentage

    def _rebuild(self, tm, routing):
        t = self.t
        self.tm = np.asarray(tm, dtype=np.float64)
        loads = np.zeros(t.E)
        crossing = [dict() for _ in range(t.E)]
        for s in range(t.N):
            for d in range(t.N):
                if s != d:
                    for e in t.sp_edges[(s, d)]:
                        loads[e] += self.tm[s, d]
                        crossing[e][(s, d)] = 1
        for (s, d), m in routing.items():          # environment15.py:508-518 (insertion order)
            if m != d:
                for e in t.sp_edges[(s, d)]:
                    loads[e] -= self.tm[s, d]
                    crossing[e].pop((s, d), None)
                for e in t.route_edges(s, d, m):
                    loads[e] += self.tm[s, d]
                    crossing[e][(s, d)] = 1
        self.loads = loads
        return crossing

    def reset_DRL_hill_sp(self, tm, best_routing):
        """environment15.py:485-544; returns -maxUti."""
        self.sp_middlepoints = dict(best_routing)
        crossing = self._rebuild(tm, self.sp_middlepoints)
        self.edge_max = _scan_max(self.loads, self.t.cap, self.t.edges)
        self.elig = _critical_demands(self.t, self.tm, self.loads,
                                      [list(c.keys()) for c in crossing], self.percentage)
        return -self.edge_max[2]

    def reset_hill_sp(self, tm):
        """environment15.py:438-469: all demands on direct SPs; every pair eligible."""
        self.sp_middlepoints = {}
        self._rebuild(tm, {})
        self.edge_max = _scan_max(self.loads, self.t.cap, self.t.edges)
        self.elig = [(s, d, self.tm[s, d]) for s in range(self.t.N) for d in range(self.t.N) if s != d]
        return -self.edge_max[2]

    def move_value(self, s, d, action):
        """get_value_sp after de-allocating the current route
        (script_eval_on_single_topology.py:404-434, 317-351)."""
        t = self.t
        bw = self.tm[s, d]
        loads = self.loads.copy()
        for e in t.route_edges(s, d, self.sp_middlepoints.get((s, d))):
            loads[e] -= bw
        for e in t.route_edges(s, d, t.mids[(s, d)][action]):
            loads[e] += bw
        cur = -1000000
        for e in range(t.E):
            u = loads[e] / t.cap[e]
            if u > cur:
                cur = u
        return -cur

    def explore(self):
        """explore_neighbourhood_DRL_sp / explore_neighbourhood_sp: first strictly best."""
        best, state = -1000000, None
        for (s, d, _) in self.elig:
            for a in range(len(self.t.mids[(s, d)])):
                v = self.move_value(s, d, a)
                if v > best:
                    best, state = v, (a, s, d)
        return best, state

    def apply(self, action, s, d):
        """de-allocate (script :494-502) + step_hill_sp (environment15.py:378-404)."""
        t = self.t
        bw = self.tm[s, d]
        for e in t.route_edges(s, d, self.sp_middlepoints.pop((s, d), None)):
            self.loads[e] -= bw
        mid = t.mids[(s, d)][action]
        for e in t.route_edges(s, d, mid):
            self.loads[e] += bw
        if mid != d:
            self.sp_middlepoints[(s, d)] = mid
        self.edge_max = _scan_max(self.loads, t.cap, t.edges)
        return -self.edge_max[2]


def local_search(env: Env15, current):
    """play_DRL_GNN_sp_hill_climbing_games loop, script :482-509; returns
    (final maxUti, move list)."""
    moves = []
    while True:
        nxt, state = env.explore()
        if nxt <= current or abs((-1) * nxt - (-1) * current) < 1e-4:
            break
        a, s, d = state
        current = env.apply(a, s, d)
        moves.append((a, s, d, -current))
    return current * (-1), moves


# ==========================================================================
# SAP baseline (environment20.py:124-196, 240-316; script_eval_on_single_topology.py:511-557)
# ==========================================================================
def k_shortest_simple_paths(topo: Topology, K=5, exhaustive=True):
    """Env20.compute_SPs, environment20.py:165-196 -> {"s:d": [node lists]}.  exhaustive=True restates the
    reference literally (every simple path of <= 2*diameter links, sorted by (len, nodes), first K: exponential, small
    graphs only); exhaustive=False reaches the same list through networkx's Yen generator (paths by length; every
    path as short as the K-th is collected before sorting)."""
    import itertools
    import networkx as nx
    G = nx.DiGraph()
    G.add_nodes_from(range(topo.N))
    G.add_edges_from(topo.edges)
    cutoff = 2 * nx.diameter(G)
    out = {}
    for n1 in range(topo.N):
        for n2 in range(topo.N):
            if n1 == n2:
                continue
            if exhaustive:
                paths = [p for p in nx.all_simple_paths(G, source=n1, target=n2, cutoff=cutoff)]
            else:
                paths, gen = [], nx.shortest_simple_paths(G, n1, n2)
                for p in gen:
                    if len(p) - 1 > cutoff or (len(paths) >= K and len(p) > len(paths[K - 1])):
                        break
                    paths.append(p)
                    paths.sort(key=lambda item: (len(item), item))
            out["%d:%d" % (n1, n2)] = sorted(paths, key=lambda item: (len(item), item))[:K]
    return out


def play_sap(topo: Topology, all_paths, tm, K=5, seed=9):
    """play_sap_games (script :543-557) with SAPAgent.act (:515-541) on Env20.reset / step (environment20.py:300-316,
    240-291) -> (final max link utilisation, list of the agent's actions, -1 = no path had room)."""
    import random
    rnd = random.Random(seed)                      # env.seed(SEED): random.seed(9), environment20.py:98-100
    tm = np.asarray(tm, dtype=np.float64)
    demands = [(s, d, tm[s, d]) for s in range(topo.N) for d in range(topo.N) if s != d]
    demands = sorted(demands, key=lambda tup: tup[2], reverse=True)            # :149
    loads = np.zeros(topo.E)
    actions = []
    for (s, d, bw) in demands:
        path_list = all_paths["%d:%d" % (s, d)]
        action = -1
        for pi, path in enumerate(path_list[:K]):
            if all(not ((loads[topo.eid[(a, b)]] + bw) / topo.cap[topo.eid[(a, b)]] > 1) for a, b in zip(path[:-1], path[1:])):
                action = pi
                break
        actions.append(action)
        if action == -1:
            action = rnd.randint(0, len(path_list) - 1)                          # :253
        path = path_list[action]
        for a, b in zip(path[:-1], path[1:]):
            loads[topo.eid[(a, b)]] += bw
    return _scan_max(loads, topo.cap, topo.edges)[2], actions


# ==========================================================================
# PPO pieces (train_Enero_3top_script.py:264-377)
# ==========================================================================
def get_advantages(values, masks, rewards, gamma=0.99, lmbda=0.95):
    """train_Enero_3top_script.py:367-377.  The reference runs under numpy 1.19.5
    (requirements.txt:41), where `python float * np.float32 scalar` promotes to float64; under
    numpy >= 2 (NEP 50) the same expression would stay float32, so the critic values are widened
    explicitly to keep the reference's arithmetic."""
    values = [np.float64(v) for v in values]
    returns = []
    gae = 0
    for i in reversed(range(len(rewards))):
        delta = rewards[i] + gamma * values[i + 1] * masks[i] - values[i]
        gae = delta + gamma * lmbda * masks[i] * gae
        returns.insert(0, gae + values[i])
    adv = np.array(returns) - values[:-1]
    return returns, (adv - np.mean(adv)) / (np.std(adv) + 1e-10)
