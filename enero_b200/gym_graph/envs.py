"""``Env16`` / ``Env15`` with the reference's method and attribute surface
(environment16.py / environment15.py), state kept on the GPU by ``EpisodeBatch`` (B = 1) and
mirrored into the numpy attributes the reference's agents read (``edge_state``,
``src_dst_k_middlepoints``, ``edgeMaxUti``, ``list_eligible_demands``, ``sp_middlepoints`` ...).
"""
from __future__ import annotations

import os
import random

import numpy as np

from .. import datasets as ds
from ..engine import EpisodeBatch
from ..runtime import default_context
from ..topology import Topology


class _Graph:
    """The little of ``defoDatasetAPI.Gbase`` the eval scripts touch (``len(Gbase.edges())``,
    script_eval_on_single_topology.py:633)."""

    def __init__(self, topo):
        self._t = topo

    def edges(self):
        return list(self._t.edges)

    def nodes(self):
        return list(range(self._t.N))

    # `for i in env.graph: for j in env.graph[i]` walks the links in edge-id order
    # (environment16.py:557-569, script_eval_on_single_topology.py:334-339)
    def __iter__(self):
        return iter(self._t.graph_file.node_order)

    def __getitem__(self, i):
        return list(self._t.graph_file.adj[i])


class _DatasetAPI:
    def __init__(self, topo, gf):
        self.Gbase = _Graph(topo)
        self.links_bw = [dict() for _ in range(gf.num_nodes)]
        self.links_weight = [dict() for _ in range(gf.num_nodes)]
        for (i, j), c in gf.links_bw.items():
            self.links_bw[i][j] = c
            self.links_weight[i][j] = gf.links_weight[(i, j)]
        self.net_size = gf.num_nodes

    def _get_traffic_matrix(self, traffic_file):
        return ds.parse_demands_file(traffic_file, self.net_size)


class _EnvBase:
    def __init__(self, context=None):
        self._ctx = context
        self.top_K_critical_demands = False
        self.num_critical_links = 5
        self.episode_over = True
        self.reward = 0
        self.sp_middlepoints = None
        self.sp_middlepoints_step = dict()
        self.list_eligible_demands = None
        self.edgeMaxUti = None
        self.TM = None

    def seed(self, seed):
        random.seed(seed)
        np.random.seed(seed)

    @property
    def context(self):
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    def _generate(self, dataset_folder_name, graph_topology_name, EPISODE_LENGTH, K, X, state_cols):
        self.episode_length = EPISODE_LENGTH
        self.graph_topology_name = graph_topology_name
        self.dataset_folder_name = dataset_folder_name
        self.list_eligible_demands = list()
        self.percentage_demands = X
        self.topology = Topology.from_dataset(dataset_folder_name, graph_topology_name, K=K)
        t = self.topology
        self.defoDatasetAPI = _DatasetAPI(t, t.graph_file)
        self.links_bw = self.defoDatasetAPI.links_bw
        self.graph = self.defoDatasetAPI.Gbase
        self.numNodes, self.numEdges, self.K = t.N, t.E, t.K
        self.maxCapacity = float(t.cap.max())
        self.edgesDict = dict(t.edgesDict)
        self.edge_state = np.zeros((t.E, state_cols))
        self.edge_state[:, 1] = t.cap
        self.first, self.second = t.first, t.second
        self.firstTrueSize = t.P
        self.link_capacity_feature = t.link_capacity_feature
        self.nodes = list(range(t.N))
        self.src_dst_k_middlepoints = t.middlepoint_dict()
        self.shortest_paths = np.zeros((t.N, t.N), dtype=object)
        for s in range(t.N):
            for d in range(t.N):
                if s != d:
                    self.shortest_paths[s, d] = t.path_nodes(s, d)
        self._batch = EpisodeBatch(self.context, t, 1, X)

    def _load_tm(self, tm_id):
        tm_file = os.path.join(self.dataset_folder_name, "TM", "%s.%s.demands" % (self.graph_topology_name, tm_id))
        self.TM = ds.parse_demands_file(tm_file, self.numNodes)
        return self.TM

    def _edge_tuple(self, edge_id, value):
        if edge_id < 0:
            return (0, 0, 0)
        i, j = self.topology.edges[edge_id]
        return (i, j, value)

    def _path_edges(self, src, dst):
        return self.topology.sp_edges(src, dst)


class Env16(_EnvBase):
    """GraphEnv-v16 (environment16.py:16-725)."""

    def generate_environment(self, dataset_folder_name, graph_topology_name, EPISODE_LENGTH, K, X):
        self._generate(dataset_folder_name, graph_topology_name, EPISODE_LENGTH, K, X, 3)
        self.iter_list_elig_demn = 0

    def reset(self, tm_id):
        """environment16.py:642-690 -> (TM[s][d], s, d)."""
        if not self.top_K_critical_demands:
            raise NotImplementedError("enero_b200 implements the ENERO setting top_K_critical_demands=True "
                                      "(train_Enero_3top_script.py:399, script_eval_on_single_topology.py:603)")
        tm = self._load_tm(tm_id)
        self._batch.reset(tm[None])
        elig, n = self._batch.eligible()
        self.list_eligible_demands = [(int(s), int(d), tm[s, d]) for s, d in elig[0, :n[0]]]
        self.sp_middlepoints = dict()
        self.iter_list_elig_demn = 1
        st = self._batch.state()
        # edgeMaxUti is taken before the first demand is removed (environment16.py:656-670)
        self.edgeMaxUti = self._edge_tuple(int(st["max_edge"][0]), st["max_uti"][0])
        self.currentVal = -self.edgeMaxUti[2]
        self.initial_maxLinkUti = -self.edgeMaxUti[2]
        s, d = int(st["cur"][0, 0]), int(st["cur"][0, 1])
        self.patMaxBandwth = (s, d, int(tm[s, d]))
        self.edge_state[:, 0] = st["loads"][0]
        self.edge_state[:, 2] = 0
        self.episode_over = False
        return tm[s, d], s, d

    def step(self, action, demand, source, destination):
        """environment16.py:582-640 -> the reference's 9-tuple."""
        if (source, destination) != (self.patMaxBandwth[0], self.patMaxBandwth[1]):
            raise ValueError("step() must be called with the demand returned by reset()/step()")
        action = int(action)
        mids = self.src_dst_k_middlepoints["%d:%d" % (source, destination)]
        if not 0 <= action < len(mids):
            raise IndexError("list index out of range")
        reward, done, st = self._batch.step_state([action])
        mid = mids[action]
        if mid != destination:
            self.sp_middlepoints["%d:%d" % (source, destination)] = mid
        self.sp_middlepoints_step = self.sp_middlepoints
        self.edgeMaxUti = self._edge_tuple(int(st["max_edge"][0]), st["max_uti"][0])
        self.currentVal = -self.edgeMaxUti[2]
        self.reward = reward[0]
        self.episode_over = bool(done[0])
        ns, nd = int(st["cur"][0, 0]), int(st["cur"][0, 1])
        if not self.episode_over:
            self.iter_list_elig_demn += 1
        self.patMaxBandwth = (ns, nd, int(self.TM[ns][nd]))
        self.sp_middlepoints.pop("%d:%d" % (ns, nd), None)
        self.edge_state[:, 0] = st["loads"][0]
        self.edge_state[:, 2] = 0
        return (self.reward, self.episode_over, 0.0, self.TM[ns][nd], ns, nd, self.edgeMaxUti, 0.0,
                np.std(self.edge_state[:, 0]))

    def run_greedy_episode(self):
        """The whole greedy loop of play_middRout_games_sp (script_eval_on_single_topology.py:169-186) for the
        traffic matrix of the last reset(), device resident: one CUDA-graph replay per decision, no host round
        trip until the episode is over.  The actor weights are whatever the context holds (the agent binds its
        model first).  -> dict(actions, rewards, max_uti per step, best, best_routing {"s:d": mid})."""
        eb = self._batch
        eb.run_greedy()
        st = eb.state()
        n = int(st["steps"][0])
        act, rew, mx = eb.history()
        best, ospf, routing = eb.best()
        # leave the attributes the way the reference's loop leaves them after the last step()
        self.sp_middlepoints = {}
        for k in range(n):
            s, d, _ = self.list_eligible_demands[k]
            mid = self.src_dst_k_middlepoints["%d:%d" % (s, d)][int(act[0, k])]
            nxt = self.list_eligible_demands[k + 1][:2] if k + 1 < n else (1, 2)
            if mid != d:
                self.sp_middlepoints["%d:%d" % (s, d)] = mid
            self.sp_middlepoints.pop("%d:%d" % nxt, None)
        self.sp_middlepoints_step = self.sp_middlepoints
        self.edgeMaxUti = self._edge_tuple(int(st["max_edge"][0]), st["max_uti"][0])
        self.currentVal = -self.edgeMaxUti[2]
        self.iter_list_elig_demn = n
        self.episode_over = True
        self.patMaxBandwth = (1, 2, int(self.TM[1][2]))
        self.edge_state[:, 0] = st["loads"][0]
        self.edge_state[:, 2] = 0
        return {"actions": act[0, :n].copy(), "rewards": rew[0, :n].copy(), "max_uti": mx[0, :n].copy(),
                "best": float(best[0]), "ospf": float(ospf[0]),
                "best_routing": {"%d:%d" % (s, d): int(m) for s, d, m in routing[0]}}

    def mark_action_sp(self, src, dst, init_source, final_destination):
        """environment16.py:711-725 -- host-side mirror for agents that still build their own
        features; the fused GPU path (enero_decide_batch) marks candidates on device."""
        bw = self.TM[init_source][final_destination]
        for e in self._path_edges(src, dst):
            self.edge_state[e][2] = bw / self.edge_state[e][1]


class Env15(_EnvBase):
    """GraphEnv-v15 (environment15.py:14-563): the Local-Search environment."""

    def generate_environment(self, dataset_folder_name, graph_topology_name, EPISODE_LENGTH, K, percentage_demands):
        self._generate(dataset_folder_name, graph_topology_name, EPISODE_LENGTH, K, percentage_demands, 2)
        self._ls = None

    def _routing_list(self, best_routing):
        out = []
        for key, mid in (best_routing or {}).items():
            s, d = (int(v) for v in key.split(":"))
            out.append((s, d, int(mid)))
        return out

    def reset_DRL_hill_sp(self, tm_id, best_routing, list_of_demands_to_change):
        """environment15.py:485-544 -> -maxUti of the restored DRL routing.  The hill climbing itself
        (script_eval_on_single_topology.py:482-509) runs on device in one launch: local_search(); callers that
        drive it move by move use allocate_to_destination_sp / decrease_links_utilization_sp / step_hill_sp."""
        tm = self._load_tm(tm_id)
        self.sp_middlepoints = dict(best_routing) if best_routing is not None else dict()
        self.list_of_demands_to_change = list_of_demands_to_change
        self._upload_tm(tm)
        self._mode, self._routing = 0, [self._routing_list(best_routing)]
        return self._reset_only(tm, True)

    def reset_hill_sp(self, tm_id):
        """environment15.py:438-469 -> -maxUti on plain shortest paths; all demands eligible."""
        tm = self._load_tm(tm_id)
        self.sp_middlepoints = dict()
        self._upload_tm(tm)
        self._mode, self._routing = 1, None
        return self._reset_only(tm, False)

    def _reset_only(self, tm, critical):
        r = self._batch.local_search(self._routing, mode=self._mode, max_moves=0, want_moves=False)
        if critical:
            elig, n = self._batch.ls_eligible()
            self.list_eligible_demands = [(int(s), int(d), tm[s, d]) for s, d in elig[0, :n[0]]]
        else:
            self.list_eligible_demands = [(s, d, tm[s, d]) for s in range(self.numNodes) for d in range(self.numNodes) if s != d]
        self.edge_state[:, 0] = self._batch.ls_loads()[0]
        e = int(self._batch.ls_max_edge()[0])
        self.edgeMaxUti = self._edge_tuple(e, r["start"][0])
        self._ls = None
        return -r["start"][0]

    # ---- per-operation methods (environment15.py:546-563, 150-165, 378-404): each one is a device call on the
    # episode's Local-Search loads, mirrored into edge_state[:, 0] -------------------------------------------
    def _route_add(self, src, dst, value):
        loads, mx, me = self._batch.ls_route_add(0, int(src), int(dst), float(value))
        self.edge_state[:, 0] = loads
        return mx, me

    def allocate_to_destination_sp(self, src, dst, init_source, final_destination):
        self._route_add(src, dst, self.TM[init_source][final_destination])

    def decrease_links_utilization_sp(self, src, dst, init_source, final_destination):
        self._route_add(src, dst, -self.TM[init_source][final_destination])

    def step_hill_sp(self, action, source, destination):
        mid = list(self.src_dst_k_middlepoints["%d:%d" % (source, destination)])[action]
        mx, me = self._route_add(source, mid, self.TM[source][destination])
        if mid != destination:
            mx, me = self._route_add(mid, destination, self.TM[source][destination])
            self.sp_middlepoints["%d:%d" % (source, destination)] = mid
        self.edgeMaxUti = self._edge_tuple(int(me), mx)
        return -self.edgeMaxUti[2]

    def _upload_tm(self, tm):
        # Env16-style reset puts the TM on the device; its demand list is not used by the LS
        self._batch.reset(tm[None])

    def local_search(self):
        """-> (final maxUti, [(action, s, d, maxUti after the move), ...]) of the hill climbing
        started by the last reset_*; updates sp_middlepoints / edgeMaxUti like the reference loop."""
        if self._ls is None:
            self._ls = self._batch.local_search(self._routing, mode=self._mode)
        ls = self._ls
        n = int(ls["n_moves"][0])
        moves = []
        for k in range(n):
            a, s, d = (int(v) for v in ls["moves"][0, k])
            mid = self.src_dst_k_middlepoints["%d:%d" % (s, d)][a]
            self.sp_middlepoints.pop("%d:%d" % (s, d), None)
            if mid != d:
                self.sp_middlepoints["%d:%d" % (s, d)] = mid
            moves.append((a, s, d, float(ls["move_val"][0, k])))
        return float(ls["final"][0]), moves


class Env20(_EnvBase):
    """GraphEnv-v20 (environment20.py:14-316): the environment of the SAP baseline.  K is fixed to 5 shortest simple
    paths per pair (environment20.py:221); ``allPaths`` is the reference's table.  The episode itself is one kernel
    over the whole sorted demand list (``run_sap``): SAPAgent.act + step for all N(N-1) demands, the per-step Python
    round trip of play_sap_games (script :543-557) folded into the device loop."""

    def generate_environment(self, dataset_folder_name, graph_topology_name, EPISODE_LENGTH, K):
        self._generate(dataset_folder_name, graph_topology_name, EPISODE_LENGTH, K, 0.15, 3)
        self.K = 5
        self.allPaths = self.topology.k_shortest_paths(self.K)
        self._seed = 9

    def seed(self, seed):
        super().seed(seed)
        self._seed = int(seed)

    def reset(self, tm_id):
        """environment20.py:300-316 -> (TM[s][d], s, d) of the largest demand."""
        tm = self._load_tm(tm_id)
        n = self.numNodes
        self.list_eligible_demands = sorted(((s, d, tm[s, d]) for s in range(n) for d in range(n) if s != d),
                                            key=lambda tup: tup[2], reverse=True)
        self.iter_list_elig_demn = 1
        s, d, bw = self.list_eligible_demands[0]
        self.patMaxBandwth = (s, d, int(bw))
        return tm[s, d], s, d

    def run_sap(self):
        """The whole play_sap_games loop for the traffic matrix of the last reset -> (maxLinkUti[2], actions)."""
        mx, acts = self._batch.sap(self.TM[None], K=self.K, seed=self._seed, want_actions=True)
        return float(mx[0]), acts[0]

    def step(self, action, demand, source, destination):
        raise NotImplementedError("Env20 runs whole SAP episodes on the device: use run_sap() "
                                  "(agents.play_sap_games does) instead of act()/step() one demand at a time")
