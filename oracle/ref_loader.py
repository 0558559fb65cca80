"""TEST INFRASTRUCTURE ONLY -- imports the reference's own environment and
Local-Search code, unmodified, from /root/reference (build container only).

Used by oracle/gen_golden.py to produce tests/golden/*; nothing in the product
path, the GPU tests, smoke() or bench.py touches this module (the reference
tree does not exist on the GPU box).

Recipe = SURVEY.md Appendix B: stub gym / matplotlib.pyplot / tensorflow,
patch networkx.edge_betweenness_centrality (networkx >= 3 keys multigraph edges
by (u, v, key), environment16.py:566 indexes by (u, v); the value is unused),
and lift the episode / Local-Search drivers out of
script_eval_on_single_topology.py by AST so they run verbatim.
"""
import ast
import os
import sys
import time as tt
from collections import defaultdict

REFERENCE_ROOT = os.environ.get("ENERO_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))

_LIFTED = ("HILL_CLIMBING", "play_middRout_games_sp",
           "play_DRL_GNN_sp_hill_climbing_games", "play_sp_hill_climbing_games",
           "SAPAgent", "play_sap_games")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gym-graph"))


def _setup_path():
    for p in (os.path.join(REFERENCE_ROOT, "gym-graph"), REFERENCE_ROOT,
              os.path.join(_HERE, "stubs")):
        if p not in sys.path:
            sys.path.insert(0, p)


def load_gym():
    """Returns the stub ``gym`` with the reference's GraphEnv-v15/16/20 registered."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _setup_path()
    import networkx as nx
    nx.edge_betweenness_centrality = lambda G, *a, **k: defaultdict(float)
    import gym
    import gym_graph  # noqa: F401  (registers the envs)
    return gym


def load_eval_functions(dataset_folder, topology_name, percentage_demands=0.15):
    """AST-lift the drivers from script_eval_on_single_topology.py:162-188,313-509
    and bind the module globals they read (lines 27-41 of that script)."""
    import numpy as np
    gym = load_gym()
    src_path = os.path.join(REFERENCE_ROOT, "script_eval_on_single_topology.py")
    tree = ast.parse(open(src_path).read())
    body = [n for n in tree.body
            if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in _LIFTED]
    mod = ast.Module(body=body, type_ignores=[])
    glb = {
        "gym": gym, "np": np, "tt": tt,
        "ENV_SIMM_ANEAL_AGENT": "GraphEnv-v15", "ENV_SAP_AGENT": "GraphEnv-v20", "SEED": 9,
        "general_dataset_folder": dataset_folder,
        "graph_topology_name": topology_name,
        "EPISODE_LENGTH_MIDDROUT": 100, "NUM_ACTIONS": 100,
        "percentage_demands": percentage_demands,
    }
    exec(compile(mod, src_path, "exec"), glb)
    return glb


def make_env16(dataset_folder, topology_name, K=100, percentage=0.15):
    gym = load_gym()
    env = gym.make("GraphEnv-v16")
    env.seed(9)
    env.generate_environment(dataset_folder, topology_name, 100, K, percentage)
    env.top_K_critical_demands = True
    return env


def make_env15(dataset_folder, topology_name, K=100, percentage=0.15):
    gym = load_gym()
    env = gym.make("GraphEnv-v15")
    env.seed(9)
    env.generate_environment(dataset_folder, topology_name, 100, K, percentage)
    return env
