"""REPETITA-format dataset ingest and the synthetic "Zoo-shaped" generators.

File formats are the drop-in input of the reference and are kept unchanged:

* ``<folder>/<topo>.graph``  -- parsed like ``defo_process_results.py:55-86``
  (``NODES n`` / header / n node lines / then every line starting with
  ``Link_`` or ``edge_``: ``label src dst weight bw delay``); writer layout as
  ``generate_link_failure_topologies.py:40-56``.
* ``<folder>/TM/<topo>.<id>.demands`` -- two header lines then
  ``label src dst bw``; bandwidths are floored (``defo_process_results.py:216-225``).
* ``<folder>/shortest_paths.json`` -- optional cache ``{"s:d": [nodes]}``
  read by ``environment16.py:483-500``.

The synthetic generators follow SURVEY.md section 8(d): ``zoo_like(N, L, seed)``
is a random recursive tree plus ``L-(N-1)`` extra distinct links; ``uniform_tm``
draws ``floor(U(0.5,1.0)*sigma)`` per ordered pair.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass

import numpy as np


@dataclass
class GraphFile:
    """Result of parsing a ``.graph`` file (insertion-ordered)."""

    num_nodes: int
    node_order: list            # nodes in first-appearance order (src before dst)
    adj: dict                   # node -> list of out-neighbours in file order
    links_bw: dict              # (src, dst) -> float capacity
    links_weight: dict          # (src, dst) -> int weight

    def edge_list(self):
        """Directed links in edge-id order (``environment16.py:557-569``)."""
        out = []
        for i in self.node_order:
            for j in self.adj[i]:
                out.append((i, j))
        return out


def parse_graph_file(path: str) -> GraphFile:
    """Same accept/skip rules as ``defo_process_results.py:55-86``."""
    with open(path) as fd:
        line = fd.readline()
        net_size = int(line.split(" ")[1])
        fd.readline()
        for _ in range(net_size):
            fd.readline()
        node_order, adj, bw, wt = [], {}, {}, {}
        for line in fd:
            if not line.startswith("Link_") and not line.startswith("edge_"):
                continue
            c = line.split(" ")
            src, dst, weight, cap = int(c[1]), int(c[2]), int(c[3]), float(c[4])
            for n in (src, dst):
                if n not in adj:
                    adj[n] = []
                    node_order.append(n)
            if dst not in adj[src]:
                # MultiDiGraph.add_edge on an existing pair adds a parallel key but
                # the adjacency (what the env iterates) keeps one entry per neighbour
                adj[src].append(dst)
            bw[(src, dst)] = cap
            wt[(src, dst)] = weight
    return GraphFile(net_size, node_order, adj, bw, wt)


def parse_demands_file(path: str, num_nodes: int) -> np.ndarray:
    """``defo_process_results.py:216-225``: fp64 matrix of floored demands."""
    tm = np.zeros((num_nodes, num_nodes))
    with open(path) as fd:
        fd.readline()
        fd.readline()
        for line in fd:
            c = line.split(" ")
            if len(c) < 4:
                continue
            tm[int(c[1]), int(c[2])] = np.floor(float(c[3]))
    return tm


def parse_demands_files(paths, num_nodes: int, threads: int = 0) -> np.ndarray:
    """many ``.demands`` files -> fp64 [n, N, N], parsed by the library's multi-threaded C++ reader
    (same accept rules as ``parse_demands_file``)."""
    import ctypes as C

    from . import _lib
    lib = _lib.load()
    paths = [os.fsencode(p) for p in paths]
    out = np.zeros((len(paths), num_nodes, num_nodes))
    if paths:
        arr = (C.c_char_p * len(paths))(*paths)
        _lib.check(lib.enero_parse_demands_batch(arr, len(paths), int(num_nodes), _lib.ptr(out), int(threads)))
    return out


# --------------------------------------------------------------------------
# synthetic generators
# --------------------------------------------------------------------------

def zoo_like(num_nodes: int, num_links: int, seed: int):
    """Random recursive tree + extra distinct links; returns sorted undirected
    links ``[(a, b)]`` with ``a < b`` and a capacity per link in {10000, 40000}."""
    rng = np.random.default_rng(seed)
    links = set()
    for v in range(1, num_nodes):
        u = int(rng.integers(0, v))
        links.add((u, v))
    max_links = num_nodes * (num_nodes - 1) // 2
    num_links = min(num_links, max_links)
    while len(links) < num_links:
        a = int(rng.integers(0, num_nodes))
        b = int(rng.integers(0, num_nodes))
        if a == b:
            continue
        links.add((min(a, b), max(a, b)))
    links = sorted(links)
    caps = [float(rng.choice([10000, 40000])) for _ in links]
    return links, caps


def write_graph_file(path: str, num_nodes: int, links, caps) -> None:
    """Both directions of every undirected link consecutively, weight 1, delay 100."""
    with open(path, "w") as fd:
        fd.write("NODES %d\n" % num_nodes)
        fd.write("label x y\n")
        for n in range(num_nodes):
            fd.write("%d_Node 0.0 0.0\n" % n)
        fd.write("\n")
        fd.write("EDGES %d\n" % (2 * len(links)))
        fd.write("label src dest weight bw delay\n")
        k = 0
        for (a, b), c in zip(links, caps):
            fd.write("edge_%d %d %d 1 %d 100\n" % (k, a, b, int(c)))
            fd.write("edge_%d %d %d 1 %d 100\n" % (k + 1, b, a, int(c)))
            k += 2


def link_failure_variants(links, caps, num_nodes: int, link_failures: int, num_variants: int, seed: int):
    """generate_link_failure_topologies.py:123-177: ``num_variants`` distinct copies of the topology with
    ``link_failures`` random bidirectional links removed, keeping only connected graphs; the surviving
    links keep their order (so edge ids of a variant follow the same insertion rule as the base graph).
    Returns ``[(links, caps, removed)]``.  The reference draws with ``random.choice``; this generator is
    seeded separately (offline data preparation, not on the parity path)."""
    rng = np.random.default_rng(seed)
    out, seen = [], set()
    links = list(links)
    if link_failures >= len(links):
        raise ValueError("cannot remove %d of %d links" % (link_failures, len(links)))
    attempts = 0
    while len(out) < num_variants:
        attempts += 1
        if attempts > 200 * max(1, num_variants):
            raise RuntimeError("could not find %d connected, distinct variants" % num_variants)
        drop = tuple(sorted(int(i) for i in rng.choice(len(links), size=link_failures, replace=False)))
        if drop in seen:
            continue
        keep = [i for i in range(len(links)) if i not in drop]
        # connectivity of the undirected remainder (nx.is_connected in the reference)
        adj = {v: [] for v in range(num_nodes)}
        for i in keep:
            a, b = links[i]
            adj[a].append(b)
            adj[b].append(a)
        stack, reached = [0], {0}
        while stack:
            v = stack.pop()
            for u in adj[v]:
                if u not in reached:
                    reached.add(u)
                    stack.append(u)
        if len(reached) != num_nodes:
            continue
        seen.add(drop)
        out.append(([links[i] for i in keep], [caps[i] for i in keep], [links[i] for i in drop]))
    return out


def uniform_tm(num_nodes: int, sigma: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    tm = np.floor(rng.uniform(0.5, 1.0, size=(num_nodes, num_nodes)) * sigma)
    np.fill_diagonal(tm, 0.0)
    return tm


def write_demands_file(path: str, tm: np.ndarray) -> None:
    n = tm.shape[0]
    with open(path, "w") as fd:
        fd.write("DEMANDS %d\n" % (n * (n - 1)))
        fd.write("label src dest bw\n")
        k = 0
        for s in range(n):
            for d in range(n):
                if s != d:
                    fd.write("demand_%d %d %d %d\n" % (k, s, d, int(tm[s, d])))
                    k += 1


def bfs_lexmin_paths(gf: GraphFile):
    """Hop-count shortest, lexicographically smallest node sequence for every
    ordered pair -- the closed form of ``environment16.py:491-495``."""
    n = gf.num_nodes
    INF = 1 << 30
    dist = [[INF] * n for _ in range(n)]
    radj = {v: [] for v in range(n)}
    for u in gf.node_order:
        for v in gf.adj[u]:
            radj[v].append(u)
    for d in range(n):
        dist[d][d] = 0
        frontier = [d]
        while frontier:
            nxt = []
            for v in frontier:
                for u in radj.get(v, []):
                    if dist[u][d] == INF:
                        dist[u][d] = dist[v][d] + 1
                        nxt.append(u)
            frontier = nxt
    paths = {}
    for s in range(n):
        for d in range(n):
            if s == d:
                continue
            if dist[s][d] >= INF:
                raise ValueError("graph is not strongly connected (%d->%d)" % (s, d))
            p = [s]
            cur = s
            while cur != d:
                cur = min(v for v in gf.adj[cur] if dist[v][d] == dist[cur][d] - 1)
                p.append(cur)
            paths["%d:%d" % (s, d)] = p
    return paths


def sigma_for_mean_utilisation(gf: GraphFile, target: float = 0.5) -> float:
    """Scale so that mean shortest-path link utilisation is about ``target``."""
    paths = bfs_lexmin_paths(gf)
    cross = {}
    for p in paths.values():
        for a, b in zip(p[:-1], p[1:]):
            cross[(a, b)] = cross.get((a, b), 0) + 1
    acc = 0.0
    for e, cap in gf.links_bw.items():
        acc += 0.75 * cross.get(e, 0) / cap
    mean_per_sigma = acc / max(1, len(gf.links_bw))
    return target / mean_per_sigma


def make_synthetic_dataset(folder: str, name: str, num_nodes: int, num_links: int,
                           topo_seed: int, num_tms: int, tm_seed0: int = 2000,
                           write_sp_cache: bool = True):
    """Emit ``<folder>/<name>.graph``, ``<folder>/TM/<name>.<j>.demands`` and
    (optionally) ``<folder>/shortest_paths.json``; returns the parsed GraphFile."""
    os.makedirs(os.path.join(folder, "TM"), exist_ok=True)
    links, caps = zoo_like(num_nodes, num_links, topo_seed)
    gpath = os.path.join(folder, name + ".graph")
    write_graph_file(gpath, num_nodes, links, caps)
    gf = parse_graph_file(gpath)
    sigma = sigma_for_mean_utilisation(gf)
    for j in range(num_tms):
        tm = uniform_tm(num_nodes, sigma, tm_seed0 + j)
        write_demands_file(os.path.join(folder, "TM", "%s.%d.demands" % (name, j)), tm)
    if write_sp_cache:
        with open(os.path.join(folder, "shortest_paths.json"), "w") as fp:
            json.dump(bfs_lexmin_paths(gf), fp)
    return gf


def max_decisions(num_nodes: int, percentage: float = 0.15) -> int:
    """``ceil(N(N-1)*X)`` exactly as ``environment16.py:199``."""
    return int(math.ceil(num_nodes * (num_nodes - 1) * percentage))
