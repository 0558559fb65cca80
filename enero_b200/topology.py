"""Topology handle: host precompute in C++ (csrc/topology.cpp) + numpy views.

Mirrors what ``Env16.generate_environment`` leaves on the environment object
(environment16.py:526-580): ``numNodes``, ``numEdges``, ``first``/``second``,
``firstTrueSize``, ``link_capacity_feature``, ``shortest_paths``,
``src_dst_k_middlepoints`` and ``edgesDict``.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import _lib
from . import datasets as ds


class Topology:
    def __init__(self, graph: ds.GraphFile, K: int = 100, paths: dict | None = None):
        lib = _lib.load()
        self.graph_file = graph
        self.edges = graph.edge_list()                       # edge-id order (Q2)
        src = np.ascontiguousarray([e[0] for e in self.edges], dtype=np.int32)
        dst = np.ascontiguousarray([e[1] for e in self.edges], dtype=np.int32)
        cap = np.ascontiguousarray([graph.links_bw[e] for e in self.edges], dtype=np.float64)
        N = graph.num_nodes
        sp_off = sp_nodes = None
        if paths is not None:                                # shortest_paths.json content
            off, nodes = [0], []
            for s in range(N):
                for d in range(N):
                    if s != d:
                        nodes.extend(int(v) for v in paths["%d:%d" % (s, d)])
                    off.append(len(nodes))
            sp_off = np.ascontiguousarray(off, dtype=np.int32)
            sp_nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        h = C.c_void_p()
        _lib.check(lib.enero_topo_build(N, len(self.edges), _lib.ptr(src), _lib.ptr(dst), _lib.ptr(cap), int(K),
                                        _lib.ptr(sp_off), _lib.ptr(sp_nodes), C.byref(h)))
        self._h = h
        sz = np.zeros(9, dtype=np.int64)
        _lib.check(lib.enero_topo_sizes(h, _lib.ptr(sz)))
        (self.N, self.E, self.P, self.K, self.off_first, self.off_second, n_mids, self.Kmax, n_sp) = (int(v) for v in sz)
        self.cap = cap
        self.edge_src, self.edge_dst = src, dst
        self.first = np.zeros(self.P, dtype=np.int32)
        self.second = np.zeros(self.P, dtype=np.int32)
        _lib.check(lib.enero_topo_get_pairs(h, _lib.ptr(self.first), _lib.ptr(self.second)))
        self.sp_off = np.zeros(N * N + 1, dtype=np.int32)
        self.sp_edge = np.zeros(max(1, n_sp), dtype=np.int32)
        _lib.check(lib.enero_topo_get_sp_edges(h, _lib.ptr(self.sp_off), _lib.ptr(self.sp_edge)))
        self.mid_off = np.zeros(N * N + 1, dtype=np.int32)
        self.mids = np.zeros(max(1, n_mids), dtype=np.int32)
        _lib.check(lib.enero_topo_get_middlepoints(h, _lib.ptr(self.mid_off), _lib.ptr(self.mids)))
        self.link_capacity_feature = np.zeros(self.E, dtype=np.float32)
        _lib.check(lib.enero_topo_get_capacity_feature(h, _lib.ptr(self.link_capacity_feature)))
        self.edgesDict = {"%d:%d" % (i, j): k for k, (i, j) in enumerate(self.edges)}

    # ---- construction helpers ---------------------------------------------------------
    @classmethod
    def from_dataset(cls, dataset_folder: str, topology_name: str, K: int = 100) -> "Topology":
        """Same inputs as generate_environment (environment16.py:537-543,483-500)."""
        gf = ds.parse_graph_file(os.path.join(dataset_folder, topology_name + ".graph"))
        sp = os.path.join(dataset_folder, "shortest_paths.json")
        paths = json.load(open(sp)) if os.path.isfile(sp) else None
        return cls(gf, K=K, paths=paths)

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().enero_topo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference-shaped views --------------------------------------------------------
    def middlepoints(self, s: int, d: int) -> list:
        i = s * self.N + d
        return [int(v) for v in self.mids[self.mid_off[i]:self.mid_off[i + 1]]]

    def middlepoint_dict(self) -> dict:
        """``src_dst_k_middlepoints`` keyed ``"s:d"`` (environment16.py:403-408)."""
        return {"%d:%d" % (s, d): self.middlepoints(s, d)
                for s in range(self.N) for d in range(self.N) if s != d}

    def sp_edges(self, s: int, d: int) -> np.ndarray:
        i = s * self.N + d
        return self.sp_edge[self.sp_off[i]:self.sp_off[i + 1]]

    def path_nodes(self, s: int, d: int) -> list:
        e = self.sp_edges(s, d)
        return [int(self.edge_src[e[0]])] + [int(self.edge_dst[k]) for k in e] if len(e) else [s]

    def paths_dict(self) -> dict:
        return {"%d:%d" % (s, d): self.path_nodes(s, d)
                for s in range(self.N) for d in range(self.N) if s != d}

    def k_shortest_paths(self, K: int = 5) -> dict:
        """``Env20.allPaths`` keyed ``"s:d"`` (environment20.py:165-196): the K simple paths with the fewest hops among
        those of at most 2*diameter links, ties in lexicographic node order; built by csrc/ksp.cpp."""
        lib = _lib.load()
        _lib.check(lib.enero_topo_build_ksp(self._h, int(K)))
        sz = np.zeros(2, dtype=np.int64)
        _lib.check(lib.enero_topo_ksp_sizes(self._h, _lib.ptr(sz)))
        pair_off = np.zeros(self.N * self.N + 1, dtype=np.int32)
        path_off = np.zeros(int(sz[0]) + 1, dtype=np.int32)
        nodes = np.zeros(max(1, int(sz[1])), dtype=np.int32)
        _lib.check(lib.enero_topo_get_ksp(self._h, _lib.ptr(pair_off), _lib.ptr(path_off), _lib.ptr(nodes)))
        out = {}
        for s in range(self.N):
            for d in range(self.N):
                if s != d:
                    i = s * self.N + d
                    out["%d:%d" % (s, d)] = [[int(v) for v in nodes[path_off[p]:path_off[p + 1]]]
                                             for p in range(pair_off[i], pair_off[i + 1])]
        return out

    def max_decisions(self, percentage: float = 0.15) -> int:
        return ds.max_decisions(self.N, percentage)
