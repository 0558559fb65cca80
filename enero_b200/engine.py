"""Device-resident batch of ENERO episodes (GraphEnv-v16 semantics for B traffic
matrices of one topology) and the two-stage optimiser built on it:

    DRL stage   play_middRout_games_sp            script_eval_on_single_topology.py:162-188
    LS stage    play_DRL_GNN_sp_hill_climbing_games                        ... :473-509

Everything numeric runs in libenero_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .runtime import Context
from .topology import Topology


class EpisodeBatch:
    def __init__(self, ctx: Context, topo: Topology, batch: int, percentage_demands: float = 0.15):
        self.ctx, self.topo, self.B = ctx, topo, int(batch)
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.enero_batch_create(ctx.handle, topo.handle, self.B, float(percentage_demands), C.byref(h)))
        self._h = h
        ctx._adopt(self)
        self.Dmax = int(self._lib.enero_batch_dmax(h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.enero_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Env16 -----------------------------------------------------------------------------
    def reset(self, tms):
        """tms: fp64 [B,N,N] (floored demands). Returns (bw[B], sd[B,2]) of the first demands."""
        tms = np.ascontiguousarray(tms, dtype=np.float64)
        if tms.shape != (self.B, self.topo.N, self.topo.N):
            raise ValueError("tms must be [B,N,N]")
        _lib.check(self._lib.enero_batch_reset(self._h, _lib.ptr(tms)))
        st = self.state()
        return st["bw"], st["cur"]

    def policy(self, want_probs: bool = True):
        probs = np.zeros((self.B, self.topo.Kmax), dtype=np.float32) if want_probs else None
        nK = np.zeros(self.B, dtype=np.int32) if want_probs else None
        _lib.check(self._lib.enero_batch_policy(self._h, _lib.ptr(probs), _lib.ptr(nK)))
        return probs, nK

    def value(self) -> np.ndarray:
        v = np.zeros(self.B, dtype=np.float32)
        _lib.check(self._lib.enero_batch_value(self._h, _lib.ptr(v)))
        return v

    def step(self, actions=None, want_outputs: bool = True):
        """actions None = greedy argmax of the last policy(). -> (reward[B], done[B]);
        want_outputs=False with actions=None only enqueues the step (asynchronous)."""
        a = None if actions is None else np.ascontiguousarray(actions, dtype=np.int32).reshape(-1)
        if a is not None and a.size != self.B:
            raise ValueError("need one action per episode")
        if a is None and not want_outputs:
            _lib.check(self._lib.enero_batch_step(self._h, None, None, None))
            return None, None
        reward = np.zeros(self.B, dtype=np.float64)
        done = np.zeros(self.B, dtype=np.uint8)
        _lib.check(self._lib.enero_batch_step(self._h, _lib.ptr(a), _lib.ptr(reward), _lib.ptr(done)))
        return reward, done.astype(bool)

    def step_state(self, actions):
        """Env16.step with host actions and the post-step state in one stream round trip
        -> (reward[B], done[B], state dict with loads, cur, bw, max_uti, max_edge)."""
        a = np.ascontiguousarray(actions, dtype=np.int32).reshape(-1)
        if a.size != self.B:
            raise ValueError("need one action per episode")
        B, E = self.B, self.topo.E
        reward = np.zeros(B)
        done = np.zeros(B, dtype=np.uint8)
        st = {"loads": np.zeros((B, E)), "cur": np.zeros((B, 2), dtype=np.int32), "bw": np.zeros(B),
              "max_uti": np.zeros(B), "max_edge": np.zeros(B, dtype=np.int32)}
        _lib.check(self._lib.enero_batch_step_state(self._h, _lib.ptr(a), _lib.ptr(reward), _lib.ptr(done), _lib.ptr(st["loads"]),
                                                    _lib.ptr(st["cur"]), _lib.ptr(st["bw"]), _lib.ptr(st["max_uti"]),
                                                    _lib.ptr(st["max_edge"])))
        return reward, done.astype(bool), st

    def run_greedy(self):
        _lib.check(self._lib.enero_batch_run_greedy(self._h))

    def snapshot(self):
        _lib.check(self._lib.enero_batch_snapshot(self._h))

    def restore(self):
        _lib.check(self._lib.enero_batch_restore(self._h))

    def state(self) -> dict:
        B, E = self.B, self.topo.E
        out = {"loads": np.zeros((B, E)), "cur": np.zeros((B, 2), dtype=np.int32), "bw": np.zeros(B),
               "max_uti": np.zeros(B), "max_edge": np.zeros(B, dtype=np.int32),
               "done": np.zeros(B, dtype=np.uint8), "steps": np.zeros(B, dtype=np.int32)}
        _lib.check(self._lib.enero_batch_get_state(self._h, *[_lib.ptr(out[k]) for k in
                                                   ("loads", "cur", "bw", "max_uti", "max_edge", "done", "steps")]))
        return out

    def eligible(self):
        elig = np.zeros((self.B, self.Dmax, 2), dtype=np.int32)
        n = np.zeros(self.B, dtype=np.int32)
        _lib.check(self._lib.enero_batch_get_eligible(self._h, _lib.ptr(elig), _lib.ptr(n)))
        return elig, n

    def history(self):
        a = np.zeros((self.B, self.Dmax), dtype=np.int32)
        r = np.zeros((self.B, self.Dmax))
        m = np.zeros((self.B, self.Dmax))
        _lib.check(self._lib.enero_batch_get_history(self._h, _lib.ptr(a), _lib.ptr(r), _lib.ptr(m)))
        return a, r, m

    def best_arrays(self):
        """-> (best max-uti[B], ospf[B], routing int32[B,Dmax,3] = (s,d,mid) in insertion order, n[B])."""
        best = np.zeros(self.B)
        ospf = np.zeros(self.B)
        routing = np.zeros((self.B, self.Dmax, 3), dtype=np.int32)
        n = np.zeros(self.B, dtype=np.int32)
        _lib.check(self._lib.enero_batch_get_best(self._h, _lib.ptr(best), _lib.ptr(ospf), _lib.ptr(routing), _lib.ptr(n)))
        return best, ospf, routing, n

    def best(self):
        """-> (best max-uti[B], ospf[B], routing list per episode [(s,d,mid), ...])."""
        best, ospf, routing, n = self.best_arrays()
        return best, ospf, [[tuple(v) for v in routing[b, :n[b]].tolist()] for b in range(self.B)]

    # ---- Env15 + HILL_CLIMBING ----------------------------------------------------------------
    def local_search(self, routing=None, mode: int = 0, max_moves: int = 4096, want_moves: bool = True):
        """routing: per-episode list of (s,d,mid) in insertion order, or None for the batch's own
        best DRL routing.  -> dict(start, final, n_moves, moves, move_val)."""
        B = self.B
        r_arr = n_arr = None
        if routing is not None:
            r_arr = np.zeros((B, self.Dmax, 3), dtype=np.int32)
            n_arr = np.zeros(B, dtype=np.int32)
            for b, lst in enumerate(routing):
                if len(lst) > self.Dmax:
                    raise ValueError("routing longer than the demand budget")
                n_arr[b] = len(lst)
                for k, e in enumerate(lst):
                    r_arr[b, k] = e
        start = np.zeros(B)
        final = np.zeros(B)
        n_moves = np.zeros(B, dtype=np.int32)
        moves = np.zeros((B, max_moves, 3), dtype=np.int32) if want_moves else None
        vals = np.zeros((B, max_moves)) if want_moves else None
        _lib.check(self._lib.enero_batch_local_search(self._h, int(mode), _lib.ptr(r_arr), _lib.ptr(n_arr), int(max_moves),
                                                      _lib.ptr(start), _lib.ptr(final), _lib.ptr(n_moves),
                                                      _lib.ptr(moves), _lib.ptr(vals)))
        return {"start": start, "final": final, "n_moves": n_moves, "moves": moves, "move_val": vals}

    def rollout(self, seed: int):
        """Sampled-action episodes (train_Enero_3top_script.py:483-625) for the whole batch, device resident
        -> dict of per-(episode, step) records: loads[B,Dmax,E], sd, bw, action, prob, value, reward, plus
        steps[B] and last_value[B]."""
        B, D, E = self.B, self.Dmax, self.topo.E
        out = {"loads": np.zeros((B, D, E)), "sd": np.zeros((B, D, 2), np.int32), "bw": np.zeros((B, D)),
               "action": np.zeros((B, D), np.int32), "prob": np.zeros((B, D), np.float32), "value": np.zeros((B, D), np.float32),
               "reward": np.zeros((B, D)), "steps": np.zeros(B, np.int32), "last_value": np.zeros(B, np.float32)}
        _lib.check(self._lib.enero_batch_rollout(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, *[_lib.ptr(out[k]) for k in
                   ("loads", "sd", "bw", "action", "prob", "value", "reward", "steps", "last_value")]))
        return out

    # ---- Env20 + SAPAgent ----------------------------------------------------------------------
    def sap(self, tms=None, K: int = 5, seed: int = 9, want_actions: bool = False):
        """play_sap_games (script_eval_on_single_topology.py:511-557) for every traffic matrix of the batch
        (tms fp64 [B,N,N], or None for the matrices of the last reset) -> final max link utilisation [B]
        (and, with want_actions, what SAPAgent.act returned per demand, int32 [B, N(N-1)], -1 = no path had room)."""
        if tms is not None:
            tms = np.ascontiguousarray(tms, dtype=np.float64)
            if tms.shape != (self.B, self.topo.N, self.topo.N):
                raise ValueError("tms must be [B,N,N]")
        out = np.zeros(self.B)
        n = self.topo.N
        acts = np.zeros((self.B, n * (n - 1)), dtype=np.int32) if want_actions else None
        _lib.check(self._lib.enero_batch_sap(self._h, _lib.ptr(tms), int(K), int(seed), _lib.ptr(out), _lib.ptr(acts)))
        return (out, acts) if want_actions else out

    def ls_route_add(self, episode, src, dst, value):
        """allocate_to_destination_sp (value > 0) / decrease_links_utilization_sp (value < 0) on the Local-Search
        loads of one episode -> (loads fp64[E], max utilisation, its edge id or -1)."""
        loads = np.zeros(self.topo.E)
        mx = np.zeros(1)
        me = np.zeros(1, dtype=np.int32)
        _lib.check(self._lib.enero_batch_ls_route_add(self._h, int(episode), int(src), int(dst), float(value),
                                                      _lib.ptr(loads), _lib.ptr(mx), _lib.ptr(me)))
        return loads, float(mx[0]), int(me[0])

    def ls_loads(self):
        """link loads of the Local-Search state, fp64[B,E]"""
        out = np.zeros((self.B, self.topo.E))
        _lib.check(self._lib.enero_batch_get_ls_state(self._h, _lib.ptr(out), None))
        return out

    def ls_max_edge(self):
        out = np.zeros(self.B, dtype=np.int32)
        _lib.check(self._lib.enero_batch_get_ls_state(self._h, None, _lib.ptr(out)))
        return out

    def ls_eligible(self):
        cap = max(self.Dmax, self.topo.N * (self.topo.N - 1))
        elig = np.zeros((self.B, cap, 2), dtype=np.int32)
        n = np.zeros(self.B, dtype=np.int32)
        _lib.check(self._lib.enero_batch_get_ls_eligible(self._h, _lib.ptr(elig), _lib.ptr(n), cap))
        return elig, n


def run_enero(ctx: Context, topo: Topology, tms, percentage_demands: float = 0.15, local_search: bool = True):
    """Full Enero inference (greedy DRL + Local Search) for a batch of traffic matrices."""
    tms = np.ascontiguousarray(tms, dtype=np.float64)
    eb = EpisodeBatch(ctx, topo, tms.shape[0], percentage_demands)
    try:
        eb.reset(tms)
        eb.run_greedy()
        best, ospf, routing, n_routing = eb.best_arrays()
        out = {"ospf": ospf, "drl": best, "routing": routing, "n_routing": n_routing, "decisions": eb.state()["steps"]}
        if local_search:
            ls = eb.local_search(None, mode=0, want_moves=False)
            out["enero"] = ls["final"]
            out["ls_moves"] = ls["n_moves"]
        return out
    finally:
        eb.close()
