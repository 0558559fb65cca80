"""PPO agent of train_Enero_3top_script.py (``PPOActorCritic``, :103-365; ``get_advantages``,
:367-377) on the GPU engine.

Same method names, arguments and hyper-parameters as the reference; what changes is the sample
record.  The reference stores, per decision, the materialised TF tensors of the K' candidate
graphs (``link_state [K'E,20]``, ``first``, ``second`` ...).  Here a sample is the compact state the
kernels rebuild those tensors from on device: ``(topology, link loads[E], s, d, bandwidth)`` --
``E*8 + 24`` bytes instead of ~``K'*E*88`` bytes.

The gradient of one minibatch is accumulated per topology (``enero_ppo_accumulate``); with a
process group (one rank per GPU) the minibatch is sharded over ranks and the flat gradient buffer
is summed by ONE all-reduce before every rank applies the identical clip + Adam step.
"""
from __future__ import annotations

import ctypes as C
import gc

import numpy as np

from . import _lib
from .models import ActorModel, CriticModel
from .optim import Adam
from .parallel import shard_indices

# train_Enero_3top_script.py:28-87
MINI_BATCH_SIZE = 55
PPO_EPOCHS = 8
hparams = {'l2': 0.0001, 'link_state_dim': 20, 'readout_units': 20, 'learning_rate': 0.0002, 'T': 5}   # :81-87
gamma = 0.99
lmbda = 0.95
clipping_val = 0.1
max_grad_norm = 0.5
ENTROPY_BETA = 0.01


def get_advantages(values, masks, rewards, context=None):
    """train_Enero_3top_script.py:367-377 -> (returns list, normalised advantages fp64[n]) via k_gae."""
    from .runtime import default_context
    ctx = context if context is not None else default_context()
    n = len(rewards)
    v = np.ascontiguousarray(np.asarray(values, dtype=np.float32).reshape(-1))
    if v.size != n + 1 or len(masks) != n:
        raise ValueError("need len(values) == len(rewards) + 1 == len(masks) + 1")
    m = np.ascontiguousarray(np.asarray(masks, dtype=np.uint8))
    r = np.ascontiguousarray(np.asarray(rewards, dtype=np.float64))
    ret = np.zeros(n)
    adv = np.zeros(n)
    _lib.check(_lib.load().enero_gae(ctx.handle, n, _lib.ptr(v), _lib.ptr(m), _lib.ptr(r), gamma, lmbda,
                                     _lib.ptr(ret), _lib.ptr(adv)))
    return list(ret), adv


class PPOActorCritic:
    def __init__(self, context=None, buff_size=None, learning_rate=None, entropy_beta=ENTROPY_BETA,
                 process_group=None, actor=None, critic=None):
        self.BUFF_SIZE = buff_size
        self.entropy_beta = float(entropy_beta)
        self.memory = []
        self.inds = None if buff_size is None else np.arange(buff_size)
        self.global_step = 0
        self.listQValues = None
        self.softMaxQValues = None
        self.optimizer = Adam(learning_rate=hparams['learning_rate'] if learning_rate is None else learning_rate,
                              beta_1=0.9, epsilon=1e-05)
        self.actor = actor if actor is not None else ActorModel(hparams, context=context)
        self.actor.build()
        self.critic = critic if critic is not None else CriticModel(hparams, context=self.actor.context)
        self.critic.build()
        self.context = self.actor.context
        self._lib = _lib.load()
        self._pg = process_group
        self._grad_view = None
        self._value_cache = None
        self.last_grad_norm = None

    def _bind(self, optimizer=False):
        """this agent's weights (and, for an update, its Adam slots) are what the context's device buffers hold"""
        self.actor.bind()
        self.critic.bind()
        if optimizer:
            owners = self.context.__dict__.setdefault("_optimizer_owner", {})
            cur = owners.get("adam")
            if cur is not self:
                if cur is not None:
                    cur.store_optimizer_state()
                self.load_optimizer_state()
                owners["adam"] = self

    # ---- acting ----------------------------------------------------------------------------
    def pred_action_distrib_sp(self, env, source, destination):
        """train script :125-171 -> (probabilities over the K' middlepoints, sample record)."""
        t = env.topology
        loads = np.array(env.edge_state[:, 0], dtype=np.float64)
        bw = float(env.TM[source][destination])
        # the rollout asks for the critic's value of the same state right after (train script :495-500): both
        # networks run in one stream round trip and the value is handed out by critic_value()
        self._bind()
        probs, nK, values = self.context.decide_value_batch(t, loads[None], np.array([[source, destination]], np.int32),
                                                            np.array([bw]))
        k = int(nK[0])
        self._value_cache = (t, loads, np.float32(values[0]), self.critic._version)
        self.listQValues = None
        self.softMaxQValues = probs[:1, :k]
        tensor = {'topology': t, 'loads': loads, 'sd': (int(source), int(destination)), 'bw': bw, 'num_candidates': k}
        return probs[0, :k].copy(), tensor

    def critic_get_graph_features(self, env):
        """train script :204-230: the critic sees [load/cap, cap/maxCap] of the same loads."""
        return {'topology': env.topology, 'loads': np.array(env.edge_state[:, 0], dtype=np.float64)}

    def critic_value(self, features):
        """self.critic(link_state_critic, first_critic, second_critic, num_edges_critic)[0].numpy()[0]"""
        cached = getattr(self, "_value_cache", None)
        if (cached is not None and cached[0] is features['topology'] and cached[3] == self.critic._version
                and np.array_equal(cached[1], features['loads'])):
            return cached[2]
        self.critic.bind()
        return self.context.value_batch(features['topology'], features['loads'][None])[0]

    # ---- learning ----------------------------------------------------------------------------
    def _group_by_topology(self, inds):
        groups = {}
        for i in inds:
            s = self.memory[int(i)]
            groups.setdefault(id(s['topology']), (s['topology'], []))[1].append(s)
        return list(groups.values())

    def _accumulate(self, samples, inv_m):
        for topo, group in samples:
            n = len(group)
            loads = np.ascontiguousarray(np.stack([s['loads'] for s in group]), dtype=np.float64)
            sd = np.ascontiguousarray([s['sd'] for s in group], dtype=np.int32)
            bw = np.ascontiguousarray([s['bw'] for s in group], dtype=np.float64)
            act = np.ascontiguousarray([s['action'] for s in group], dtype=np.int32)
            oldp = np.ascontiguousarray([s['old_prob'] for s in group], dtype=np.float32)
            adv = np.ascontiguousarray([s['advantage'] for s in group], dtype=np.float32)
            ret = np.ascontiguousarray([s['return'] for s in group], dtype=np.float32)
            _lib.check(self._lib.enero_ppo_accumulate(self.context.handle, topo.handle, n, _lib.ptr(loads), _lib.ptr(sd),
                                                      _lib.ptr(bw), _lib.ptr(act), _lib.ptr(oldp), _lib.ptr(adv), _lib.ptr(ret),
                                                      float(inv_m), self.entropy_beta, clipping_val, _lib.HOST))

    def _allreduce(self):
        """sum the flat gradient buffer over the ranks of the process group (the path's only collective).  The
        collective is enqueued from the library's own stream (torch sees it as an ExternalStream): c10d orders NCCL
        after the kernels already on that stream and the following clip + Adam kernel after NCCL, all on the device --
        no host synchronisation on either side."""
        import torch
        import torch.distributed as dist
        if self._grad_view is None:
            self._grad_view = _device_view(self._lib.enero_ppo_grad_buffer(self.context.handle), _lib.GRAD_FLOATS,
                                           self.context.device)
            self._ext_stream = torch.cuda.ExternalStream(self.context.stream, device=torch.device("cuda", self.context.device))
        _lib.check(self._lib.enero_ppo_grad_ready(self.context.handle))      # lanes joined, gradient buffer complete
        if dist.get_backend(self._pg) == "nccl":
            with torch.cuda.stream(self._ext_stream):
                dist.all_reduce(self._grad_view, op=dist.ReduceOp.SUM, group=self._pg)
        else:                                     # gloo (CPU tests of the N > 1 logic): staged through the host
            self.context.sync()
            dist.all_reduce(self._grad_view, op=dist.ReduceOp.SUM, group=self._pg)
            torch.cuda.current_stream(self.context.device).synchronize()

    def gradients(self, inds):
        """accumulate the minibatch gradient without applying it -> (actor[4201], critic[4201], sums[3])"""
        self._bind()
        _lib.check(self._lib.enero_ppo_zero_grad(self.context.handle))
        self._accumulate(self._group_by_topology(inds), 1.0 / len(inds))
        ga, gc_, sums = np.zeros(_lib.NPARAMS, np.float32), np.zeros(_lib.NPARAMS, np.float32), np.zeros(3, np.float32)
        _lib.check(self._lib.enero_ppo_grad_download(self.context.handle, _lib.ptr(ga), _lib.ptr(gc_), _lib.ptr(sums)))
        return ga, gc_, sums

    # ---- device-resident buffer of one ppo_update -------------------------------------------------------
    def _store_memory(self):
        """upload self.memory once, grouped by topology -> per-sample (topology, index inside its store)"""
        groups = {}
        where = []
        for smp in self.memory:
            t = smp['topology']
            g = groups.setdefault(id(t), (t, []))
            where.append((id(t), len(g[1])))
            g[1].append(smp)
        for t, group in groups.values():
            loads = np.ascontiguousarray(np.stack([s['loads'] for s in group]), dtype=np.float64)
            sd = np.ascontiguousarray([s['sd'] for s in group], dtype=np.int32)
            bw = np.ascontiguousarray([s['bw'] for s in group], dtype=np.float64)
            act = np.ascontiguousarray([s['action'] for s in group], dtype=np.int32)
            oldp = np.ascontiguousarray([s['old_prob'] for s in group], dtype=np.float32)
            adv = np.ascontiguousarray([s['advantage'] for s in group], dtype=np.float32)
            ret = np.ascontiguousarray([s['return'] for s in group], dtype=np.float32)
            _lib.check(self._lib.enero_ppo_store(self.context.handle, t.handle, len(group), _lib.ptr(loads), _lib.ptr(sd),
                                                 _lib.ptr(bw), _lib.ptr(act), _lib.ptr(oldp), _lib.ptr(adv), _lib.ptr(ret)))
        self._store = ({k: g[0] for k, g in groups.items()}, where)

    def _accumulate_stored(self, inds, inv_m):
        topos, where = self._store
        per = {}
        for i in inds:
            k, j = where[int(i)]
            per.setdefault(k, []).append(j)
        for k, idx in per.items():
            idx = np.ascontiguousarray(idx, dtype=np.int32)
            _lib.check(self._lib.enero_ppo_accumulate_stored(self.context.handle, topos[k].handle, idx.size, _lib.ptr(idx),
                                                             float(inv_m), self.entropy_beta, clipping_val))

    def _train_step_combined(self, inds, hist_slot=None):
        """train script :293-324 -> (actor_loss, critic_loss, final_entropy) of the minibatch.  With hist_slot the
        step is only enqueued (its losses and gradient norm go to the device history, ppo_update reads them once)."""
        m = len(inds)
        mine = inds
        if self._pg is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(self._pg), dist.get_world_size(self._pg)
            mine = shard_indices(inds, rank, world)
        self._bind(optimizer=True)
        _lib.check(self._lib.enero_ppo_zero_grad(self.context.handle))
        if len(mine):
            if getattr(self, "_store", None) is not None:
                self._accumulate_stored(mine, 1.0 / m)
            else:
                self._accumulate(self._group_by_topology(mine), 1.0 / m)
        if self._pg is not None:
            self._allreduce()
        self.optimizer.iterations += 1
        self.actor.mark_device_dirty()
        self.critic.mark_device_dirty()
        self._value_cache = None
        args = (self.context.handle, float(self.optimizer.learning_rate), float(self.optimizer.beta_1),
                float(self.optimizer.beta_2), float(self.optimizer.epsilon), max_grad_norm, int(self.optimizer.iterations))
        if hist_slot is not None:
            _lib.check(self._lib.enero_ppo_apply(*args, None, int(hist_slot)))
            return None
        _lib.check(self._lib.enero_ppo_apply(*args, None, 0))
        return self._step_record(0, m)

    def _step_record(self, slot, m):
        rec = np.zeros(4, np.float32)
        _lib.check(self._lib.enero_ppo_history(self.context.handle, int(slot), 1, _lib.ptr(rec)))
        self.last_grad_norm = float(rec[0])
        final_entropy = rec[2] / m
        actor_loss = rec[1] / m - self.entropy_beta * final_entropy
        critic_loss = rec[3] / m
        return float(actor_loss), float(critic_loss), float(final_entropy)

    def ppo_update(self, actions, actions_probs, tensors, critic_features, returns, advantages, shuffle=np.random.shuffle):
        """train script :326-365.  ``actions`` are the one-hot vectors the reference stores.  The buffer goes to the
        device once; the 8 x ceil(n / 55) Adam steps are enqueued without a host synchronisation and the losses of the
        last minibatch (what the reference returns and logs) are read back at the end."""
        n = len(tensors)
        if self.inds is None or len(self.inds) != n:
            self.inds = np.arange(n)
        for pos in range(n):
            a = np.asarray(actions[pos])
            act = int(np.argmax(a)) if a.ndim else int(a)
            sample = dict(tensors[pos])
            # ratio's denominator: reduce_sum(old_act * old_policy_probs)  (:285); device rollouts record it directly
            if actions_probs[pos] is not None:
                sample["old_prob"] = np.float32(np.asarray(actions_probs[pos], dtype=np.float32)[act])
            sample.update(action=act, advantage=np.float32(advantages[pos]), **{'return': np.float32(returns[pos])})
            self.memory.append(sample)
        self._bind(optimizer=True)
        self._store_memory()
        slot, last_m = 0, 1
        try:
            for _ in range(PPO_EPOCHS):
                shuffle(self.inds)
                for start in range(0, n, MINI_BATCH_SIZE):
                    mb = self.inds[start:start + MINI_BATCH_SIZE]
                    slot = (slot + 1) % _lib.PPO_HISTORY
                    last_m = len(mb)
                    self._train_step_combined(mb, hist_slot=slot)
        finally:
            self._store = None
        actor_loss, critic_loss, _ = self._step_record(slot, last_m)
        self.memory.clear()
        self.actor.sync_from_device()
        self.critic.sync_from_device()
        gc.collect()
        return actor_loss, critic_loss

    # ---- optimiser slots <-> device (checkpoint save / restore) --------------------------------
    def upload_slots(self, which, m, v):
        from .runtime import flatten_weights
        fm, fv = flatten_weights(m), flatten_weights(v)
        _lib.check(self._lib.enero_adam_slots_upload(self.context.handle, which, _lib.ptr(fm), _lib.ptr(fv)))

    def load_optimizer_state(self):
        """after Checkpoint.restore: push the restored Adam slots of both models to the device"""
        for which, model in ((_lib.ACTOR, self.actor), (_lib.CRITIC, self.critic)):
            m, v = self.optimizer.get_slots(model)
            self.upload_slots(which, m, v)
        self.context.__dict__.setdefault("_optimizer_owner", {})["adam"] = self

    def store_optimizer_state(self):
        """before Checkpoint.save: pull weights and Adam slots of both models back to the host objects"""
        if self.context.__dict__.get("_optimizer_owner", {}).get("adam") is not self:
            return                              # the device holds another agent's slots; ours are already on the host
        for which, model in ((_lib.ACTOR, self.actor), (_lib.CRITIC, self.critic)):
            model.sync_from_device()
            m, v = self.download_slots(which)
            self.optimizer.set_slots(model, m, v)

    def download_slots(self, which):
        from .runtime import unflatten_weights
        fm, fv = np.zeros(_lib.NPARAMS, np.float32), np.zeros(_lib.NPARAMS, np.float32)
        _lib.check(self._lib.enero_adam_slots_download(self.context.handle, which, _lib.ptr(fm), _lib.ptr(fv)))
        return unflatten_weights(fm), unflatten_weights(fv)



def _device_view(ptr, n_floats, device):
    """torch fp32 view of library-owned device memory (tensor plumbing for torch.distributed)."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device="cuda:%d" % device)
