"""Training loop of train_Enero_3top_script.py (:379-722) on the GPU engine: rollouts on three
topologies with per-topology sample quotas, bootstrap value, GAE, ``ppo_update`` (8 epochs x
minibatches of 55), greedy evaluation on 20 TMs x 3 topologies, the tagged log-line protocol
(``a,`` ``c,`` ``;,`` ``+,`` ``<,`` ``>,`` ``ENTR,`` ``REW,`` ``lr,`` ``MAX REWD: x REWD_ID: n,``),
``ckpt_ACT-n`` / ``ckpt_CRT-n`` checkpoints and the ``(max_reward, lr)`` pickle.

Differences from the reference in how the work is laid out:
  * the 60 greedy evaluation episodes run as three device-resident batches of 20 instead of 60
    sequential Python episodes;
  * with a process group, every rank owns the rollouts of the topologies ``rank::world`` (per-topology
    numpy RNG streams), the sample records are all-gathered (a few KB each), and each PPO minibatch is
    split over the ranks with one gradient all-reduce per Adam step; only rank 0 writes the log
    and the checkpoints.
``run()`` reproduces the launch schedule of train_Enero_3top.py / train_Enero_3top_script.py:428-444 as the
reference EXECUTES it (the logged learning rate compounds on the pickled value, the optimiser keeps the
checkpointed learning rate, the Python / numpy RNGs are re-seeded at every 20-episode launch); the behaviour
the script's author probably intended (the decayed rate actually applied) is behind ``apply_decayed_lr=True``.
"""
from __future__ import annotations

import os
import pickle
import random
import tempfile
import time

import numpy as np

from . import datasets as ds
from . import gym_graph
from .checkpoint import Checkpoint
from .engine import EpisodeBatch
from .ppo import ENTROPY_BETA, PPOActorCritic, get_advantages, hparams

SEED = 9
EVALUATION_EPISODES = 20
DECAY_STEPS = 60
DECAY_RATE = 0.96
ENTROPY_STEP = 60
QUOTA_MULTIPLIERS = (5, 4, 6)       # num_samples_top{1,2,3} = ceil(pct * N(N-1)) * {5,4,6}   (train script :39-41)


def decayed_learning_rate(step, base=0.0002):
    """train script :97-101"""
    lr = base * (DECAY_RATE ** (step / DECAY_STEPS))
    return 10e-5 if lr < 10e-5 else lr


def launch_schedule(first_iteration, n_iterations, lr, episodes_per_launch=20, base_entropy_beta=ENTROPY_BETA):
    """(iteration, learning rate as logged / pickled, entropy weight, launch boundary?) for every PPO episode of
    train_Enero_3top.py's loop, as the reference executes it (train_Enero_3top_script.py:428-444): at a launch
    boundary whose iteration is a multiple of 60 the rate becomes lr * 0.96 ** (i / 60) floored at 1e-4, compounding
    on the value pickled by the previous launch -- Logs/expSP_3top_15_B_NEWLogs.txt shows exactly this sequence
    (0.0002, 0.000192, 0.0001769472, 0.0001565515579392, ...)."""
    for i in range(first_iteration, first_iteration + n_iterations):
        boundary = (i - first_iteration) % episodes_per_launch == 0
        if boundary and i % DECAY_STEPS == 0:
            lr = decayed_learning_rate(i, lr)
        yield i, lr, (base_entropy_beta / 10 if i >= ENTROPY_STEP else base_entropy_beta), boundary


class TrainingRun:
    def __init__(self, ctx, train_envs, eval_envs, quotas, percentage_demands=0.15, learning_rate=None,
                 entropy_beta=ENTROPY_BETA, log_path=None, checkpoint_dir=None, counter=1, max_reward=-1000.0,
                 process_group=None, eval_episodes=EVALUATION_EPISODES, rollout_mode="device"):
        self.ctx = ctx
        self.train_envs, self.eval_envs, self.quotas = list(train_envs), list(eval_envs), list(quotas)
        self.pct = percentage_demands
        self.entropy_beta = entropy_beta
        self.lr = hparams['learning_rate'] if learning_rate is None else learning_rate
        self.agent = PPOActorCritic(context=ctx, buff_size=sum(self.quotas), learning_rate=self.lr,
                                    entropy_beta=entropy_beta, process_group=process_group)
        self.pg = process_group
        self._rank0 = True
        if process_group is not None:
            import torch.distributed as dist
            self._rank0 = dist.get_rank(process_group) == 0
        self.log = open(log_path, "a") if (log_path and self._rank0) else None
        self.checkpoint_dir = checkpoint_dir
        self.counter = counter
        self.max_reward = max_reward
        self.reward_id = 0
        self.eval_episodes = eval_episodes
        self.training_tm_ids = list(range(100))
        self.rollout_mode = rollout_mode
        self._tm_cache = {}
        self._eval_tms = {}
        # rollouts / greedy evaluation of the topologies run side by side: one helper context (stream + scratch) per
        # topology on the same device, driven from one host thread each -- a batch of 4-6 episodes is bound by launch
        # latency, not by the GPU (ENERO_ROLLOUT_THREADS=0: one after the other on the training context)
        self._helpers = None
        self._pool = None
        self._use_helpers = os.environ.get("ENERO_ROLLOUT_THREADS", "1") != "0"
        self.checkpoint_actor = Checkpoint(model=self.agent.actor, optimizer=self.agent.optimizer)
        self.checkpoint_critic = Checkpoint(model=self.agent.critic, optimizer=self.agent.optimizer)
        random.seed(SEED)
        np.random.seed(SEED)

    # ---- construction helpers ------------------------------------------------------------------
    @classmethod
    def synthetic(cls, ctx, specs, n_train_tms=100, n_eval_tms=EVALUATION_EPISODES, folder=None, **kw):
        """three synthetic Zoo-shaped datasets (TRAIN / EVALUATE splits like convert_dataset.py:50-70)"""
        folder = folder or tempfile.mkdtemp(prefix="enero_train_")
        train_envs, eval_envs, quotas = [], [], []
        for i, (n, l, seed) in enumerate(specs):
            name = "syn%d" % i
            for split, count, first in (("TRAIN", n_train_tms, 2000), ("EVALUATE", n_eval_tms, 5000)):
                ds.make_synthetic_dataset(os.path.join(folder, name, split), name, n, l, seed, count, first)
            for split, lst in (("TRAIN", train_envs), ("EVALUATE", eval_envs)):
                env = gym_graph.make('GraphEnv-v16', context=ctx)
                env.seed(SEED)
                env.generate_environment(os.path.join(folder, name, split), name, 100, 100, kw.get("percentage_demands", 0.15))
                env.top_K_critical_demands = True
                lst.append(env)
            quotas.append(int(np.ceil(kw.get("percentage_demands", 0.15) * n * (n - 1))) * QUOTA_MULTIPLIERS[i % 3])
        run = cls(ctx, train_envs, eval_envs, quotas, **kw)
        run.training_tm_ids = list(range(n_train_tms))
        return run

    def restore(self, counter):
        """train script :452-457"""
        d = self.checkpoint_dir
        self.checkpoint_actor.restore(os.path.join(d, "ckpt_ACT-%d" % counter))
        self.checkpoint_critic.restore(os.path.join(d, "ckpt_CRT-%d" % counter))
        self.agent.load_optimizer_state()

    # ---- rollouts (train script :483-625) ---------------------------------------------------------
    def _collect_topology(self, env, quota, rng_choice, sample_tm):
        buf = {k: [] for k in ("tensors", "critic_features", "actions", "values", "masks", "rewards", "actions_probs")}
        agent = self.agent
        reached = False
        tm_id = sample_tm()
        while not reached:
            demand, source, destination = env.reset(tm_id)
            while 1:
                action_dist, tensor = agent.pred_action_distrib_sp(env, source, destination)
                features = agent.critic_get_graph_features(env)
                q_value = agent.critic_value(features)
                action = rng_choice(len(action_dist), action_dist)
                onehot = np.zeros(len(action_dist), dtype=np.float32)
                onehot[action] = 1.0
                reward, done, _, new_demand, new_source, new_destination, _, _, _ = env.step(action, demand, source, destination)
                buf["tensors"].append(tensor)
                buf["critic_features"].append(features)
                buf["actions"].append(onehot)
                buf["values"].append(q_value)
                buf["masks"].append(not done)
                buf["rewards"].append(reward)
                buf["actions_probs"].append(action_dist)
                demand, source, destination = new_demand, new_source, new_destination
                if len(buf["tensors"]) == quota:
                    reached = True
                    break
                if done:
                    break
        return buf

    # ---- rollouts, device resident and batched ------------------------------------------------------
    def _tm(self, env, tm_id):
        key = (id(env), int(tm_id))
        if key not in self._tm_cache:
            self._tm_cache[key] = ds.parse_demands_file(os.path.join(env.dataset_folder_name, "TM", "%s.%d.demands"
                                                                     % (env.graph_topology_name, tm_id)), env.numNodes)
        return self._tm_cache[key]

    def _side_contexts(self, n):
        """n contexts holding the agent's current weights (None when running on the training context alone)"""
        self.agent._bind()
        if not self._use_helpers or n < 2:
            return None
        from .runtime import Context
        from concurrent.futures import ThreadPoolExecutor
        if self._helpers is None:
            self._helpers = []
        while len(self._helpers) < n:
            h = Context(self.ctx.device)
            h.math, h.engine = self.ctx.math, self.ctx.engine
            self._helpers.append(h)
        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=8)
        self.ctx.sync()
        for which in (0, 1):
            w = self.ctx.download_weights(which)
            for h in self._helpers[:n]:
                h.upload_weights(which, w)
        return self._helpers[:n]

    def _map_topologies(self, fn, items):
        """fn(context, item) for every item, concurrently on the helper contexts when there are any"""
        helpers = self._side_contexts(len(items))
        if helpers is None:
            return [fn(self.ctx, it) for it in items]
        return list(self._pool.map(lambda hi: fn(hi[0], hi[1]), zip(helpers, items)))

    def close(self):
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None
        for h in self._helpers or []:
            h.close()
        self._helpers = None

    def collect_device(self):
        """The rollouts of train script :483-625 as device-resident batches: every topology runs `episodes per quota`
        episodes (quota / episode length = 5, 4, 6 with the reference's multipliers) at once through
        enero_batch_rollout -- actor, critic, action sampling (Philox stream per episode) and Env16.step without a host
        round trip -- and the records are cut to the topology's quota in (episode, step) order, so the buffer has the
        reference's shape: topology 1, 2, 3, quota-truncated tails with mask = True, GAE running across the
        boundaries (quirk Q5).  With a process group the (topology, episode) slots are dealt round-robin to the
        ranks -- every GPU rolls out -- and the records are all-gathered.  Traffic matrices and action streams are drawn
        from generators keyed by (SEED, PPO episode, topology, slot): the result does not depend on the number of ranks."""
        world, rank = 1, 0
        if self.pg is not None:
            import torch.distributed as dist
            world, rank = dist.get_world_size(self.pg), dist.get_rank(self.pg)
        self.agent._bind()
        per_topo = []
        for i, (env, quota) in enumerate(zip(self.train_envs, self.quotas)):
            d = env.topology.max_decisions(self.pct)
            per_topo.append(max(1, -(-quota // max(1, d))))
        rounds = 0
        recs = {}                                    # (topology, slot) -> record dict of one episode
        need = [True] * len(self.train_envs)
        while any(need):
            slots = [(i, rounds * per_topo[i] + j) for i in range(len(self.train_envs)) if need[i] for j in range(per_topo[i])]
            mine = slots[rank::world]
            out = {}

            def roll(ctx, i):
                env = self.train_envs[i]
                js = [j for (ii, j) in mine if ii == i]
                tm_ids = [int(np.random.RandomState((SEED, self.counter, i, j)).choice(self.training_tm_ids)) for j in js]
                eb = EpisodeBatch(ctx, env.topology, len(js), self.pct)
                try:
                    eb.reset(np.stack([self._tm(env, t) for t in tm_ids]))
                    r = eb.rollout(seed=(SEED << 40) ^ (self.counter << 20) ^ (i << 12) ^ js[0])
                finally:
                    eb.close()
                res = {}
                for b, j in enumerate(js):
                    n = int(r["steps"][b])
                    res[(i, j)] = {k: np.array(r[k][b, :n]) for k in ("loads", "sd", "bw", "action", "prob", "value", "reward")}
                    res[(i, j)]["last_value"] = float(r["last_value"][b])
                return res

            for res in self._map_topologies(roll, sorted({s[0] for s in mine})):
                out.update(res)
            if world > 1:
                import torch.distributed as dist
                gathered = [None] * world
                dist.all_gather_object(gathered, out, group=self.pg)
                for g in gathered:
                    recs.update(g)
            else:
                recs.update(out)
            rounds += 1
            for i, quota in enumerate(self.quotas):
                have = sum(len(v["bw"]) for (ii, _), v in recs.items() if ii == i)
                need[i] = have < quota
        buf = {k: [] for k in ("tensors", "critic_features", "actions", "values", "masks", "rewards", "actions_probs")}
        boot = 0.0
        for i, (env, quota) in enumerate(zip(self.train_envs, self.quotas)):
            topo, taken = env.topology, 0
            for j in sorted(jj for (ii, jj) in recs if ii == i):
                e = recs[(i, j)]
                n = len(e["bw"])
                for s in range(n):
                    if taken == quota:
                        break
                    k = len(topo.middlepoints(int(e["sd"][s, 0]), int(e["sd"][s, 1])))
                    buf["tensors"].append({"topology": topo, "loads": e["loads"][s], "sd": (int(e["sd"][s, 0]), int(e["sd"][s, 1])),
                                           "bw": float(e["bw"][s]), "num_candidates": k, "old_prob": np.float32(e["prob"][s])})
                    buf["critic_features"].append({"topology": topo, "loads": e["loads"][s]})
                    buf["actions"].append(int(e["action"][s]))
                    buf["actions_probs"].append(None)
                    buf["values"].append(np.float32(e["value"][s]))
                    buf["masks"].append(s != n - 1)
                    buf["rewards"].append(float(e["reward"][s]))
                    taken += 1
                    # value of the state this environment is left in after the sample (train script :622-625)
                    boot = float(e["value"][s + 1]) if s + 1 < n else e["last_value"]
                if taken == quota:
                    break
        buf["values"] = buf["values"] + [np.float32(boot)]
        return buf

    def collect(self):
        """-> the experience buffer of one PPO episode (lists, topology 1 then 2 then 3) + bootstrap value.
        rollout_mode "device" (default): batched device-resident rollouts (collect_device); "host": one environment
        stepped from Python with the reference's global RNGs (random.sample for the TM, np.random.choice for the action),
        which replays the reference's sampling sequence decision by decision."""
        if self.rollout_mode == "device":
            return self.collect_device()
        world, rank = 1, 0
        if self.pg is not None:
            import torch.distributed as dist
            world, rank = dist.get_world_size(self.pg), dist.get_rank(self.pg)
        parts = [None] * len(self.train_envs)
        for i, (env, quota) in enumerate(zip(self.train_envs, self.quotas)):
            if world == 1:
                # the reference's global RNGs: random.sample for the TM, np.random.choice for the action
                choice = lambda n, p: int(np.random.choice(n, p=p))
                parts[i] = self._collect_topology(env, quota, choice, lambda: random.sample(self.training_tm_ids, 1)[0])
            elif i % world == rank:
                rs = np.random.RandomState((SEED, self.counter, i))
                choice = lambda n, p, rs=rs: int(rs.choice(n, p=p))
                parts[i] = self._collect_topology(env, quota, choice, lambda rs=rs: int(rs.choice(self.training_tm_ids)))
        if world > 1:
            import torch.distributed as dist
            mine = {i: _portable(p) for i, p in enumerate(parts) if p is not None}
            gathered = [None] * world
            dist.all_gather_object(gathered, mine, group=self.pg)
            for g in gathered:
                for i, p in g.items():
                    parts[i] = _attach(p, self.train_envs[i].topology)
        buf = {k: [] for k in parts[0]}
        for p in parts:
            for k in buf:
                buf[k] += p[k]
        # bootstrap value of the state the last environment was left in (train script :622-625)
        last = self.train_envs[-1]
        if world == 1 or (len(self.train_envs) - 1) % world == rank:
            boot = float(self.agent.critic_value(self.agent.critic_get_graph_features(last)))
        else:
            boot = 0.0
        if world > 1:
            import torch.distributed as dist
            lst = [boot]
            dist.broadcast_object_list(lst, src=(len(self.train_envs) - 1) % world, group=self.pg)
            boot = lst[0]
        buf["values"] = buf["values"] + [np.float32(boot)]
        return buf

    # ---- greedy evaluation (train script :633-693), batched --------------------------------------------
    def evaluate(self):
        def greedy(ctx, env):
            key = id(env)
            if key not in self._eval_tms:             # the evaluation matrices are fixed: parsed once
                self._eval_tms[key] = np.stack([ds.parse_demands_file(os.path.join(env.dataset_folder_name, "TM", "%s.%d.demands"
                                                                                     % (env.graph_topology_name, t)), env.numNodes)
                                                for t in range(self.eval_episodes)])
            tms = self._eval_tms[key]
            eb = EpisodeBatch(ctx, env.topology, tms.shape[0], self.pct)
            try:
                eb.reset(tms)
                eb.run_greedy()
                st = eb.state()
                _, r, _ = eb.history()
            finally:
                eb.close()
            rewards = []
            for i in range(tms.shape[0]):
                rew = 0
                for v in r[i, :st["steps"][i]]:         # rewardAddTest += reward, in step order
                    rew += v
                rewards.append(rew)
            return rewards, list(st["max_uti"]), [float(np.std(st["loads"][i])) for i in range(tms.shape[0])]

        rewards, max_uti, uti_std = [], [], []
        for a, b, c in self._map_topologies(greedy, list(self.eval_envs)):     # (with whatever weights the agent holds)
            rewards += a
            max_uti += b
            uti_std += c
        return np.array(rewards), np.array(max_uti), np.array(uti_std)

    # ---- one PPO episode ---------------------------------------------------------------------------
    def _log(self, line):
        if self.log:
            self.log.write(line)

    def ppo_episode(self):
        t0 = time.perf_counter()
        buf = self.collect()
        t1 = time.perf_counter()
        returns, advantages = get_advantages(buf["values"], buf["masks"], buf["rewards"], context=self.ctx)
        shuffle = np.random.shuffle
        if self.pg is not None:                     # every rank must draw the same permutations
            rs = np.random.RandomState((SEED, self.counter, 99))
            shuffle = rs.shuffle
        actor_loss, critic_loss = self.agent.ppo_update(buf["actions"], buf["actions_probs"], buf["tensors"],
                                                        buf["critic_features"], returns, advantages, shuffle=shuffle)
        self.ctx.sync()
        t2 = time.perf_counter()
        self._log("a,%s,\n" % actor_loss)
        self._log("c,%s,\n" % critic_loss)
        rewards_test, max_link_uti, uti_std = self.evaluate()
        t3 = time.perf_counter()
        eval_mean = float(np.mean(rewards_test))
        self._log(";,%s,\n" % np.mean(uti_std))
        self._log("+,%s,\n" % 0.0)
        self._log("<,%s,\n" % np.amax(max_link_uti))
        self._log(">,%s,\n" % 0.0)
        self._log("ENTR,%s,\n" % self.entropy_beta)
        self._log("REW,%s,\n" % eval_mean)
        self._log("lr,%s,\n" % self.lr)
        if eval_mean > self.max_reward:
            self.max_reward = eval_mean
            self.reward_id = self.counter
            self._log("MAX REWD: %s REWD_ID: %s,\n" % (self.max_reward, self.reward_id))
        if self.log:
            self.log.flush()
        if self.checkpoint_dir:
            self.agent.store_optimizer_state()
            if self._rank0:                       # one writer; files appear under their final names atomically
                os.makedirs(self.checkpoint_dir, exist_ok=True)
                tmp_dir = tempfile.mkdtemp(prefix=".ckpt_tmp_", dir=self.checkpoint_dir)
                # a restored checkpoint carries save_counter = counter - 1, so save() writes ckpt_*-<counter>
                self.checkpoint_actor.save_counter = self.counter - 1
                self.checkpoint_critic.save_counter = self.counter - 1
                self.checkpoint_actor.save(os.path.join(tmp_dir, "ckpt_ACT"))
                self.checkpoint_critic.save(os.path.join(tmp_dir, "ckpt_CRT"))
                names = sorted(os.listdir(tmp_dir), key=lambda f: f == "checkpoint")     # the `checkpoint` index file last
                for f in names:
                    os.replace(os.path.join(tmp_dir, f), os.path.join(self.checkpoint_dir, f))
                os.rmdir(tmp_dir)
            if self.pg is not None:
                import torch.distributed as dist
                dist.barrier(group=self.pg)
        self.counter += 1
        n = len(buf["tensors"])
        steps = 8 * ((n + 54) // 55)
        return {"samples": n, "adam_steps": steps, "collect_s": t1 - t0, "update_s": t2 - t1, "eval_s": t3 - t2,
                "actor_loss": actor_loss, "critic_loss": critic_loss, "eval_mean_reward": eval_mean,
                "eval_max_uti": float(np.amax(max_link_uti)), "update_samples_per_s": 8 * n / (t2 - t1)}

    def run(self, first_iteration, n_iterations, episodes_per_launch=20, base_entropy_beta=ENTROPY_BETA,
            apply_decayed_lr=False):
        """The outer schedule of train_Enero_3top.py:28-34 + train_Enero_3top_script.py:428-444, one "launch" every
        `episodes_per_launch` PPO episodes, exactly as the reference executes it:
          * ``hparams['learning_rate']`` starts from the value the previous launch pickled (``self.lr``) and, when the
            iteration is a multiple of 60, becomes ``lr * 0.96 ** (i / 60)`` floored at 1e-4 -- it COMPOUNDS on the
            pickled value (:97-101, :433-440); this is the number the ``lr,`` log lines and the pickle carry;
          * the optimiser does not see it: ``checkpoint.restore`` (:452-457) overwrites the freshly built Adam's
            learning rate with the checkpointed one, so training continues at the rate of the first launch
            (SURVEY quirk Q6).  ``apply_decayed_lr=True`` applies the decayed rate instead;
          * the entropy weight is a tenth of its base value from iteration 60 on (:442-443);
          * every launch is a fresh process that seeds ``random`` and ``numpy.random`` with 9 (:34-37), which restarts
            the traffic-matrix and action sampling sequences."""
        stats = []
        for i, lr, beta, boundary in launch_schedule(first_iteration, n_iterations, self.lr, episodes_per_launch, base_entropy_beta):
            if boundary:
                self.lr = lr
                if apply_decayed_lr:
                    self.agent.optimizer.learning_rate = np.float32(lr)
                self.entropy_beta = beta
                self.agent.entropy_beta = float(beta)
                random.seed(SEED)
                np.random.seed(SEED)
            stats.append(self.ppo_episode())
        return stats

    def save_state(self, path):
        """train script :720-722"""
        with open(path, "wb") as f:
            pickle.dump((self.max_reward, self.lr), f)


def _portable(buf):
    """sample records without the (non-picklable) topology handles"""
    out = dict(buf)
    out["tensors"] = [{k: v for k, v in t.items() if k != "topology"} for t in buf["tensors"]]
    out["critic_features"] = [{k: v for k, v in t.items() if k != "topology"} for t in buf["critic_features"]]
    return out


def _attach(buf, topo):
    for t in buf["tensors"]:
        t["topology"] = topo
    for t in buf["critic_features"]:
        t["topology"] = topo
    return buf
