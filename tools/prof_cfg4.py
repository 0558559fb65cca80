import cProfile, pstats, sys, io, time
sys.path.insert(0, '.')
from enero_b200.runtime import Context
from enero_b200 import workloads as wl
from enero_b200.train import TrainingRun
ctx = Context(0)
ctx.upload_weights(0, wl.shipped_actor_weights())
run = TrainingRun.synthetic(ctx, ((20, 31, 1003), (23, 25, 1004), (17, 31, 1005)))
buf = run.collect()
from enero_b200.ppo import get_advantages
returns, adv = get_advantages(buf["values"], buf["masks"], buf["rewards"], context=ctx)
# warm
run.agent.ppo_update(buf["actions"], buf["actions_probs"], buf["tensors"], buf["critic_features"], returns, adv)
ctx.sync()
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
run.agent.ppo_update(buf["actions"], buf["actions_probs"], buf["tensors"], buf["critic_features"], returns, adv)
ctx.sync()
pr.disable()
print("update seconds", time.perf_counter() - t0)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])
