"""enero_b200 -- B200-native (sm_100a) implementation of ENERO's hot path behind the reference's own
class / file contracts.  All numerics run in ``libenero_b200.so`` (hand-written CUDA, C ABI in
``include/enero_b200.h``); there is no CPU fallback.

    runtime.Context            device context + stateless batched operators (actor / critic / fused decisions)
    topology.Topology          host precompute of a WAN topology (edge ids, message pairs, shortest paths, middlepoints)
    engine.EpisodeBatch        device-resident batch of GraphEnv-v16 episodes + Local Search (GraphEnv-v15)
    models / actorPPOmiddR / criticPPO     drop-in ``myModel`` classes
    gym_graph                  drop-in ``gym.make('GraphEnv-v15' / '-v16')``
    agents, ppo, train         the eval agent / episode drivers, ``PPOActorCritic``, the training loop
    evaluate, workloads        batched eval worker (+ result records, driver CLI), benchmark workloads
    checkpoint, datasets       TensorBundle reader / writer, ``.graph`` / ``.demands`` ingest + synthetic generators
    parallel                   one-process-per-GPU sharding helpers
"""
__version__ = "0.1.0"
