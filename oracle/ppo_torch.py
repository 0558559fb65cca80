"""TEST INFRASTRUCTURE ONLY -- torch-CPU fp32 restatement of ENERO's PPO update
(train_Enero_3top_script.py:264-324) used as the gradient / optimiser oracle.

PARITY UNPINNED against TensorFlow (not installable here, no golden vectors shipped): the
functions restate the published Keras semantics op for op and rely on torch autograd for the
derivatives.  Only tests/ and bench.py's CPU legs import this module.
"""
from __future__ import annotations

import numpy as np
import torch

from . import enero_oracle as orc

SELU_SCALE = 1.0507009873554804934193349852946
SELU_SCALE_ALPHA = 1.7580993408473768599402175208123


def selu(x):
    return torch.where(x < 0, SELU_SCALE_ALPHA * (torch.exp(torch.clamp(x, max=0.0)) - 1.0), SELU_SCALE * x)


def to_params(weights, dtype=torch.float32):
    """dict name -> numpy  ==>  dict name -> leaf tensors requiring grad.  dtype=torch.float64 gives the
    fp64 ARBITER of the same formulas on the same fp32 inputs: when the CUDA kernels and this fp32
    restatement disagree in the last bits, the distance to the fp64 result says which is closer."""
    return {k: torch.tensor(np.asarray(weights[k], dtype=np.float32), dtype=dtype, requires_grad=True) for k in orc.WEIGHT_NAMES}


def message_passing(w, link_state, first, second, num_edges, T=5):
    """actorPPOmiddR.py:45-63 / criticPPO.py:41-58"""
    dt = w["msg_kernel"].dtype
    h = link_state.to(dt)
    for _ in range(T):
        cat = torch.cat([h[first], h[second]], dim=1)
        m = selu(cat @ w["msg_kernel"] + w["msg_bias"])
        agg = torch.zeros((num_edges, 20), dtype=dt).index_add(0, second, m)
        x = agg @ w["gru_kernel"] + w["gru_bias"][0]
        y = h @ w["gru_recurrent_kernel"] + w["gru_bias"][1]
        z = torch.sigmoid(x[:, :20] + y[:, :20])
        r = torch.sigmoid(x[:, 20:40] + y[:, 20:40])
        hh = torch.tanh(x[:, 40:] + r * y[:, 40:])
        h = z * h + (1 - z) * hh
    return h


def readout(w, g):
    a = selu(g @ w["r1_kernel"] + w["r1_bias"])
    a = selu(a @ w["r2_kernel"] + w["r2_bias"])
    return a @ w["r3_kernel"] + w["r3_bias"]


def actor_logits(w, batch):
    ls = torch.tensor(batch["link_state"])
    first = torch.tensor(batch["first"].astype(np.int64))
    second = torch.tensor(batch["second"].astype(np.int64))
    gid = torch.tensor(batch["graph_id"].astype(np.int64))
    h = message_passing(w, ls, first, second, batch["num_edges"])
    G = int(batch["graph_id"].max()) + 1
    g = torch.zeros((G, 20), dtype=h.dtype).index_add(0, gid, h)
    return readout(w, g).reshape(-1)


def critic_value(w, link_state, first, second, num_edges):
    h = message_passing(w, torch.tensor(link_state), torch.tensor(first.astype(np.int64)),
                        torch.tensor(second.astype(np.int64)), num_edges)
    return readout(w, h.sum(dim=0, keepdim=True)).reshape(())


def minibatch_loss(wa, wc, samples, entropy_beta, clipping_val=0.1):
    """_train_step_combined (train script :293-315).  samples: list of dicts with the actor batch
    ('link_state','graph_id','first','second','num_edges'), critic inputs ('link_state_critic',
    'first_critic','second_critic','num_edges_critic'), 'old_act' (one-hot), 'old_policy_probs',
    'advantage', 'return' (fp32)."""
    actor_losses, entropies, critic_losses = [], [], []
    for s in samples:
        logits = actor_logits(wa, s)
        p = torch.softmax(logits, dim=0)
        dt = logits.dtype
        old_act = torch.tensor(np.asarray(s["old_act"], dtype=np.float32), dtype=dt)
        old_p = torch.tensor(np.asarray(s["old_policy_probs"], dtype=np.float32), dtype=dt)
        adv = torch.tensor(np.float32(s["advantage"]), dtype=dt)
        newp = torch.sum(old_act * p)
        ratio = torch.exp(torch.log(newp) - torch.log(torch.sum(old_act * old_p)))
        surr1 = -ratio * adv
        surr2 = -torch.clamp(ratio, 1 - clipping_val, 1 + clipping_val) * adv
        actor_losses.append(torch.maximum(surr1, surr2))
        entropies.append(-torch.sum(torch.log(p) * p))
        v = critic_value(wc, s["link_state_critic"], s["first_critic"], s["second_critic"], s["num_edges_critic"])
        critic_losses.append((torch.tensor(np.float32(s["return"]), dtype=v.dtype) - v) ** 2)
    critic_loss = torch.stack(critic_losses).mean()
    final_entropy = torch.stack(entropies).mean()
    actor_loss = torch.stack(actor_losses).mean() - entropy_beta * final_entropy
    return actor_loss + critic_loss, actor_loss, critic_loss, final_entropy


def flat(tensors, dtype=np.float32):
    return np.concatenate([np.asarray(t, dtype=dtype).reshape(-1) for t in tensors])


def gradients(wa, wc, samples, entropy_beta):
    """-> (flat actor grad[4201], flat critic grad[4201], actor_loss, critic_loss, entropy)"""
    for w in (wa, wc):
        for t in w.values():
            t.grad = None
    total, al, cl, ent = minibatch_loss(wa, wc, samples, entropy_beta)
    total.backward()
    nd = np.float64 if wa["msg_kernel"].dtype == torch.float64 else np.float32
    ga = flat([wa[k].grad.numpy() if wa[k].grad is not None else np.zeros(tuple(wa[k].shape)) for k in orc.WEIGHT_NAMES], nd)
    gc = flat([wc[k].grad.numpy() if wc[k].grad is not None else np.zeros(tuple(wc[k].shape)) for k in orc.WEIGHT_NAMES], nd)
    return ga, gc, float(al.detach()), float(cl.detach()), float(ent.detach())


def clip_by_global_norm(grads, clip_norm, dtype=np.float32):
    """tf.clip_by_global_norm over the concatenation of all gradient tensors (train script :319)"""
    g = np.concatenate(grads).astype(dtype)
    norm = np.sqrt(np.sum(g ** 2, dtype=dtype))
    scale = dtype(clip_norm) / max(norm, dtype(clip_norm))
    return [dtype(scale) * x.astype(dtype) for x in grads], norm


def adam_step(theta, m, v, g, lr, step, beta1=0.9, beta2=0.999, eps=1e-5, dtype=np.float32):
    """Keras OptimizerV2 Adam dense update as TF's ApplyAdam CPU kernel evaluates it (fp32):
    alpha = lr * sqrt(1 - b2^t) / (1 - b1^t);  m += (g - m) * (1 - b1);  v += (g*g - v) * (1 - b2);
    theta -= (m * alpha) / (sqrt(v) + eps);  t = optimizer.iterations + 1.  dtype=np.float64: the arbiter."""
    f = dtype
    t = f(step)
    alpha = f(lr) * np.sqrt(f(1) - np.power(f(beta2), t)) / (f(1) - np.power(f(beta1), t))
    m = (m + (g - m) * (f(1) - f(beta1))).astype(dtype)
    v = (v + (g * g - v) * (f(1) - f(beta2))).astype(dtype)
    theta = (theta - (m * f(alpha)) / (np.sqrt(v) + f(eps))).astype(dtype)
    return theta, m, v
