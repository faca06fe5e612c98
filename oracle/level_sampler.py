"""PLR buffer logic oracle (numpy).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates the index work of environments/level_sampler.py one statement at a time on plain arrays:
``_reset_lowest_scoring`` (:331-353, Q3 reproduced), ``_replay_from_buffer`` (:355-387),
``_sample_random_from_buffer`` (:389-408) and the buffer update / replay-vs-random selection of ``sample``
(:183-234).  Levels are represented by their buffer ids only (the level payload follows the ids)."""
from __future__ import annotations

import numpy as np

from . import prng


def reset_lowest_scoring(score, active, new, minimum_new):
    """-> (reset_ids, score, active, new)"""
    level_scores = np.where(new, -np.inf, score).astype(np.float32)
    level_scores = np.where(active, np.inf, level_scores)
    reset_ids = np.argsort(level_scores, kind="stable")[:minimum_new]
    score2, active2 = score.copy(), active.copy()
    score2[reset_ids] = 0.0
    active2[reset_ids] = False
    new2 = active.copy()                      # Q3: `new=level_buffer.active.at[reset_ids].set(True)`
    new2[reset_ids] = True
    return reset_ids, score2, active2, new2


def replay_scores(score, invalid, temperature=1.0):
    """exp(score / T) masked and normalised (level_sampler.py:365-370).  Float contract shared with the device kernel
    (csrc/plr.cu) and the host path, since these floats can create or break ties of the ranking: ``exp_portable`` on the
    argument clamped to [-80, 80] (the reference's jnp.exp overflows beyond that anyway), a LEFT-TO-RIGHT fp32 sum,
    IEEE division."""
    from .rollout import exp_portable
    F32 = np.float32
    x = np.clip((np.asarray(score, F32) / F32(temperature)).astype(F32), F32(-80), F32(80))
    s = np.where(invalid, F32(0), exp_portable(x)).astype(F32)
    total = np.cumsum(s, dtype=F32)[-1]                      # sequential accumulation
    with np.errstate(invalid="ignore", divide="ignore"):
        return (s / total).astype(F32)


def replay_ids(score, active, new, batch_size, temperature=1.0):
    """rank transform: flip(argsort(p))[:batch]"""
    B = len(score)
    invalid = new | active
    s = replay_scores(score, invalid, temperature)
    p = np.where(B - invalid.sum() < batch_size, np.ones_like(s), s)
    return np.flip(np.argsort(p, kind="stable"))[:batch_size]


def random_ids(key, active, new, batch_size):
    mask = new & ~active
    return prng.choice_no_replace_p(key, mask.astype(np.float32), batch_size)


def plr_select(rng, score, active, new, old_ids, terminated, new_scores, p_replay):
    """level_sampler.py:183-234 given the regret scores of the terminated agents.
    Returns (new_ids, score, active, new) where new_ids are the buffer ids chosen for every agent
    (old ids kept where not terminated)."""
    B, n = len(score), len(old_ids)
    sc, ac, nw = score.copy(), active.copy(), new.copy()
    for i in range(n):                        # .at[old_ids].set(...) — duplicates resolve last-wins
        sc[old_ids[i]] = new_scores[i] if terminated[i] else score[old_ids[i]]
        ac[old_ids[i]] = False if terminated[i] else active[old_ids[i]]
        nw[old_ids[i]] = False if terminated[i] else new[old_ids[i]]
    ks = prng.split(rng, 3); rng, replay_rng, random_rng = ks[0], ks[1], ks[2]
    rep = replay_ids(sc, ac, nw, n)
    rnd = random_ids(random_rng, ac, nw, n)
    ks = prng.split(rng, 2); rng, k = ks[0], ks[1]
    n_to_replay = int((prng.uniform(k, (n,)) < np.float32(p_replay)).sum())
    use = np.arange(n) < n_to_replay
    n_replayable = B - int((nw | ac).sum())
    use = use & (n_replayable >= n)
    ks = prng.split(rng, 2); rng, k = ks[0], ks[1]
    use = use[prng.shuffle(k, n)]
    ids = np.where(use, rep, rnd)
    ids = np.where(terminated, ids, old_ids)
    ac[ids] = True
    return ids, sc, ac, nw, rng
