"""Prioritised Level Replay index work on the device (csrc/plr.cu; reference environments/level_sampler.py:183-234,
331-408) against the oracle restatement (oracle/level_sampler.py): ids, scores and flags bit-exact, including quirk Q3,
duplicate buffer ids among the agents (last writer wins), tied scores, too few replayable levels and buffer sizes that
are not powers of two; then the product sampler with the device kernels against the same sampler on its numpy path."""
import numpy as np
import pytest
import torch

from oracle import prng
from oracle import level_sampler as O

pytestmark = pytest.mark.gpu


def _dev(a, dt):
    return torch.from_numpy(np.ascontiguousarray(a, dt)).cuda()


def _random_state(rs, B, p_active=0.2, p_new=0.4, ties=False):
    score = rs.randn(B).astype(np.float32)
    if ties:
        score = np.round(score * 2).astype(np.float32) / 2          # many exactly tied scores
    return score, rs.rand(B) < p_active, rs.rand(B) < p_new


@pytest.mark.parametrize("B,m,seed,ties", [(64, 8, 0, False), (4000, 512, 1, False), (2048, 512, 2, True), (100, 100, 3, True)])
def test_reset_lowest_matches_oracle_including_q3(built_lib, B, m, seed, ties):
    from to_ued_b200 import _lib
    rs = np.random.RandomState(seed)
    score, active, new = _random_state(rs, B, ties=ties)
    ids, sc, ac, nw = O.reset_lowest_scoring(score, active, new, m)
    d_s, d_a, d_n = _dev(score, np.float32), _dev(active, np.uint8), _dev(new, np.uint8)
    out = torch.empty(m, dtype=torch.int32, device="cuda")
    _lib.call("toued_plr_reset_lowest", _lib.ptr(d_s), _lib.ptr(d_a), _lib.ptr(d_n), B, m, _lib.ptr(out), _lib.stream_ptr())
    np.testing.assert_array_equal(out.cpu().numpy(), ids)
    np.testing.assert_array_equal(d_s.cpu().numpy(), sc)
    np.testing.assert_array_equal(d_a.cpu().numpy().astype(bool), ac)
    np.testing.assert_array_equal(d_n.cpu().numpy().astype(bool), nw)          # Q3: rebuilt from `active`


CASES = [
    # B, n, seed, p_active, p_new, p_term, ties, temperature
    (64, 8, 0, 0.2, 0.4, 0.5, False, 1.0),
    (4000, 512, 1, 0.15, 0.3, 0.3, False, 1.0),          # the reference's default buffer size (not a power of two)
    (2048, 512, 2, 0.25, 0.3, 1.0, True, 0.3),           # GROOVE bench shape, every agent terminated, tied scores
    (600, 512, 3, 0.5, 0.6, 0.7, False, 1.0),            # too few replayable levels: uniform scores, no replay
    (1024, 256, 4, 0.0, 0.0, 0.4, True, 2.0),            # nothing new: the random ids come from inadmissible levels
    (8192, 4096, 5, 0.1, 0.45, 0.5, False, 1.0),         # maximum size: 2 shuffle rounds
]


@pytest.mark.parametrize("B,n,seed,pa,pn,pt,ties,temp", CASES)
def test_select_matches_oracle(built_lib, B, n, seed, pa, pn, pt, ties, temp):
    from to_ued_b200 import _lib
    rs = np.random.RandomState(seed)
    score, active, new = _random_state(rs, B, pa, pn, ties)
    old_ids = rs.randint(0, B, n).astype(np.int32)
    old_ids[: n // 4] = old_ids[n // 4: 2 * (n // 4)]                # duplicate ids: .at[].set() resolves last-wins
    terminated = rs.rand(n) < pt
    new_scores = (rs.randn(n) * 3).astype(np.float32)
    key = prng.PRNGKey(100 + seed)
    # oracle.plr_select uses temperature 1 inside replay_ids: call its pieces with the case's temperature
    want_ids, sc, ac, nw, rng_after = _oracle_select(key, score, active, new, old_ids, terminated, new_scores, 0.6, temp)
    d_s, d_a, d_n = _dev(score, np.float32), _dev(active, np.uint8), _dev(new, np.uint8)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(float(0xFFFFFFFF))))
    p = _lib.ptr
    d_ids, d_term, d_ns = _dev(old_ids, np.int32), _dev(terminated, np.uint8), _dev(new_scores, np.float32)   # (kept alive)
    _lib.call("toued_plr_select", int(key[0]), int(key[1]), p(d_s), p(d_a), p(d_n), B, p(d_ids), p(d_term), p(d_ns), n, 0.6,
              float(temp), max(1, rounds), p(out), _lib.stream_ptr())
    got = out.cpu().numpy()
    np.testing.assert_array_equal(got, want_ids)
    np.testing.assert_array_equal(d_s.cpu().numpy(), sc)
    np.testing.assert_array_equal(d_a.cpu().numpy().astype(bool), ac)
    np.testing.assert_array_equal(d_n.cpu().numpy().astype(bool), nw)
    assert (got[~terminated] == old_ids[~terminated]).all()


def _oracle_select(rng, score, active, new, old_ids, terminated, new_scores, p_replay, temperature):
    """oracle.level_sampler.plr_select with the temperature exposed (level_sampler.py:183-234)."""
    B, n = len(score), len(old_ids)
    sc, ac, nw = score.copy(), active.copy(), new.copy()
    for i in range(n):
        sc[old_ids[i]] = new_scores[i] if terminated[i] else score[old_ids[i]]
        ac[old_ids[i]] = False if terminated[i] else active[old_ids[i]]
        nw[old_ids[i]] = False if terminated[i] else new[old_ids[i]]
    ks = prng.split(rng, 3); rng, random_rng = ks[0], ks[2]
    rep = O.replay_ids(sc, ac, nw, n, temperature)
    rnd = O.random_ids(random_rng, ac, nw, n)
    ks = prng.split(rng, 2); rng, k = ks[0], ks[1]
    n_to_replay = int((prng.uniform(k, (n,)) < np.float32(p_replay)).sum())
    use = (np.arange(n) < n_to_replay) & (B - int((nw | ac).sum()) >= n)
    ks = prng.split(rng, 2); rng, k = ks[0], ks[1]
    use = use[prng.shuffle(k, n)]
    ids = np.where(terminated, np.where(use, rep, rnd), old_ids).astype(np.int32)
    ac[ids] = True
    return ids, sc, ac, nw, rng


def test_oracle_select_helper_is_the_oracle():
    """The temperature-exposing copy above must stay identical to oracle.level_sampler.plr_select at temperature 1."""
    rs = np.random.RandomState(9)
    score, active, new = _random_state(rs, 256)
    old_ids = rs.randint(0, 256, 32).astype(np.int32)
    term, ns = rs.rand(32) < 0.5, rs.randn(32).astype(np.float32)
    a = _oracle_select(prng.PRNGKey(4), score, active, new, old_ids, term, ns, 0.6, 1.0)
    b = O.plr_select(prng.PRNGKey(4), score, active, new, old_ids, term, ns, 0.6)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("mode,B", [("mazes", 64), ("all_vrandlife", 96)])
def test_sampler_device_plr_equals_host_plr(built_lib, mode, B):
    """LevelSampler._plan_sample on the device PLR path against the same sampler forced onto its numpy path, over several
    calls (no terminations, some, all): buffer state, chosen ids, lifetimes, level records and the returned key chain."""
    from to_ued_b200.experiments.parse_args import parse_args
    from to_ued_b200.environments.level_sampler import LevelSampler
    n = 16
    args = parse_args(["--env_mode", mode, "--num_agents", str(n), "--num_mini_batches", "1", "--score_function", "alg_regret",
                       "--buffer_size", str(B)])
    out = {}
    for dev_plr in (True, False):
        ls = LevelSampler(args)
        ls.device_plr = dev_plr
        buf = ls.initialize_buffer(prng.PRNGKey(3))
        buf = buf.replace(active=np.arange(B) < n)
        level = _index(ls, buf, np.arange(n))
        rs = np.random.RandomState(0)
        trace = []
        for call in range(4):
            host_step = np.where(rs.rand(n) < (0.0, 0.5, 1.0, 0.3)[call], level.lifetime, 0).astype(np.int32)
            scores = rs.randn(n).astype(np.float32)

            def score_fn(keys, only, device_result=False, scores=scores):
                s = np.where(only, scores, 0).astype(np.float32)
                return torch.from_numpy(s).cuda() if device_result else s
            assert ls._plr_on_device(buf, n) == dev_plr
            buf, plan = ls._plan_sample(prng.PRNGKey(50 + call), buf, host_step, level, True, score_fn)
            if plan is not None:
                level = plan[1]
            trace.append((buf.score.copy(), buf.active.copy(), buf.new.copy(), buf.level.lifetime.copy(),
                          buf.level.packed.cpu().numpy().copy(), level.buffer_id.copy(), level.lifetime.copy(),
                          level.packed.cpu().numpy().copy(), None if plan is None else plan[2].copy()))
        out[dev_plr] = trace
    for a, b in zip(out[True], out[False]):
        for x, y in zip(a, b):
            if x is None or y is None:
                assert x is None and y is None
            else:
                np.testing.assert_array_equal(x, y)


def _index(ls, buf, ids):
    from to_ued_b200.environments.level_sampler import _index_level
    return _index_level(buf.level, ids)
