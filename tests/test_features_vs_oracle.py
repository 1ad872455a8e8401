"""CleanupFeatures / HarvestFeatures (cleanup_features.py:156-284, harvest_features.py:173-336) on the CUDA path vs the C
oracle on seeded random rollouts: every output and the full state, bit by bit, over whole episodes, with masked resets and
the next-step auto-reset, for agent counts that fill an octet (8) and that do not (2, 4, 5), batches that do not fill
their last warp, and short horizons that force many episode boundaries."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def _maps(kind):
    from contracts_b200.maps import CLEANUP_MAP, HARVEST_MAP
    return CLEANUP_MAP if kind == "cleanup" else HARVEST_MAP


def _contract(kind):
    return "CleanupContract" if kind == "cleanup" else "HarvestFeaturemodLocalContract"


def _compare_state(env, orc, ctx):
    st, so = env.get_state(), orc.get_state()
    for k in ("pos", "ori", "cells", "theta", "t"):
        gu.assert_same(k, st[k].cpu().numpy(), so[k], ctx)


def _compare_step(env, o, ctx):
    gu.assert_same("obs", env.obs.cpu().numpy(), o["obs"], ctx)
    gu.assert_same("rew", env.rew.cpu().numpy(), o["rew"], ctx)
    gu.assert_same("base_rew", env.base_rew.cpu().numpy(), o["base_rew"], ctx)
    gu.assert_same("transfers", env.transfers.cpu().numpy(), o["transfers"], ctx)
    gu.assert_same("info", env.info.cpu().numpy()[..., :2].astype(np.int32), o["info"][..., :2], ctx)
    gu.assert_same("done", env.done.cpu().numpy(), o["done"], ctx)


@pytest.mark.parametrize("kind,n,E,steps,horizon,contract", [
    ("cleanup", 8, 203, 260, 1000, True),
    ("cleanup", 5, 67, 200, 1000, True),
    ("cleanup", 2, 64, 200, 1000, False),
    ("harvest", 8, 203, 260, 1000, True),
    ("harvest", 4, 130, 200, 1000, True),
    ("harvest", 2, 33, 150, 1000, False),
])
def test_feature_rollout_matches_oracle(oracle_lib, kind, n, E, steps, horizon, contract):
    import torch
    from contracts_b200.features import BatchedFeatureEnv
    c = _contract(kind) if contract else None
    env = BatchedFeatureEnv(kind, E, n, horizon=horizon, contract=c, seed=4242, first_env_id=1000)
    orc = oracle_lib.FeatOracle(kind, E, n, _maps(kind), horizon=horizon, contract=c, seed=4242, first_env_id=1000)
    gu.assert_same("reset obs", env.reset().cpu().numpy(), orc.reset(), "reset")
    _compare_state(env, orc, "reset")
    rng = np.random.default_rng(7)
    nact = 9 if kind == "cleanup" else 8
    for t in range(steps):
        a = rng.integers(0, nact, size=(E, n))
        if t % 3 == 0:                               # bursts of cleaning / eating so that list removals are frequent
            a[rng.random((E, n)) < 0.4] = 7 if kind == "cleanup" else rng.integers(0, 4)
        o = orc.step(a)
        env.step(torch.as_tensor(a.astype(np.uint8)).cuda())
        ctx = "%s n=%d step %d" % (kind, n, t)
        _compare_step(env, o, ctx)
        if t % 10 == 0 or t == steps - 1:
            _compare_state(env, orc, ctx)
        if t % 50 == 49:                             # masked reset of a third of the envs mid-episode
            m = (rng.random(E) < 0.33).astype(np.uint8)
            want = orc.reset(m)
            got = env.reset(torch.as_tensor(m).cuda()).cpu().numpy()
            sel = m.astype(bool)
            gu.assert_same("masked reset obs", got[sel], want[sel], ctx)
            _compare_state(env, orc, ctx + " after masked reset")
    gu.assert_same("metrics", env.metrics_raw().cpu().numpy(), orc.metrics_raw(), "end")


@pytest.mark.parametrize("kind,n", [("cleanup", 8), ("harvest", 8), ("cleanup", 4)])
def test_feature_auto_reset_matches_oracle(oracle_lib, kind, n):
    """Short horizon + ssd_feat_io.auto_reset: an env that finished is reset by the next step call (its actions ignored,
    zero rewards, done cleared).  One single-env oracle per env: a finished one is reset instead of stepped."""
    import torch
    from contracts_b200.features import BatchedFeatureEnv
    E, horizon = 41, 23
    c = _contract(kind)
    env = BatchedFeatureEnv(kind, E, n, horizon=horizon, contract=c, seed=99, first_env_id=5)
    orcs = [oracle_lib.FeatOracle(kind, 1, n, _maps(kind), horizon=horizon, contract=c, seed=99, first_env_id=5 + i) for i in range(E)]
    env.reset()
    for o in orcs:
        o.reset()
    rng = np.random.default_rng(3)
    nact = 9 if kind == "cleanup" else 8
    done_prev = np.zeros(E, bool)
    for t in range(100):
        a = rng.integers(0, nact, size=(E, n))
        if t == 7:                                   # stagger the episodes
            m = (np.arange(E) % 3 == 0).astype(np.uint8)
            for i in np.nonzero(m)[0]:
                orcs[i].reset()
            env.reset(torch.as_tensor(m).cuda())
            done_prev[m.astype(bool)] = False
        env.step(torch.as_tensor(a.astype(np.uint8)).cuda(), auto_reset=True)
        got = {k: getattr(env, k).cpu().numpy() for k in ("obs", "rew", "base_rew", "transfers", "done")}
        st = {k: v.cpu().numpy() for k, v in env.get_state().items()}
        for i in range(E):
            ctx = "%s auto-reset step %d env %d" % (kind, t, i)
            if done_prev[i]:
                gu.assert_same("restart obs", got["obs"][i], orcs[i].reset()[0], ctx)
                assert not got["rew"][i].any() and not got["base_rew"][i].any() and not got["transfers"][i].any() and got["done"][i] == 0, ctx
            else:
                o = orcs[i].step(a[i][None])
                for k in ("obs", "rew", "base_rew", "transfers", "done"):
                    gu.assert_same(k, got[k][i], o[k][0], ctx)
            so = orcs[i].get_state()
            for k in ("pos", "ori", "cells", "theta", "t"):
                gu.assert_same(k, st[k][i], so[k][0], ctx)
        done_prev = got["done"].astype(bool)


def _fits_i8(w):
    return (w == np.round(w)) & (np.abs(w) <= 127) & ~(np.signbit(w) & (w == 0))


@pytest.mark.parametrize("kind,n,E,theta", [("cleanup", 8, 1500, 0.15), ("harvest", 8, 900, 2.5), ("cleanup", 5, 301, 0.0)])
def test_feature_step_host_async_equals_step(kind, n, E, theta):
    """ssd_feat_step_host_async / ssd_step_host_wait (two slots, compact int8 + sparse float64 result block) delivers bit
    for bit what ssd_feat_step produces: observations (device), dones, and — after ssd_host_result_expand — the float64
    rewards; with the next-step auto-reset on a short horizon."""
    import torch
    from contracts_b200.features import BatchedFeatureEnv
    c = _contract(kind)
    a = BatchedFeatureEnv(kind, E, n, horizon=17, contract=c, seed=5, first_env_id=11)
    b = BatchedFeatureEnv(kind, E, n, horizon=17, contract=c, seed=5, first_env_id=11)
    a.reset(); b.reset()
    a.set_contract_params(theta); b.set_contract_params(theta)
    rng = np.random.RandomState(1)
    nact = 9 if kind == "cleanup" else 8
    T = 60
    acts = [torch.as_tensor(rng.randint(0, nact, size=(E, n)).astype(np.uint8)).pin_memory() for _ in range(T)]
    res = [b.new_host_result(), b.new_host_result()]
    want = []
    for t in range(T):
        obs_a, rew_a, done_a, _ = a.step(acts[t].cuda(), auto_reset=True)
        want.append((rew_a.cpu().numpy().copy(), done_a.cpu().numpy().copy(), obs_a.cpu().numpy().copy() if t % 7 == 0 else None))
        if t % 17 == 16:                                # (auto-reset draws a new theta: pin it again on both sides)
            pass
    tickets, sparse_total = [], 0

    def check(t):
        nonlocal sparse_total
        b.step_host_wait(tickets[t])
        r = res[t & 1]
        assert np.array_equal(r.rewards().view(np.uint64), want[t][0].view(np.uint64)), (kind, t)
        assert np.array_equal(r.done, want[t][1]), (kind, t)
        envs = r.rec_env[:r.count]
        assert len(np.unique(envs)) == r.count
        sparse_env = np.nonzero(~_fits_i8(want[t][0]).all(1))[0]
        assert np.array_equal(np.sort(envs), sparse_env), (kind, t)
        sparse_total += r.count
    for t in range(T):
        tickets.append(b.step_host_async(acts[t], res[t & 1], auto_reset=True))
        if want[t][2] is not None:
            assert np.array_equal(b.obs.cpu().numpy().view(np.uint64), want[t][2].view(np.uint64)), (kind, t)
        if t >= 1:
            check(t - 1)
    check(T - 1)
    assert sparse_total > 0                             # transfers were paid (theta is redrawn at the auto-reset): records exercised
    from contracts_b200 import _lib
    t0 = b.step_host_async(acts[0], res[0]); t1 = b.step_host_async(acts[1], res[1])
    with pytest.raises(_lib.SsdError):
        b.step_host_async(acts[2], res[0])
    b.step_host_wait(t0); b.step_host_wait(t1)


# custom maps: cramped ones (agents block each other, share squares, fire into walls at point-blank range), a cleanup map
# without waste points and one whose point lists need all eight mask words
_CRAMPED_CLEANUP = ["@@@@@@", "@PPPP@", "@PPPP@", "@HBBR@", "@PPPP@", "@@@@@@"]
_CRAMPED_HARVEST = ["@@@@@@", "@PPPP@", "@PAAP@", "@PAAP@", "@PPPP@", "@@@@@@"]
_NO_WASTE_CLEANUP = ["@@@@@@@", "@PPBBP@", "@PBBBP@", "@PPSSP@", "@@@@@@@"]
_BIG_CLEANUP = ["@" * 34] + ["@" + "P" * 4 + "B" * 12 + "R" * 8 + "H" * 8 + "@"] * 15 + ["@" * 34]      # 180 apple, 240 waste points
_BIG_HARVEST = ["@" * 30] + ["@" + "P" * 4 + "A" * 24 + "@"] * 10 + ["@" * 30]                         # 240 apple points


@pytest.mark.parametrize("kind,amap,n,nact", [
    ("cleanup", _CRAMPED_CLEANUP, 8, 9), ("cleanup", _CRAMPED_CLEANUP, 3, 9), ("harvest", _CRAMPED_HARVEST, 8, 8),
    ("harvest", _CRAMPED_HARVEST, 1, 8), ("cleanup", _NO_WASTE_CLEANUP, 4, 9), ("cleanup", _BIG_CLEANUP, 8, 9),
    ("harvest", _BIG_HARVEST, 7, 8)], ids=["cramped_cleanup_n8", "cramped_cleanup_n3", "cramped_harvest_n8", "cramped_harvest_n1",
                                          "cleanup_no_waste_n4", "big_cleanup_n8", "big_harvest_n7"])
def test_feature_rollout_custom_maps(oracle_lib, kind, amap, n, nact):
    import torch
    from contracts_b200.features import BatchedFeatureEnv
    E, steps = 50, 150
    c = _contract(kind) if n > 1 else None
    env = BatchedFeatureEnv(kind, E, n, ascii_map=amap, horizon=60, contract=c, seed=11, first_env_id=77)
    orc = oracle_lib.FeatOracle(kind, E, n, amap, horizon=60, contract=c, seed=11, first_env_id=77)
    gu.assert_same("reset obs", env.reset().cpu().numpy(), orc.reset(), "reset")
    _compare_state(env, orc, "reset")
    rng = np.random.default_rng(n)
    for t in range(steps):
        a = rng.integers(0, nact, size=(E, n))
        o = orc.step(a)
        env.step(torch.as_tensor(a.astype(np.uint8)).cuda())
        ctx = "%s custom map n=%d step %d" % (kind, n, t)
        _compare_step(env, o, ctx)
        _compare_state(env, orc, ctx)
        if o["done"].any():
            m = o["done"].astype(np.uint8)
            want = orc.reset(m)
            got = env.reset(torch.as_tensor(m).cuda()).cpu().numpy()
            gu.assert_same("reset obs", got[m.astype(bool)], want[m.astype(bool)], ctx)
    gu.assert_same("metrics", env.metrics_raw().cpu().numpy(), orc.metrics_raw(), "end")
