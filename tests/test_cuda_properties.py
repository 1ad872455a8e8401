"""GPU: size-independent properties at the full benchmark size (BASELINE configs: 131072 envs x 8 agents cleanup,
16384 envs x 4 agents harvest), where the CPU oracle is only used on a slice."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PALETTE = {(0, 0, 0), (180, 180, 180), (0, 255, 0), (99, 156, 194), (113, 75, 24), (0, 0, 255), (2, 81, 154),
           (204, 0, 204), (216, 30, 54), (254, 151, 0), (100, 255, 255), (99, 99, 255), (250, 204, 255)}
AGENT_RGB = [(0, 0, 255), (2, 81, 154), (204, 0, 204), (216, 30, 54), (254, 151, 0), (100, 255, 255), (99, 99, 255),
             (250, 204, 255)]


@pytest.mark.parametrize("kind,E,n,nact,contract", [("cleanup", 131072, 8, 9, "CleanupContract"),
                                                    ("harvest", 16384, 4, 8, "HarvestFeaturemodLocalContract")])
def test_full_size_rollout_properties(oracle_lib, kind, E, n, nact, contract):
    import torch
    from contracts_b200.batched import BatchedGridEnv
    from contracts_b200.maps import CLEANUP_MAP, HARVEST_MAP
    amap = CLEANUP_MAP if kind == "cleanup" else HARVEST_MAP
    big = BatchedGridEnv(kind + "_new", E, n, horizon=40, contract=contract, seed=73907)
    k0, ks = E - 77, 64                                   # a slice near the end of the batch
    small = BatchedGridEnv(kind + "_new", ks, n, horizon=40, contract=contract, seed=73907, first_env_id=k0)
    orc = oracle_lib.GridOracle(kind, ks, n, amap, horizon=40, contract=contract, seed=73907, first_env_id=k0)
    ob, osm = big.reset(), small.reset()
    assert torch.equal(ob[k0:k0 + ks], osm)
    assert np.array_equal(osm.cpu().numpy(), orc.reset())
    cleaned_total = torch.zeros((), dtype=torch.float64, device=big.device)
    for t in range(90):                                    # crosses the horizon twice (masked re-resets)
        a = big.random_actions(t, nact)
        a_small = small.random_actions(t, nact)
        assert torch.equal(a[k0:k0 + ks], a_small)        # actions are keyed by the global env id too
        obs, rew, done, info = big.step(a)
        o2, r2, d2, i2 = small.step(a_small)
        # 1. sharding invariance: a slice of the big batch == a small handle with the same global ids == the oracle
        assert torch.equal(obs[k0:k0 + ks], o2) and torch.equal(rew[k0:k0 + ks], r2) and torch.equal(info[k0:k0 + ks], i2)
        o = orc.step(a_small.cpu().numpy(), want_features=False)
        assert np.array_equal(o2.cpu().numpy(), o["obs"]) and np.array_equal(r2.cpu().numpy().view(np.uint64), o["rew"].view(np.uint64))
        # 2. transfers redistribute, they do not create reward
        assert float((rew.sum(1) - big.base_rew.sum(1)).abs().max()) < 1e-9
        if kind == "cleanup":
            cleaned_total += info[:, :, 1].sum()
        if t % 30 == 0:
            # 3. every pixel is a palette colour; the centre pixel of agent i is an agent colour >= its own index
            px = obs[::97].reshape(-1, 3).cpu().numpy()
            assert {tuple(p) for p in np.unique(px, axis=0)} <= PALETTE
            centre = obs[::97, :, 7, 7, :].cpu().numpy()
            for i in range(n):
                allowed = set(AGENT_RGB[i:n])
                assert {tuple(p) for p in np.unique(centre[:, i], axis=0)} <= allowed
        if bool(done.any()):
            assert bool(done.all())                        # equal horizons
            big.reset(done)
            small.reset(d2)
            orc.reset(d2.cpu().numpy())
            cleaned_total.zero_()
    m = big.metrics_raw()
    assert float(m[:, 5].max()) == 0.0                     # no device error flags
    if kind == "cleanup":
        assert float(m[:, 4].sum()) == float(cleaned_total)          # dirt_cleaned == sum of cleaned_squares infos
        st = big.get_state()
        assert int((st["map"] == ord("H")).sum()) > 0
    # 4. state round trip: get_state -> set_state on a fresh handle reproduces the next step
    st = small.get_state()
    clone = BatchedGridEnv(kind + "_new", ks, n, horizon=40, contract=contract, seed=73907, first_env_id=k0)
    clone.reset()
    for _ in range(10):                                    # bring the clone's episode counter / t in line
        pass
    clone.set_state(map=st["map"], pos=st["pos"], ori=st["ori"], t=st["t"], theta=st["theta"])
    st2 = clone.get_state()
    for k in ("map", "pos", "ori", "t", "theta"):
        assert torch.equal(st[k], st2[k]), k


def test_step_host_equals_step():
    """ssd_step_host (host buffers, overlapped D2H copy) produces exactly what ssd_step + explicit copies produce."""
    import torch
    from contracts_b200.batched import BatchedGridEnv
    for kind, contract, nact in (("cleanup_new", "CleanupContract", 8), ("harvest_new", "HarvestFeaturemodLocalContract", 8)):
        E, n = 300, 5
        a = BatchedGridEnv(kind, E, n, contract=contract, seed=12, first_env_id=40, horizon=25)
        b = BatchedGridEnv(kind, E, n, contract=contract, seed=12, first_env_id=40, horizon=25)
        a.reset(); b.reset()
        rng = np.random.RandomState(0)
        rew_h = torch.empty((E, n), dtype=torch.float64).pin_memory()
        done_h = torch.empty((E,), dtype=torch.uint8).pin_memory()
        for t in range(30):
            acts = torch.as_tensor(rng.randint(0, nact, size=(E, n)).astype(np.uint8)).pin_memory()
            obs_a, rew_a, done_a, _ = a.step(acts.cuda())
            obs_b, _, _ = b.step_host(acts, rew_h, done_h)
            assert torch.equal(obs_a.cpu(), obs_b.cpu()), (kind, t)
            assert np.array_equal(rew_a.cpu().numpy().view(np.uint64), rew_h.numpy().view(np.uint64)), (kind, t)
            assert torch.equal(done_a.cpu(), done_h), (kind, t)
            if done_h.any():
                a.reset(done_a); b.reset(done_h.cuda())
