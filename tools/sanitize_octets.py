"""Small workload over the eight-lanes-per-env kernels (feature envs, selfdrive, small-batch gridworld logic) for
compute-sanitizer runs: a few dozen steps incl. masked resets and the next-step auto-reset, results checked against the oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from contracts_b200.batched import BatchedGridEnv          # noqa: E402
from contracts_b200.features import BatchedFeatureEnv      # noqa: E402
from contracts_b200.maps import CLEANUP_MAP, HARVEST_MAP   # noqa: E402
from contracts_b200.selfdrive import BatchedCarEnv         # noqa: E402
from oracle import oracle as O                             # noqa: E402

O.build()
rng = np.random.default_rng(0)
steps = int(os.environ.get("SAN_STEPS", "25"))
E = 37
for kind, amap, c, nact in (("cleanup", CLEANUP_MAP, "CleanupContract", 9), ("harvest", HARVEST_MAP, "HarvestFeaturemodLocalContract", 8)):
    env = BatchedFeatureEnv(kind, E, 8, horizon=10, contract=c, seed=3, first_env_id=9)
    orc = O.FeatOracle(kind, E, 8, amap, horizon=10, contract=c, seed=3, first_env_id=9)
    assert np.array_equal(env.reset().cpu().numpy(), orc.reset())
    for t in range(steps):
        a = rng.integers(0, nact, size=(E, 8))
        if t % 2 == 0:
            a[rng.random((E, 8)) < 0.4] = 7 if kind == "cleanup" else 1
        o = orc.step(a)
        obs, rew, done, _ = env.step(torch.as_tensor(a.astype(np.uint8)).cuda())
        assert np.array_equal(obs.cpu().numpy(), o["obs"]) and np.array_equal(rew.cpu().numpy(), o["rew"]), (kind, t)
        if o["done"].any():
            m = o["done"].astype(np.uint8)
            want = orc.reset(m)
            got = env.reset(torch.as_tensor(m).cuda()).cpu().numpy()
            assert np.array_equal(got[m.astype(bool)], want[m.astype(bool)])
    print("features", kind, "ok")
car = BatchedCarEnv(E, 8, contract="SelfdriveContractDistprop", seed=3, first_env_id=9)
oc = O.CarOracle(E, 8, contract=True, seed=3, first_env_id=9)
assert np.array_equal(car.reset().cpu().numpy(), oc.reset())
for t in range(4 * steps):
    a = (rng.uniform(-0.7, 1.0, size=(E, 8)) * 0.3).astype(np.float32)
    o = oc.step(a)
    obs, rew, done, _ = car.step(torch.from_numpy(a).cuda())
    assert np.array_equal(obs.cpu().numpy(), o["obs"]) and np.array_equal(rew.cpu().numpy(), o["rew"]), t
print("selfdrive ok")
os.environ["SSD_LOGIC8"] = "1"
for kind, amap, c, nact in (("cleanup", CLEANUP_MAP, "CleanupContract", 9), ("harvest", HARVEST_MAP, "HarvestFeaturemodLocalContract", 8)):
    g = BatchedGridEnv(kind + "_new", E, 8, horizon=1000, contract=c, seed=3, first_env_id=9)
    og = O.GridOracle(kind, E, 8, amap, horizon=1000, contract=c, seed=3, first_env_id=9)
    assert np.array_equal(g.reset().cpu().numpy(), og.reset())
    for t in range(steps):
        a = rng.integers(0, nact, size=(E, 8))
        o = og.step(a, want_features=False)
        obs, rew, done, _ = g.step(torch.as_tensor(a.astype(np.uint8)).cuda())
        assert np.array_equal(obs.cpu().numpy(), o["obs"]) and np.array_equal(rew.cpu().numpy(), o["rew"]), (kind, t)
    print("grid logic8", kind, "ok")
torch.cuda.synchronize()
print("done")
