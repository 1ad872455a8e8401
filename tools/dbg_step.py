"""debug: a few steps of a small cleanup batch vs the oracle (run under `timeout`)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from contracts_b200.batched import BatchedGridEnv
from contracts_b200.maps import CLEANUP_MAP, HARVEST_MAP
from oracle import oracle

kind = sys.argv[1] if len(sys.argv) > 1 else "cleanup"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
n = int(sys.argv[4]) if len(sys.argv) > 4 else 8
amap = CLEANUP_MAP if kind == "cleanup" else HARVEST_MAP
contract = "CleanupContract" if kind == "cleanup" else "HarvestFeaturemodLocalContract"
env = BatchedGridEnv(kind + "_new", E, n, contract=contract, seed=73907, device="cuda:0")
orc = oracle.GridOracle(kind, E, n, amap, contract=contract, seed=73907)
o0 = env.reset(); torch.cuda.synchronize(); print("reset done", flush=True)
assert np.array_equal(o0.cpu().numpy(), orc.reset())
rng = np.random.RandomState(0)
na = 9 if kind == "cleanup" else 8
for t in range(steps):
    a = rng.randint(0, na, size=(E, n))
    o = orc.step(a, want_features=False)
    obs, rew, done, info = env.step(torch.as_tensor(a.astype(np.uint8)).cuda())
    torch.cuda.synchronize()
    ok_obs = np.array_equal(obs.cpu().numpy(), o["obs"])
    ok_rew = np.array_equal(rew.cpu().numpy().view(np.uint64), o["rew"].view(np.uint64))
    print("step", t, "obs", ok_obs, "rew", ok_rew, flush=True)
    if not (ok_obs and ok_rew):
        bad = np.nonzero((obs.cpu().numpy() != o["obs"]).reshape(E, -1).any(1))[0]
        print("bad envs (obs)", bad[:10])
        badr = np.nonzero((rew.cpu().numpy() != o["rew"]).any(1))[0]
        print("bad envs (rew)", badr[:10])
        sys.exit(1)
print("ok")
