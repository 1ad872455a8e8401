"""Generate the golden vectors in tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE (oracle/), container-only: needs /root/reference.  Run as
    python -m oracle.make_golden
Each fixture is the reference's own output (oracle/ref_harness.py: reference code + RNG
injection) for a seeded action sequence: per-step grid, agent state, uint8 observations,
rewards before/after contract transfers, transfers, infos and feature_obs, plus episode
metrics.  Both the C oracle (CPU tests) and the CUDA path (GPU tests) must reproduce them
bit-exactly.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CRAMPED_CLEANUP = ["@@@@@@", "@PPPP@", "@PPPP@", "@HBBR@", "@PPPP@", "@@@@@@"]
CRAMPED_HARVEST = ["@@@@@@", "@PPPP@", "@PAAP@", "@PAAP@", "@PPPP@", "@@@@@@"]
OPEN_CLEANUP = [" PBH", "PRB ", "HPPB", "BPRP", "PPHP"]          # no walls: exercises map-edge bounds

# name -> (kind, n, seed, env_id, map (None = stock), horizon, episodes, steps/episode, action ids, action probs)
SCENARIOS = {
    "cleanup_n2": ("cleanup", 2, 73907, 0, None, 1000, 1, 300, 9, None),
    "cleanup_n8": ("cleanup", 8, 73907, 12345, None, 1000, 1, 250, 9, None),
    "cleanup_n5_short_horizon": ("cleanup", 5, 11, 3, None, 40, 3, 40, 8, None),
    "harvest_n4": ("harvest", 4, 73907, 77, None, 1000, 1, 250, 8, None),
    "harvest_n8_short_horizon": ("harvest", 8, 5, 4000000000, None, 30, 3, 30, 7, None),
    "cleanup_cramped_n8": ("cleanup", 8, 42, 9, CRAMPED_CLEANUP, 1000, 1, 400, 9,
                           [.2, .2, .2, .2, .04, .04, .04, .06, .02]),
    "harvest_cramped_n8": ("harvest", 8, 43, 10, CRAMPED_HARVEST, 1000, 1, 400, 8,
                           [.22, .22, .22, .22, .03, .03, .03, .03]),
    "cleanup_open_n6": ("cleanup", 6, 44, 11, OPEN_CLEANUP, 1000, 1, 300, 9, None),
    "cleanup_n8_nocontract": ("cleanup", 8, 3, 5, None, 1000, 1, 120, 9, None),
}


def run_scenario(name):
    from .ref_harness import RefGridEnv
    kind, n, seed, env_id, amap, horizon, episodes, steps, act_hi, act_p = SCENARIOS[name]
    contract = not name.endswith("nocontract")
    ref = RefGridEnv(kind, n, seed, env_id, contract=contract, ascii_map=amap, horizon=horizon)
    b = ref.base
    ascii_map = amap if amap is not None else ["".join(ch.decode() for ch in row) for row in b.base_map]
    rng = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    rec = {k: [] for k in ("actions", "map", "pos", "ori", "obs", "rew", "done", "eaten_apples", "info1",
                           "feature_obs", "base_rew", "transfers")}
    rst = {k: [] for k in ("map", "pos", "ori", "obs", "theta")}
    metrics = []
    for ep in range(episodes):
        s = ref.reset()
        for k in ("map", "pos", "ori", "obs"):
            rst[k].append(s[k])
        rst["theta"].append(s.get("theta", np.float64(0.0)))
        for k in rec:
            rec[k].append([])
        for t in range(steps):
            a = rng.choice(act_hi, size=n, p=act_p).astype(np.int32)
            o = ref.step(a)
            rec["actions"][-1].append(a)
            for k in ("map", "pos", "ori", "obs", "rew", "eaten_apples", "feature_obs"):
                rec[k][-1].append(o[k])
            rec["done"][-1].append(o["done"])
            rec["info1"][-1].append(o["cleaned_squares"] if kind == "cleanup" else o["eaten_close_apples"])
            rec["base_rew"][-1].append(o.get("base_rew", o["rew"]))
            rec["transfers"][-1].append(o.get("transfers", np.zeros(n)))
        m = ref.metrics()
        metrics.append(m)
    out = {"kind": kind, "n": n, "seed": seed, "env_id": env_id, "horizon": horizon, "contract": contract,
           "ascii_map": np.array(ascii_map), "metric_keys": np.array(sorted(metrics[0].keys())),
           "metrics": np.array([[m[k] for k in sorted(m.keys())] for m in metrics], dtype=np.float64)}
    for k, v in rec.items():
        out[k] = np.array(v)
    for k, v in rst.items():
        out["reset_" + k] = np.array(v)
    return out


def main(names=None):
    os.makedirs(OUT, exist_ok=True)
    for name in (names or SCENARIOS):
        data = run_scenario(name)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **data)
        print("%-28s %7.1f KiB  steps=%s" % (name, os.path.getsize(path) / 1024, data["actions"].shape))


if __name__ == "__main__":
    main(sys.argv[1:] or None)
