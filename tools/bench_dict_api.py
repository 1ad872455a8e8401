#!/usr/bin/env python
"""Cost of the dict APIs (GPU box): env-steps/s of (1) the single-env drop-in façade (`env_creator('CleanupNew')` +
`ContractWrapperSubgame`, one env, dict in / dict out), (2) `SSDVectorEnv.poll()/send_actions()` with the full dict
layer, (3) `SSDVectorEnv.poll_arrays()/send_action_array()` (one copy each way, no dicts), next to the reference's Python
loop on the same box when it is staged (oracle/_ref).  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from contracts_b200.contract import contract_list
    from contracts_b200.utils.env_creator_functions import env_creator
    from contracts_b200.vector_env import SSDVectorEnv
    n, out = 8, {}
    base = env_creator("CleanupNew", dict(num_agents=n, env_params={}, image_obs=True))
    env = env_creator("ContractWrapperSubgame", dict(num_agents=n, base_env=base, contract=contract_list.CleanupContract(n), convolutional=True))
    env.reset()
    rng = np.random.RandomState(0)
    keys = ["a%d" % i for i in range(n)]
    for steps in (20, 300):
        t0 = time.perf_counter()
        for _ in range(steps):
            env.step({k: int(rng.randint(8)) for k in keys})
        dt = time.perf_counter() - t0
    out["facade_single_env"] = {"env_steps_per_s": steps / dt, "ms_per_env_step": dt / steps * 1e3}
    for E in (256, 4096):
        v = SSDVectorEnv("cleanup_new", E, n, contract="CleanupContract")
        v.poll()
        acts = {e: {k: int(rng.randint(8)) for k in keys} for e in range(E)}
        K = 5 if E > 1000 else 20
        t0 = time.perf_counter()
        for _ in range(K):
            v.send_actions(acts)
            v.poll()
        dt = time.perf_counter() - t0
        out["vector_dicts_E%d" % E] = {"env_steps_per_s": E * K / dt, "ms_per_poll": dt / K * 1e3}
        a = np.random.randint(0, 8, size=(E, n)).astype(np.uint8)
        K = 200
        v.send_action_array(a); v.poll_arrays()
        t0 = time.perf_counter()
        for _ in range(K):
            v.send_action_array(a)
            s = v.poll_arrays()
            done = np.nonzero(s["done"])[0]
            for e in done:
                v.try_reset(int(e))
        dt = time.perf_counter() - t0
        out["vector_arrays_E%d" % E] = {"env_steps_per_s": E * K / dt, "ms_per_poll": dt / K * 1e3}
        v.stop()
    try:
        from oracle import ref_bench
        if ref_bench.available():
            rate, procs, ms, sample = ref_bench.run("cleanup8", steps=100, warmup=3, procs=1)
            out["reference_python_1proc"] = {"env_steps_per_s": rate / n, "ms_per_env_step": ms}
    except Exception as exc:
        out["reference_python_1proc"] = {"error": str(exc)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
