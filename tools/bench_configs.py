#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs on one GPU (the headline config is bench.py's).

  harvest_new  n=4  E=16384   HarvestFeaturemodLocalContract          (configs[1])
  cleanup / harvest feature envs n=8, E=131072 per GPU (1M / 8 GPUs)  (configs[3])
  selfdrive    n=8  E=131072  SelfdriveContractDistprop               (configs[4])
Each line: agent-steps/s over K device-resident steps (random actions on device, masked re-resets of finished envs
inside the timed region), the step kernel's mean CUDA-event time, algorithmic bytes per agent-step and the
fraction of the measured HBM peak.  Usage: python tools/bench_configs.py [--steps K]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def run(name, env, n, nact, alg_bytes, steps, selfdrive=False, horizon=1000):
    E = env.E
    env.reset()
    acts = None

    def step(i):
        nonlocal acts
        if selfdrive:
            acts = env.random_actions(i, -0.1, 0.1, out=acts)
        else:
            acts = env.random_actions(i, nact, out=acts)
        return env.step(acts, extras=False)

    for i in range(30):
        step(i)
    env.reset()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    acting = 0
    for i in range(steps):
        if selfdrive:
            acts = env.random_actions(i, -0.1, 0.1, out=acts)
        else:
            acts = env.random_actions(i, nact, out=acts)
        ev[i][0].record()
        obs, rew, done, info = env.step(acts, extras=False)
        ev[i][1].record()
        if selfdrive:
            if i % 16 == 15:
                env.reset(done[:, -1])           # finished episodes restart (masked reset)
        elif (i + 1) % horizon == 0:
            env.reset(done)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    kms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    rate = E * n * steps / (ms * 1e-3)
    achieved = alg_bytes * E * n / (kms * 1e-3) / 1e9
    line = {"config": name, "envs": E, "agents": n, "steps": steps, "agent_steps_per_s": rate, "ms_per_step": ms / steps,
            "kernel_ms": kms, "alg_bytes_per_agent_step": alg_bytes, "achieved_GBps": achieved, "hbm_peak_GBps": peak(),
            "frac_of_measured_hbm": achieved / peak()}
    print(json.dumps(line), flush=True)
    return line


def run_views(E=131072, n=8, reps=20):
    """JointEnv output layouts and the policy-side consumer (ssd_views.cuh): HBM GB/s of each kernel at the headline
    batch size (inputs and outputs exceed the L2, so every launch streams from / to HBM)."""
    from contracts_b200.batched import BatchedGridEnv
    env = BatchedGridEnv("cleanup_new", E, n, contract="CleanupContract")
    env.reset()
    for i in range(5):
        env.step(env.random_actions(i, 8), extras=False)
    gv = torch.empty((E, env.H, env.W, 3), dtype=torch.uint8, device=env.device)
    cc = torch.empty((E, 15, 15, 3 * n), dtype=torch.uint8, device=env.device)
    cases = [("global_view_kernel (JointEnv global_obs)", lambda: env.global_view(gv), env.state_map_bytes + 32 + env.H * env.W * 3),
             ("concat_obs_kernel (JointEnv concatenated_obs)", lambda: env.concatenated_obs(cc), 2 * n * 675)]
    for dt, size in ((torch.float32, 4), (torch.bfloat16, 2)):
        img = torch.empty((E * n, 3, 15, 15), dtype=dt, device=env.device)
        con = torch.empty((E * n, 10), dtype=dt, device=env.device)
        cases.append(("policy_inputs_kernel %s (VisionNetwork preprocessing)" % str(dt).split(".")[1],
                      lambda img=img, con=con, dt=dt: env.policy_inputs(dt, image=img, contract=con), n * 675 * (1 + size) + n * 10 * size + 8))
    for name, fn, bytes_per_env in cases:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = bytes_per_env * E / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "envs": E, "agents": n, "kernel_ms": ms, "alg_bytes_per_env": bytes_per_env,
                          "achieved_GBps": gbs, "hbm_peak_GBps": peak(), "frac_of_measured_hbm": gbs / peak()}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    args = ap.parse_args()
    from contracts_b200.batched import BatchedGridEnv
    from contracts_b200.features import BatchedFeatureEnv
    from contracts_b200.selfdrive import BatchedCarEnv
    K = args.steps
    run_views()
    # algorithmic bytes per agent-step: observation + state read/write + rewards/actions/infos
    run("harvest_new n=4 E=16384 HarvestFeaturemodLocalContract", BatchedGridEnv("harvest_new", 16384, 4, contract="HarvestFeaturemodLocalContract"),
        4, 7, (2700 + 2 * 1008 + 4 * (1 + 8 + 4) + 1) / 4, K)
    run("harvest_new n=8 E=131072 HarvestFeaturemodLocalContract", BatchedGridEnv("harvest_new", 131072, 8, contract="HarvestFeaturemodLocalContract"),
        8, 7, (5400 + 2 * 1008 + 8 * (1 + 8 + 4) + 1) / 8, K)
    run("cleanup (features) n=8 E=131072 CleanupContract", BatchedFeatureEnv("cleanup", 131072, 8, contract="CleanupContract"),
        8, 8, (8 * 20 * 8 + 2 * (64 + 32 + 16 + 8 * 36) + 8 * 13 + 1) / 8, K)
    run("harvest (features) n=8 E=131072 HarvestFeaturemodLocalContract", BatchedFeatureEnv("harvest", 131072, 8, contract="HarvestFeaturemodLocalContract"),
        8, 7, (8 * 26 * 8 + 2 * (64 + 32 + 16 + 8 * 36) + 8 * 13 + 1) / 8, K)
    run("selfdrive n=8 E=131072 SelfdriveContractDistprop", BatchedCarEnv(131072, 8, contract="SelfdriveContractDistprop"),
        8, 0, (8 * 21 * 8 + 2 * (16 * 8 + 44) + 8 * (4 + 8) + 9) / 8, K, selfdrive=True)
    run("selfdrive n=8 E=1048576 SelfdriveContractDistprop", BatchedCarEnv(1048576, 8, contract="SelfdriveContractDistprop"),
        8, 0, (8 * 21 * 8 + 2 * (16 * 8 + 44) + 8 * (4 + 8) + 9) / 8, K, selfdrive=True)
    run("cleanup (features) n=8 E=1048576 CleanupContract", BatchedFeatureEnv("cleanup", 1048576, 8, contract="CleanupContract"),
        8, 8, (8 * 20 * 8 + 2 * (64 + 32 + 16 + 8 * 36) + 8 * 13 + 1) / 8, K)


if __name__ == "__main__":
    main()
