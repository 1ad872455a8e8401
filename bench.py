#!/usr/bin/env python
"""bench.py — throughput of the hot path on B200 (and the CPU reference arm).

Metric (BASELINE.json): agent-steps/s, cleanup_new, 8 agents, observations included.
Workload (`config.workload`): BASELINE configs[2] — CleanupEnv(num_agents=8) + CleanupContract with the
two-stage negotiation prologue each episode, horizon 1000, iid uniform random actions, E envs
per GPU (weak scaling: every rank owns E envs with global ids rank*E .. rank*E+E-1; no
collective on the step path; one NCCL all-gather of episode statistics per episode).

A "step" = one `ssd_step` launch over the rank's E envs (+ the action-generation launch).
  value  : device-resident (actions generated on device, outputs stay in HBM), CUDA-event timed
  e2e    : same step through the public API with HOST buffers: pinned actions -> H2D, step,
           rewards + dones -> D2H, host sync every step (observations stay in the policy's
           device batch tensor, as designed; `e2e_obs_to_host` additionally copies them out)
  roofline: algorithmic bytes (840 B/agent-step) / measured ssd_step duration
  cpu_baseline / --impl reference: the C oracle (port of the reference algorithm) on host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_AGENTS = 8
HORIZON = 1000
N_ACTIONS = 8                      # cleanup_new with disable_firing (cleanup_new.py:90-92): Discrete(8)
ALG_BYTES_PER_AGENT_STEP = 840.0
SEED = 73907                       # reference seed multiplier (runner.py:130)
WORKLOAD = "cleanup_new n=8 CleanupContract + negotiation prologue, horizon 1000, random actions"


def ncu_traffic(envs, agents):
    """DRAM bytes per launch of the step kernel from the committed ncu capture (same workload and size), else None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if t["envs"] == envs and t["agents"] == agents:
            return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def _nvml_sample(self):
        nv, h = self.nv, self.nvml_handle
        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        r = int(self.get_reasons(h))
        self.rows.append([str(sm), str(self.mx), "0"] + ["Active" if r & b else "Not Active" for _, b in self.bits])

    def _nvml_loop(self):
        while not self.stop_flag.wait(0.005):
            self._nvml_sample()

    def start(self):
        """NVML polled every 5 ms from a thread (the timed region can be shorter than one nvidia-smi period), first
        sample taken synchronously, last one at stop(); nvidia-smi -lms as the fallback."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.nv, self.nvml_handle = nv, nv.nvmlDeviceGetHandleByIndex(idx)
            names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                     ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                     ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                     ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))
            self.bits = [(n, getattr(nv, a, None) or getattr(nv, b, 0)) for n, a, b in names]
            self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.nvml_handle, nv.NVML_CLOCK_SM))
            self._nvml_sample()
            self.stop_flag = threading.Event()
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            self.proc = "nvml"
            return
        except Exception:
            self.proc = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.proc == "nvml":
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            try:
                if len(self.rows) < 3:               # a region shorter than two polling periods
                    self._nvml_sample()
            except Exception:
                pass
        else:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def run_cpu(envs_per_thread, steps, warmup):
    """The oracle port on all host cores (OpenMP).  Returns (agent-steps/s, cores, sample string)."""
    from oracle import oracle
    from contracts_b200.maps import CLEANUP_MAP
    cores = oracle.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: ask for every core
    E = envs_per_thread * cores
    o = oracle.GridOracle("cleanup", E, N_AGENTS, CLEANUP_MAP, horizon=HORIZON, contract="CleanupContract", seed=SEED)
    o.reset()
    rng = np.random.RandomState(0)
    acts = rng.randint(0, N_ACTIONS, size=(16, E, N_AGENTS)).astype(np.int32)
    for t in range(warmup):
        o.step(acts[t % 16], want_features=False)
    t0 = time.perf_counter()
    for t in range(steps):
        o.step(acts[t % 16], want_features=False)
    dt = time.perf_counter() - t0
    return E * N_AGENTS * steps / dt, cores, dt / steps * 1e3, \
        "C oracle port, OpenMP x%d, %d envs x %d steps of the same workload (no negotiation prologue, no feature_obs)" % (cores, E, steps)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # ~10-30 s of CPU work: 64 envs per core, steps scaled so the run is bounded
    steps, warm = max(args.steps, 1), max(args.warmup, 3)
    per_core = 64
    budget_steps = 300
    if steps > budget_steps:
        steps_run = budget_steps
    else:
        steps_run = steps
    v, cores, ms, sample = run_cpu(per_core, steps_run, min(warm, 50))
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": v, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": steps_run, "warmup": min(warm, 50), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs": per_core * cores, "agents": N_AGENTS, "horizon": HORIZON},
        "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference itself is Python (cannot travel to the GPU box); this is its C restatement (oracle/) "
                "on all host cores — ~100x faster per core than the Python reference (BASELINE.md: 3.9k/s/core)",
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=1000)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=131072, help="envs per GPU")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--graph-steps", type=int, default=50, help="env steps per captured CUDA graph (1 = no graphs)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from contracts_b200.batched import BatchedGridEnv
    from contracts_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)
    E, n = args.envs, N_AGENTS
    K, W = args.steps, max(args.warmup, 3)

    env = BatchedGridEnv("cleanup_new", E, n, horizon=HORIZON, contract="CleanupContract", seed=SEED,
                         first_env_id=rank * E, device=dev)
    actions = torch.empty((E, n), dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED + rank)

    def new_episode():
        """reset + negotiation prologue (two_stage_train.py:257-281): a0 proposes theta ~ U[0, 0.2], others accept ~ U[0,1]."""
        env.reset()
        proposals = torch.rand((E,), dtype=torch.float64, device=dev, generator=gen) * 0.2
        accept = torch.rand((E, n), dtype=torch.float64, device=dev, generator=gen)
        env.negotiate(proposals, accept)

    def end_episode():
        """episode statistics -> one small NCCL all-gather (the only collective of the workload)."""
        g = sharding.gather_episode_stats(sharding.local_episode_stats(env.metrics_raw()))
        total = g.sum(0)
        total[7] = g[:, 7].max()
        return total

    state = {"t": 0, "global_step": 0, "stats": None}

    def one_step(device_actions=True, host_actions=None, host_out=None):
        if state["t"] == 0:
            new_episode()
        if device_actions:
            env.random_actions(state["global_step"], N_ACTIONS, out=actions)
            env.step(actions, extras=False)
        elif host_out is not None:      # the C ABI's host-buffer step: H2D actions, step, D2H rewards + dones, synchronous
            env.step_host(host_actions, host_out[0], host_out[1])
        else:
            actions.copy_(host_actions, non_blocking=True)
            env.step(actions, extras=False)
        state["t"] += 1
        state["global_step"] += 1
        if state["t"] == HORIZON:
            state["stats"] = end_episode()
            state["t"] = 0

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up, then the timed device-resident run -------------------------------------------
    # The inner loop (action generation + step) is launch-bound from Python, so GRAPH_STEPS consecutive steps are
    # captured once in a CUDA graph and replayed; episode boundaries (reset + negotiation prologue + statistics
    # all-gather) stay outside the graph.  Fresh actions every replay come from the handle's device step counter.
    G = max(1, args.graph_steps)
    while HORIZON % G or K % G:                                  # the timed region is EXACTLY K steps: G divides K
        G -= 1
    for _ in range(max(W, 3)):
        one_step()
    barrier()
    graph = None
    if G > 1:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):                                   # warm the capture stream
                env.random_actions(None, N_ACTIONS, out=actions)
                env.step(actions, extras=False)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        lcap = env.kernel_launches
        with torch.cuda.graph(graph, stream=side):
            for _ in range(G):
                env.random_actions(None, N_ACTIONS, out=actions)
                env.step(actions, extras=False)
        torch.cuda.synchronize(dev)
        per_graph = env.kernel_launches - lcap                   # kernels captured in one graph (counted by the C ABI)
    state["t"] = 0                                               # start the timed region at an episode boundary

    def run_steps(k):
        """k env steps (k a multiple of G when graphs are on), crossing episode boundaries as needed."""
        done = 0
        launches = 0
        while done < k:
            if state["t"] == 0:
                new_episode()
            if graph is not None:
                graph.replay()
                m, launches = G, launches + per_graph
            else:
                env.random_actions(state["global_step"], N_ACTIONS, out=actions)
                env.step(actions, extras=False)
                m = 1
            done += m
            state["t"] += m
            state["global_step"] += m
            if state["t"] >= HORIZON:
                state["stats"] = end_episode()
                state["t"] = 0
        return launches

    K = max(G, (K // G) * G)
    # The episode-end statistics all-gather is the workload's only collective: run it once untimed so that NCCL's lazy
    # communicator set-up (tens of ms) never lands inside a timed region whose warm-up is shorter than an episode.
    end_episode()
    run_steps(max(G, (W // G) * G))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = env.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    graph_launches = run_steps(K)
    ev1.record()
    barrier()
    launches = (env.kernel_launches - l0) + graph_launches       # eager launches are counted by the C ABI itself
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    value = world * E * n * K / (ms_total * 1e-3)

    # ---- the step kernel alone (roofline): CUDA events around each ssd_step of an un-graphed stretch -------
    Kk = HORIZON                     # one whole episode: the cost of a step varies along it (spawning starts late)
    state["t"] = 0
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kk)]
    for k in range(Kk):
        if state["t"] == 0:
            new_episode()
        env.random_actions(state["global_step"], N_ACTIONS, out=actions)
        kev[k][0].record()
        env.step(actions, extras=False)
        kev[k][1].record()
        state["t"] += 1
        state["global_step"] += 1
        if state["t"] == HORIZON:
            state["stats"] = end_episode()
            state["t"] = 0
    barrier()
    step_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    # per-kernel split of the step (library-side CUDA events around each kernel), sampled every 10th step of one more episode
    kern_ms = None
    try:
        env.enable_timing(True)
        acc = []
        for k in range(HORIZON):
            if state["t"] == 0:
                new_episode()
            env.random_actions(state["global_step"], N_ACTIONS, out=actions)
            env.step(actions, extras=False)
            if k % 10 == 5:
                acc.append(env.step_times_ms())
            state["t"] += 1
            state["global_step"] += 1
            if state["t"] == HORIZON:
                state["stats"] = end_episode()
                state["t"] = 0
        kern_ms = (float(np.mean([a for a, _ in acc])), float(np.mean([b for _, b in acc])))
    except Exception as exc:                                     # single-kernel fallback (SSD_GRID_KERNEL=v3) has no split
        print("per-kernel timing unavailable: %s" % exc, file=sys.stderr)
    finally:
        env.enable_timing(False)
    barrier()

    # ---- e2e: host buffers in, host results out, sync every step ----------------------------------
    Ke = min(args.e2e_steps, K)
    host_actions = [torch.randint(0, N_ACTIONS, (E, n), dtype=torch.uint8).pin_memory() for _ in range(4)]
    host_rew = torch.empty((E, n), dtype=torch.float64).pin_memory()
    host_done = torch.empty((E,), dtype=torch.uint8).pin_memory()
    for i in range(3):
        one_step(False, host_actions[i % 4], (host_rew, host_done))
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(Ke):
        one_step(False, host_actions[i % 4], (host_rew, host_done))     # returns with host_rew / host_done valid
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * E * n * Ke / (float(e2e_ms.item()) * 1e-3)
    # informational: also ship the observations to the host (PCIe-bound by construction)
    obs_to_host = None
    if rank == 0:
        host_obs = torch.empty(env._obs_buf.shape, dtype=torch.uint8).pin_memory()
        Ko = 10
        torch.cuda.synchronize(dev)
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record()
        for i in range(Ko):
            one_step(False, host_actions[i % 4])
            host_rew.copy_(env.rew, non_blocking=True)
            host_obs.copy_(env._obs_buf, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        o1.record()
        torch.cuda.synchronize(dev)
        obs_to_host = E * n * Ko / (o0.elapsed_time(o1) * 1e-3)
    barrier()

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = ALG_BYTES_PER_AGENT_STEP * E * n / (step_ms * 1e-3) / 1e9
        line = {
            "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": E, "agents": n, "horizon": HORIZON,
                       "l2": "working set %.0f MB/step (obs %.0f MB + state) exceeds the 126 MB L2" % (
                           (E * (n * 675 + 2 * env.state_bytes_per_env)) / 1e6, E * n * 675 / 1e6),
                       "parallelism": "env batch sharded over %d GPU(s), no collective on the step path" % world,
                       "cuda_graph_steps": G},
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": E * n,
                    "d2h_bytes_per_step": E * n * 8 + E, "steps": Ke,
                    "note": "ssd_step_host: pinned host actions in, rewards+dones out, synchronous every step (the D2H copy overlaps the observe kernel); observations stay in the device batch tensor"},
            "e2e_obs_to_host": {"value": obs_to_host, "unit": "agent-steps/s", "d2h_bytes_per_step": E * n * 675 + E * n * 8},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(E, n), "alg_bytes_per_launch": ALG_BYTES_PER_AGENT_STEP * E * n,
                         "peak_source": peak_src,
                         "kernel": "one env step = grid_logic_kernel + grid_obs_kernel (cleanup), CUDA events around ssd_step",
                         "kernel_ms": step_ms, "alg_bytes_per_agent_step": ALG_BYTES_PER_AGENT_STEP},
        }
        if kern_ms is not None:
            # the dominant kernel on its own: observation stream + the map / header words it reads (+ the map when the
            # spawn changed it, ~1/3 of the steps): 5400 + 464 + 56 + 464/3 B per env
            obs_alg = (n * 675 + env.state_map_bytes + 56 + env.state_map_bytes / 3.0) * E
            line["roofline"]["kernels"] = [
                {"name": "grid_logic_kernel", "ms": kern_ms[0]},
                {"name": "grid_obs_kernel", "ms": kern_ms[1], "alg_bytes_per_launch": obs_alg,
                 "achieved": obs_alg / (kern_ms[1] * 1e-3) / 1e9, "frac": obs_alg / (kern_ms[1] * 1e-3) / 1e9 / peak}]
        if state["stats"] is not None:
            s = state["stats"].tolist()
            line["episode_stats"] = {"apples_eaten": s[0], "raw_env_rewards": s[1], "transfers": s[2],
                                     "dirt_cleaned": s[3], "envs": s[6], "err_flags": s[7]}
        if world == 1 and not args.no_cpu:
            v, cores, ms, sample = run_cpu(64, 150, 20)
            line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
