"""Contract definitions, name- and signature-compatible with the reference's contract/contract_list.py.

On the device the transfer rule of each class is a fused function inside the step kernel
(`contracts_b200/csrc/ssd_grid.cuh`, selected by `type(contract).__name__`); the
`compute_transfer` methods here keep the reference's host-side interface for callers that
evaluate a contract on dict data themselves (they are never on the env step path).
"""
import numpy as np

from ..spaces import Box
from .contract import Contract


class CleanupContract(Contract):
    """theta in [0, 0.2]: payment per waste cell cleaned, paid evenly by the others (contract_list.py:7-27)."""

    def __init__(self, num_agents, low_val=0, high_val=0.2):
        super().__init__(Box(shape=(1,), low=low_val, high=high_val), np.array([0.0]), num_agents)

    def compute_transfer(self, obs, acts, rews, params, infos=None):
        return {k: -params[k][0] * infos[k]["cleaned_squares"] for k in acts.keys()}


class HarvestFeaturemodLocalContract(Contract):
    """theta in [0, 10]: eating an apple in a low-density region transfers theta (contract_list.py:29-54)."""

    def __init__(self, num_agents, low_val=0, high_val=10.0):
        super().__init__(Box(shape=(1,), low=low_val, high=high_val), np.array([0.0]), num_agents)

    def compute_transfer(self, obs, acts, rews, params, infos=None):
        out = {}
        for k in acts.keys():
            low = infos[k]["feature_obs"][8] < 4 and infos[k]["eaten_close_apples"] > 0
            out[k] = params[k][0] if low else 0
        return out


class SelfdriveContractDistprop(Contract):
    """theta in [0, 100]: distance-proportional transfers when the ambulance merges (contract_list.py:56-102)."""

    def __init__(self, num_agents):
        super().__init__(Box(shape=(1,), low=0, high=100.0), np.array([0.0]), num_agents)

    def compute_transfer(self, obs, acts, rews, params, infos=None):
        transfers = {"a0": 0}
        n_slots = len(list(obs.values())[0]) // 2
        if "a0" in acts and infos["a0"]["just_passed"]:
            behind = [i for i in range(1, n_slots) if obs["a0"][2 + i] < 0]
            if behind:
                dists = {"a%d" % i: -obs["a0"][2 + i] for i in behind}
                total = 0
                for i in behind:
                    total += dists["a%d" % i]
                transfers["a0"] = (params["a0"][0] * total, {k: d / total for k, d in dists.items()})
            for i in range(1, n_slots):
                k = "a%d" % i
                if k in acts and i not in behind:
                    transfers[k] = (params["a0"][0] * obs["a0"][2 + i], {"a0": 1})
        for i in range(1, n_slots):
            transfers.setdefault("a%d" % i, 0)
        return transfers
