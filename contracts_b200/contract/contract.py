"""Contract base type — same fields as the reference's contract/contract.py:1-9."""


class Contract:
    def __init__(self, contract_space, default_contract, num_agents):
        self.contract_space = contract_space
        self.default_contract = default_contract
        self.num_agents = num_agents

    def compute_transfer(self, obs, acts, rews, params, infos=None):
        raise NotImplementedError
