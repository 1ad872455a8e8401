import sys, torch
sys.path.insert(0, "/root/repo")
from contracts_b200.selfdrive import BatchedCarEnv
env = BatchedCarEnv(131072, 8, contract="SelfdriveContractDistprop")
env.reset()
acts = None
for i in range(120):
    acts = env.random_actions(i, -0.1, 0.1, out=acts)
    obs, rew, done, info = env.step(acts, extras=False)
    if i % 16 == 15:
        env.reset(done[:, -1])
torch.cuda.synchronize()
