import sys, torch
sys.path.insert(0, "/root/repo")
from contracts_b200.features import BatchedFeatureEnv
env = BatchedFeatureEnv("cleanup", 131072, 8, contract="CleanupContract")
env.reset()
acts = None
for i in range(300):
    acts = env.random_actions(i, 8, out=acts)
    env.step(acts, extras=False)
torch.cuda.synchronize()
