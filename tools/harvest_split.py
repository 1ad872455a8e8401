import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from contracts_b200.batched import BatchedGridEnv
for E, n in ((131072, 8), (16384, 4)):
    env = BatchedGridEnv("harvest_new", E, n, contract="HarvestFeaturemodLocalContract")
    env.reset()
    acts = None
    for i in range(200):
        acts = env.random_actions(i, 7, out=acts); env.step(acts, extras=False)
    env.enable_timing(True)
    acc = []
    for i in range(200, 300):
        acts = env.random_actions(i, 7, out=acts); env.step(acts, extras=False); torch.cuda.synchronize(); acc.append(env.step_times_ms())
    print(E, n, "logic %.4f  obs+reward %.4f ms" % (np.mean([a for a, _ in acc]), np.mean([b for _, b in acc])))
