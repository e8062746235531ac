import sys, os, torch
sys.path.insert(0, os.getcwd())
from collision_avoidance_b200 import envs
Es = [int(x) for x in sys.argv[1:]] or [100000, 320, 64, 1000]
for E in Es:
    try:
        env = envs.Collision_Avoidance_Env(numAgents=10, num_envs=E, seed=3)
        torch.cuda.synchronize()
        th = torch.zeros(E, 10, device="cuda")
        for _ in range(5):
            env.step(th)
        torch.cuda.synchronize()
        print("E", E, "ok", float(env.obs.abs().sum()))
    except Exception as ex:
        print("E", E, "FAIL", str(ex)[:200]); break
