import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from traffic_env_b200 import VecTrafficEnv
E=16384
env = VecTrafficEnv(m=10, n=10, length=500.0, num_envs=E, local_cars_per_sec=0.12, arrivals="philox", seed=2026, ticks_per_step=10, remi=True)
env.reset()
I=env.intersections
h_act=np.zeros((E,I),np.uint8)
for s in range(300):
    if s%3==0: h_act[:]=env.greedy_actions()
    env.step(h_act)
def timeit(f,n=30):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for i in range(n): f(i)
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print("step only           %.3f ms"%timeit(lambda i: env.step(h_act)))
print("last kernel span    %.3f ms"%env.last_kernel_ms())
print("greedy_actions      %.3f ms"%timeit(lambda i: env.greedy_actions()))
def hs(i):
    if i%3==0: h_act[:]=env.greedy_actions()
    o,r,d=env.step(h_act); return float(r[0,0])+float(o[0,0])
print("bench host_step     %.3f ms"%timeit(hs))
d_act=torch.zeros((E,I),dtype=torch.uint8,device='cuda'); d_obs=torch.empty((E,env.obs_len),dtype=torch.float32,device='cuda'); d_rew=torch.empty((E,I),dtype=torch.float32,device='cuda'); d_done=torch.empty((E,),dtype=torch.uint8,device='cuda')
print("device step         %.3f ms"%timeit(lambda i: env.step_device(d_act,d_obs,d_rew,d_done)))
