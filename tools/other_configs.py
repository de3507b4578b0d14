"""Timings of the other BASELINE configurations (profiles/README.md table): solve-only formation batches,
20-piece problems on the condensed and on the pivoted path, and the pose-batch collision query."""
import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drone_path_planning_python_b200 as mst
rng=np.random.default_rng(1)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
# config 2: 4096 formations x 5 drones x 10 x 3 ; scaled x64 to get a measurable time
for F in (4096, 262144):
    D,n,K=5,10,3
    T=rng.uniform(0.5,2,(F,n)); t=torch.as_tensor(np.concatenate([np.zeros((F,1)),np.cumsum(T,1)],1),device='cuda')
    wp=torch.as_tensor(np.cumsum(rng.normal(0,0.3,(F*D,n+1,K)),1),device='cuda')
    ms=timeit(lambda: mst.solve_batch(wp,t,share_time_group=D))
    print("config2-style F=%d: %.3f ms -> %.1f M trajectories/s" % (F, ms, F*D/ms/1e3))
# config 3: 65536 x 20 pieces, wide spreads (pivoted solver)
B,n,K=65536,20,3
T0=rng.uniform(0.5,2,(B,n)); xi=rng.normal(size=(B,n))
wp=torch.as_tensor(np.cumsum(rng.normal(0,0.3,(B,n+1,K)),1),device='cuda')
for r in (0,4):
    T=np.clip(T0*np.exp(0.25*r*xi),0.05,5.0)
    t=torch.as_tensor(np.concatenate([np.zeros((B,1)),np.cumsum(T,1)],1),device='cuda')
    ms=timeit(lambda: mst.solve_batch(wp,t),3)
    print("config3 r=%d (share pivoted %.2f): %.2f ms -> %.2f M solves/s" % (r,(T.max(1)/T.min(1)>4).mean(), ms, B/ms/1e3))
# config 4: 1M poses
from bench import mesh_soups
robot_s, env_s = mesh_soups()
robot,env=mst.Mesh(robot_s),mst.Mesh(env_s)
P=1<<20
flat=env_s.reshape(-1,3)
poses=torch.as_tensor(np.concatenate([rng.uniform(flat.min(0)-0.8,flat.max(0)+0.8,(P,3)),rng.uniform(-np.pi,np.pi,(P,1))],1),device='cuda')
ms=timeit(lambda: mst.collide_poses(robot,env,poses))
print("config4 1M random (x,y,z,yaw) poses: %.3f ms -> %.0f M poses/s" % (ms, P/ms/1e3))
# config 2 with its re-solves: one iteration of the time-allocation search = 6 line-search solves + costs +
# gradients per problem (the gradient comes from the coefficients: mst_time_gradient)
B,n,K=65536,20,3
T=rng.uniform(0.6,1.6,(B,n)); t=torch.as_tensor(np.concatenate([np.zeros((B,1)),np.cumsum(T,1)],1),device='cuda')
wp=torch.as_tensor(np.cumsum(rng.normal(0,1.0,(B,n+1,K)),1),device='cuda')
iters=4
mst.optimize_time_allocation(wp,t,iters=1); torch.cuda.synchronize()
dt=1e9
for _ in range(3):   # wall clock (the search is host-driven): best of three
    t0=time.perf_counter(); tn,cost=mst.optimize_time_allocation(wp,t,iters=iters); torch.cuda.synchronize(); dt=min(dt,time.perf_counter()-t0)
print("config2 time-allocation search: %d problems x %d pieces, %d iterations: %.1f ms (%.1f ms / iteration, %.1f M re-solves/s); median cost ratio %.3f" % (
    B, n, iters, dt*1e3, dt*1e3/iters, B*6*iters/dt/1e6, float((cost[-1]/cost[0]).median())))
# the benchmark's pipeline with the yaw axis as well (K = 4: rotated culls instead of the translation tables)
B,n,K,S=1<<20,10,4,100
T=rng.uniform(0.5,2,(B,n)); t=torch.as_tensor(np.concatenate([np.zeros((B,1)),np.cumsum(T,1)],1),device='cuda')
wp4=np.zeros((B,n+1,K)); wp4[:,:,:3]=rng.uniform([-2.2,2.8,0.5],[2.2,5.0,2.5],(B,1,3))+np.cumsum(rng.normal(0,0.3,(B,n+1,3)),1)
wp4[:,:,3]=np.cumsum(rng.normal(0,0.1,(B,n+1)),1)
wp4=torch.as_tensor(wp4,device='cuda')
out=mst.pipeline(wp4,t,S,robot,env)
ms=timeit(lambda: mst.pipeline(wp4,t,S,robot,env,out=out),5)
print("pipeline K=4 (x,y,z,yaw), 1M trajectories: %.2f ms -> %.1f M trajectories/s, any-hit rate %.2f" % (ms, B/ms/1e3, float(out.any_hit.float().mean())))
