"""Second hunt for an observable effect of the reference's stored-node rewrite (cpp/parallel_weighted_astar.cpp:255-257): 2400 small
oracle searches (puzzle15, cube3; batch 3 / 8 / 20; weights 0.2 - 0.5) under a RANDOM heuristic (a hash of the state, 0 at the goal), with
and without the rewrite.  Result (recorded in DESIGN.md section 2): 0 of 2400 differ in moves or nodes generated.  CPU only."""
import random, sys, numpy as np, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import oracle_env as O
from oracle.oracle_bwas import bwas
def rnd_h(env, amp, seed):
    wj = (np.arange(env.state_dim) * 2654435761 % 1000003 + 1 + seed).astype(np.int64)
    goal = env.goal
    def h(states):
        r = ((states.astype(np.int64) * wj[None]).sum(axis=1) * 2654435761 % 1000003).astype(np.float32) / np.float32(1000003.0)
        out = (r * np.float32(amp)).astype(np.float32)
        out[(states == goal[None]).all(axis=1)] = 0
        return out
    return h
t0=time.time(); tot=0; diff=0; found=[]
for name,back in [("puzzle15",(4,10)),("cube3",(2,5))]:
    env=O.get_oracle_env(name)
    for seed in range(400):
        np.random.seed(seed); random.seed(seed)
        st,_=env.generate_states(1,back)
        for batch,w,amp in [(3,0.3,4.0),(8,0.5,6.0),(20,0.2,5.0)]:
            h=rnd_h(env,amp,seed)
            a=bwas(env,st[0],h,w,batch,mutate_stored=True,max_iters=400)
            b=bwas(env,st[0],h,w,batch,mutate_stored=False,max_iters=400)
            tot+=1
            if a["moves"]!=b["moves"] or a["nodes_generated"]!=b["nodes_generated"]:
                diff+=1; found.append((name,seed,batch,w,amp,a["moves"],b["moves"],a["nodes_generated"],b["nodes_generated"],a["links"]))
                print("DIFF",found[-1],flush=True)
        if time.time()-t0>1500: break
    print(name,"cases",tot,"diff",diff,"t",time.time()-t0,flush=True)
