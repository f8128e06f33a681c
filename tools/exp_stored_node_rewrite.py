"""Does the reference's in-place rewrite of the stored CLOSED node (cpp/parallel_weighted_astar.cpp:255-257) ever change a result?
60 searches of the oracle (cube3 / puzzle15, several weights and batch sizes) under a strongly INCONSISTENT heuristic, with and without
the rewrite, and the "min" in-batch rule of r01 against the reference's sequential rule.  CPU only (oracle = test infrastructure).
Result recorded in DESIGN.md section 2: 8,514 rewrites, moves / nodes generated identical in 60/60; min-vs-sequential differs in 1/60."""
import random, sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import oracle_env as O
from oracle.oracle_bwas import bwas, misplaced_heuristic
def noisy(env, amp):
    base = misplaced_heuristic(env)
    wj = (np.arange(env.state_dim) * 2654435761 % 1000003 + 1).astype(np.int64)
    def h(states):
        r = ((states.astype(np.int64) * wj[None]).sum(axis=1) * 2654435761 % 1000003).astype(np.float32) / np.float32(1000003.0)
        return (base(states)*np.float32(1.5) + r*np.float32(amp)).astype(np.float32)
    return h
tot=0; dm=0; dn=0; dl=0; links=0; dmin=0; dminlen=0
for name,back,batch,w,amp in [("cube3",(5,10),50,0.8,3.0),("cube3",(6,11),200,0.6,4.0),("puzzle15",(15,40),50,0.8,4.0),("cube3",(5,9),20,0.2,3.0),("puzzle15",(15,40),100,0.3,6.0)]:
    env=O.get_oracle_env(name); h=noisy(env,amp)
    np.random.seed(3); random.seed(3)
    st,_=env.generate_states(12,back)
    for s in st:
        a=bwas(env,s,h,w,batch,mutate_stored=True,max_iters=300)
        b=bwas(env,s,h,w,batch,mutate_stored=False,max_iters=300)
        c=bwas(env,s,h,w,batch,batch_dedup="min",mutate_stored=False,max_iters=300)
        tot+=1; links+=a["links"]
        if a["moves"]!=b["moves"]: dm+=1
        if a["nodes_generated"]!=b["nodes_generated"]: dn+=1
        if a["moves"] is not None and b["moves"] is not None and len(a["moves"])!=len(b["moves"]): dl+=1
        if c["nodes_generated"]!=b["nodes_generated"] or c["moves"]!=b["moves"]: dmin+=1
        if c["moves"] is not None and b["moves"] is not None and len(c["moves"])!=len(b["moves"]): dminlen+=1
    print(name,back,batch,w,"cases",tot,"links",links,"moves differ",dm,"nodes differ",dn,"len differ",dl,"min-vs-seq differ",dmin,"len",dminlen,flush=True)
