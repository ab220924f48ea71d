import sys; sys.path.insert(0,"."); sys.path.insert(0,"tests")
import numpy as np
from muzero_jl_b200 import capi
import common
from oracle import oracle as O
mode = int(sys.argv[1]); slots=int(sys.argv[2]); S=int(sys.argv[3]); games=int(sys.argv[4])
ctx=capi.Context(capi.default_config(num_slots=slots,num_iters=S,replay_buffer_size=max(1024,slots),nn_mode=mode)); ctx.init_weights(1337)
ocfg=common.oracle_config(ctx.cfg); blob=ctx.get_weights()
n=min(slots,64)
st,legal,tp=common.random_stacked(ocfg,n,seed=1)
vc,rv=ctx.run_mcts(st,legal,tp,True,np.arange(n,dtype=np.uint64),np.ones(n,np.int32))
same=sum(vc[i].tolist()==O.run_mcts(ocfg,blob,st[i],int(legal[i]),int(tp[i]),True,i,1)[0].tolist() for i in range(n))
print("run_mcts ok", same, "/", n, flush=True)
s,m=ctx.self_play(0,games,1.0)
print("self_play ok",s,m, flush=True)
ctx.close()
