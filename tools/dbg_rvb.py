import sys; sys.path[:0]=['/root/repo','/root/repo/oracle','/root/repo/tests']
import numpy as np, torch, ctm_oracle as orc, helpers as H
from peps_torch_b200.engine import CtmEngine
eng=CtmEngine(); dev=torch.device('cuda:0')
z=np.load(H.GOLD+'/rvb_c4v_known_answer.npz'); a=torch.from_numpy(z['site']); chi=16
C,T=orc.init_env_c4v(a,chi); Cg,Tg=C.to(dev),T.to(dev); ag=a.to(dev)
for i in range(60):
    C,T,(M,D,U)=orc.ctm_move_c4v(a,C,T,chi,return_decomp=True)
    Cg,Tg,Dg=eng.move_c4v(ag,Cg,Tg,chi)
    if i%6==0 or i<4:
        e1=orc.energy_j1j2_c4v(a,C,T,1.0,0.5); e2=orc.energy_j1j2_c4v(a,Cg.cpu(),Tg.cpu(),1.0,0.5)
        print(i,'E cpu %.10f gpu %.10f'%(e1,e2),'D cpu',D[-4:].numpy(),'gpu',Dg.cpu()[-4:].numpy(), 'dC %.2e'%float((C-Cg.cpu()).abs().max()))
