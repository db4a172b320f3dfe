import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import solver as S, cnn as OC
import ml_accelerated_simulation_b200 as m
from ml_accelerated_simulation_b200 import models
import safetensors.torch as st
DEV='cuda:0'
def rel(a,b):
    a,b=a.double().cpu(),b.double().cpu(); return ((a-b).norm()/b.norm()).item()
# projection 2048x64
nx,ny,batch=2048,64,1
cfg=S.SolverConfig(nx=nx,ny=ny)
g=torch.Generator().manual_seed(nx+ny)
u64=torch.randn(batch,nx,ny,generator=g,dtype=torch.float64); v64=torch.randn(batch,nx,ny,generator=g,dtype=torch.float64)
plan=m.Plan(m.PlanConfig(nx=nx,ny=ny,batch=batch))
u,v=u64.float().to(DEV),v64.float().to(DEV)
p=plan.project(u,v)
ou,ov,op=S.pressure_projection(u64,v64,cfg); qu,qv,qp=S.pressure_projection(u64.float(),v64.float(),cfg)
print("proj u", rel(u,ou), rel(qu,ou), "v", rel(v,ov), rel(qv,ov), "p", rel(p,op), rel(qp,op))
d=plan.divergence(u,v); dq=S.divergence(qu,qv,cfg).abs().max().item()
vmax=max(u.abs().max().item(),v.abs().max().item())
print("div", d.abs().max().item(), dq, 1e-5*vmax/min(cfg.hx,cfg.hy))
u2,v2=u.clone(),v.clone(); plan.project(u2,v2,want_p=False)
print("idem", rel(u2,u), rel(v2,v), "pmean", abs(p.mean().item()), p.abs().max().item())
plan.close()
# rollout 64 100
gd=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'tests','golden')
cfgj=json.load(open(os.path.join(gd,'MLACFD.json'))); sd=st.load_file(os.path.join(gd,'mlacfd_weights.safetensors'))
model=models.buildModel(cfgj); model.load_state_dict(sd)
cfg=S.SolverConfig(nx=64,ny=64); u64,v64=S.filtered_velocity_field(cfg,1,seed=42)
plan=m.Plan(m.PlanConfig(nx=64,ny=64,batch=1)); model.attach(plan)
u,v=u64.float().to(DEV),v64.float().to(DEV)
ou,ov=u64.clone(),v64.clone(); qu,qv=u64.float(),v64.float(); bu,bv=u64.clone(),v64.clone()
done=0
for n in (1,10,30,100):
    plan.step(u,v,nsteps=n-done,apply_cnn=True)
    ou,ov,_=OC.rollout(ou,ov,cfg,n-done,cfgj,sd,1.0); qu,qv,_=OC.rollout(qu,qv,cfg,n-done,cfgj,sd,1.0)
    bu,bv,_=OC.rollout(bu,bv,cfg,n-done,cfgj,sd,1.0,round_bf16=True)
    done=n
    print(f"rollout n={n}: K vs O64 {rel(u,ou):.3e} | O32 vs O64 {rel(qu,ou):.3e} | O64(bf16emu cnn) vs O64 {rel(bu,ou):.3e} | K vs O(bf16emu) {rel(u,bu):.3e}")
