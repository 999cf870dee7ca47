"""Quick wall-clock timing of forward and forward + backward on the headline workload."""
import sys, time, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/3d-gaussian-splatting-for-novel-view-synthesis_b200')
import b200gs
from oracle import gs_oracle as O
for (n,W,H,ls) in [(1_000_000,1920,1080,-5.5),(3_000_000,1920,1080,-5.5)]:
    sc={k:v.cuda() for k,v in O.make_scene(n,seed=0,log_scale=ls).items()}
    cam=O.make_camera(W,H); c2w=cam['c2w'].cuda()
    with torch.no_grad():
        sigma=b200gs.build_sigma_from_params(sc['scale_raw'],sc['q_raw'])
        for it in range(3):
            col=b200gs.evaluate_sh(sc['f_dc'],sc['f_rest'],sc['pos'],c2w)
            img=b200gs.render(sc['pos'],col,sc['opacity_raw'],sigma,c2w,H,W,cam['fx'],cam['fy'],cam['cx'],cam['cy'])
        torch.cuda.synchronize(); t=time.time()
        for it in range(20):
            col=b200gs.evaluate_sh(sc['f_dc'],sc['f_rest'],sc['pos'],c2w)
            img=b200gs.render(sc['pos'],col,sc['opacity_raw'],sigma,c2w,H,W,cam['fx'],cam['fy'],cam['cx'],cam['cy'])
        torch.cuda.synchronize(); dt=(time.time()-t)/20
    print(n,W,H,'fwd ms',dt*1e3,'fps',1/dt, 'mean',img.mean().item())
    leaves={k:v.clone().requires_grad_(True) for k,v in sc.items()}
    for it in range(4):
        if it==1: torch.cuda.synchronize(); t=time.time()
        sigma=b200gs.build_sigma_from_params(leaves['scale_raw'],leaves['q_raw'])
        col=b200gs.evaluate_sh(leaves['f_dc'],leaves['f_rest'],leaves['pos'],c2w)
        img=b200gs.render(leaves['pos'],col,leaves['opacity_raw'],sigma,c2w,H,W,cam['fx'],cam['fy'],cam['cx'],cam['cy'])
        img.mean().backward()
    torch.cuda.synchronize(); dt=(time.time()-t)/3
    print(n,'fwd+bwd ms',dt*1e3)
