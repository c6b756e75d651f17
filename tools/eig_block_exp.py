"""CTA eigensolver at the cfg4 shape (512 covariances of 64 x 64): block-pair scheme (option eig_onesided = 1, shipped) against the
column-pair scheme (2) and the two-sided kernel (0): time, projector and eigenvalue error against float64."""
import sys; sys.path.insert(0,'.')
import torch
from gr_doa_b200 import synth
import gr_doa_b200 as doa
M,N,T,B=64,16384,8,512
x,_=synth.frames_torch(B,M,N,[30.0+120.0*i/7 for i in range(8)],jitter_deg=2.0,device="cuda",chunk=32)
R=doa.autocorrelate(M,N,0,0,max_frames=B).work_device(x)
del x
mus=doa.MUSIC_lin_array(0.5,T,M,1024,max_frames=B)
def ev(fn,reps=5):
    fn();fn();torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record();torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
Rm=R.view(B,M,M).transpose(1,2).to(torch.complex128)
Rm=torch.triu(Rm)+torch.triu(Rm,1).conj().transpose(1,2)
w,V=torch.linalg.eigh(Rm); En=V[:,:,:M-T]; Gref=En@En.conj().transpose(1,2)
for opt in (1,2,0):
    mus.set_option("eig_onesided",opt)
    ms=ev(lambda: mus.noise_subspace_device(R))
    G,u,wv=mus.noise_subspace_device(R)
    e=(G.view(B,M,M).transpose(1,2).to(torch.complex128)-Gref).abs().amax(dim=(1,2))
    print(f"eig_onesided={opt}: {ms:.3f} ms, projector err mean {e.mean().item():.2e} max {e.max().item():.2e}, eigenvalue err {(wv.double()-w).abs().max().item():.2e}")
