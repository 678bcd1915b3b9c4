import sys, subprocess, os
code = '''
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import fem_helmholtz_eqn_b200 as fh
from oracle import banded1d
n=int(sys.argv[1]); rng=np.random.default_rng(0)
dl=rng.normal(size=(2,n))+1j*rng.normal(size=(2,n)); du=rng.normal(size=(2,n))+1j*rng.normal(size=(2,n))
d=4+rng.normal(size=(2,n))+1j*rng.normal(size=(2,n)); b=rng.normal(size=(2,n))+1j*rng.normal(size=(2,n))
u,st=fh.solve_tridiagonal(dl,d,du,b); torch.cuda.synchronize()
lo=dl[1].copy(); hi=du[1].copy(); lo[0]=0; hi[-1]=0
ref=banded1d.solve_bands(lo,d[1],hi,b[1])
print(n,"ok err",np.linalg.norm(u[1].cpu().numpy()-ref)/np.linalg.norm(ref), st.tolist())
'''
open('/tmp/probe_one.py','w').write(code)
for n in [2048, 16, 2056, 101, 8, 5, 4096, 4097]:
    r=subprocess.run([sys.executable,'/tmp/probe_one.py',str(n)],capture_output=True,text=True,env=dict(os.environ,CUDA_LAUNCH_BLOCKING='1'))
    print(n, (r.stdout.strip() or r.stderr.strip().splitlines()[-1][:200]))
