set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0,'.')
import hydracore_b200 as hc
from hydracore_b200 import scene as S
from tests import refapi, scenes
t=time.time()
scn = S.Scene(1920,1080,S.Camera(pos=(0,7,9),look_at=(0,0,0),fov=45))
scn.add_instance(scn.add_mesh(S.grid_mesh(708,707))); scn.add_material(np.zeros(192,np.float32)); scn.build()
print('scene build s', time.time()-t, scn.bvh['nodes'].shape, scn.bvh['tris'].shape)
lay=hc.CudaLayer(); lay.LoadScene(scn)
rays=lay.MakeEyeRays(1920,1080,None)
for i in range(5):
    h=lay.TraceClosest(rays); print('primary ms', lay.last_trace_ms(), 'Mrays/s', rays.shape[0]/lay.last_trace_ms()/1e3)
hit=h['primId']>=0
# shadow rays to a point light
pos=rays[:,0:3]+rays[:,4:7]*h['t'][:,None]
L=np.array([0,20,0],np.float32)
sh=np.zeros_like(rays); d=L-pos; dist=np.linalg.norm(d,axis=1); sh[:,0:3]=pos+0.001*d/dist[:,None]; sh[:,4:7]=d/dist[:,None]; sh[:,7]=dist*0.995
sh=sh[hit]
for i in range(5):
    v=lay.TraceShadow(sh); print('shadow ms', lay.last_trace_ms(), 'Mrays/s', sh.shape[0]/lay.last_trace_ms()/1e3, v.mean())
# incoherent: cosine-ish random dirs from hit points
rng=np.random.RandomState(1)
n=hit.sum(); dd=rng.standard_normal((n,3)).astype(np.float32); dd/=np.linalg.norm(dd,axis=1,keepdims=True); dd[:,1]=np.abs(dd[:,1])
inc=np.zeros((n,8),np.float32); inc[:,0:3]=pos[hit]+0.001*np.array([0,1,0],np.float32); inc[:,4:7]=dd; inc[:,7]=3e38
for i in range(5):
    h2=lay.TraceClosest(inc); print('incoherent ms', lay.last_trace_ms(), 'Mrays/s', n/lay.last_trace_ms()/1e3, (h2['primId']>=0).mean())
orc=refapi.Oracle()
idx=np.arange(0,rays.shape[0],211)
t=time.time(); ho,cnt=orc.trace_closest(scn.bvh['nodes'],scn.bvh['tris'],rays[idx],count=True); print('oracle s',time.time()-t,'QLT/ray',cnt/len(idx))
idx2=np.arange(0,n,211)
ho,cnt=orc.trace_closest(scn.bvh['nodes'],scn.bvh['tris'],inc[idx2],count=True); print('incoherent QLT/ray',cnt/len(idx2))
PY
