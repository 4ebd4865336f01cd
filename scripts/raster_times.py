import sys, json
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from legitengine_b200 import abi, harness, scene
W,H=3840,2160
mesh=scene.scene_mesh(0xC0FFEE)
r=harness.Renderer(W,H)
r.upload_mesh(mesh)
for _ in range(3):
    r.render_frame(harness.MODE_FUSED,0,abi.GI_DEFAULT)
r.sync()
acc={}
for _ in range(5):
    r.render_frame(harness.MODE_FUSED,0,abi.GI_DEFAULT,profile=True); r.sync()
    for n,ms in r.profile(): acc[n]=acc.get(n,0)+ms/5
print(json.dumps({k:round(v,4) for k,v in acc.items()}))
