import sys, ctypes as C, subprocess
sys.path.insert(0,'.')
import numpy as np
subprocess.run("nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -o /tmp/pose_dbg.so tools/dbg/pose_dbg.cu",shell=True,check=True)
import mjpl_b200 as mj
from mjpl_b200 import models,_abi
np.set_printoptions(precision=5,suppress=True,linewidth=220)
model=models.load('ur5e_scene'); q0=model.keyframe('home').qpos.copy(); site='attachment_site'
ref=mj.site_pose(model,q0,site); lim=(-0.1,0.1)
pc=mj.PoseConstraint(model,site,ref,z_translation=(-0.05,0.2),roll=lim,pitch=lim,q_step=0.5)
class Dump(C.Structure): _fields_=[('dx',C.c_double*6),('J',(C.c_double*8)*6),('A',(C.c_double*6)*6),('y',C.c_double*6),('site',C.c_double*7),('anchor',(C.c_double*3)*8),('axis',(C.c_double*3)*8)]
L=C.CDLL('/tmp/pose_dbg.so')
desc,keep=_abi.make_desc(model,[])
q=np.array([-1.559552,-1.839694,1.657243,-1.377122,-1.491698,0.292108])
h,d=Dump(),Dump(); sp=pc._spec()
print('rc',L.run(C.byref(desc),C.byref(sp),q.ctypes.data_as(C.c_void_p),C.byref(h),C.byref(d)))
for name in ('dx','site','y','anchor','axis','J','A'):
    a=np.array(getattr(h,name)); b=np.array(getattr(d,name)); print(name,'max diff',np.abs(a-b).max())
    if np.abs(a-b).max()>1e-9: print(' host\n',a,'\n dev\n',b)
