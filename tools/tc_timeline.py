import os, sys, json, ctypes as C
os.environ["DSPGN_CLK"]="1"
sys.path.insert(0, "/root/repo")
import numpy as np
from dsp_slam_b200 import synth, load_config, _lib
from dsp_slam_b200.optimizer import Optimizer
cfg=load_config("config_kitti.json")
objs=synth.make_batch(32,2048)
opt=Optimizer("/root/repo/tests/golden/decoder_cars.npz", cfg, sdf_only=True, engine="tc")
ins=[dict(t_cam_obj=o["t_cam_obj_init"], pts=o["pts"]) for o in objs]
opt.solver.upload(ins)
for _ in range(3): opt.solver.run(0); opt.solver.results_raw()
n=4*18*8
buf=(C.c_longlong*n)()
_lib.check(_lib.load().dspgn_debug_clocks(opt.solver.handle, buf, n))
a=np.array(buf[:]).reshape(4,18,8)
t0=a[0,0,4]
np.set_printoptions(linewidth=200)
print("tile step | mma:a_ready_seen  mma:issue_done | epi:acc_seen epi:math_done epi:st_done epi:arrived   (cycles rel. to first a_ready)")
for t in range(4):
    for s in range(17):
        r=a[t,s]-t0
        print(t, s, "|", r[4], r[5], "|", r[0], r[1], r[2], r[3], "| mma_issue", r[5]-r[4], "mma_run", r[0]-r[4], "epi", r[3]-r[0])
n2=4*18*8+16
buf2=(C.c_longlong*n2)()
_lib.check(_lib.load().dspgn_debug_clocks(opt.solver.handle, buf2, n2))
sv=np.array(buf2[4*18*8:])
print("k_solve stamps (cycles from start): start, loss-reductions, rot-prior+sync, tile-partial loads, As fill, elimination+backsub, end")
print((sv[:7]-sv[0]).tolist())
# effective SM clock during the kernel: clock64 vs %globaltimer (ns) between the first steps of tile 0 and tile 3
dc = a[3,0,4]-a[0,0,4]; dg = a[3,0,6]-a[0,0,6]
print("effective SM clock over tiles 0..3: %.3f GHz  (%d cycles in %d ns)" % (dc/max(dg,1), dc, dg))
