import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
from dsp_slam_b200 import synth, load_config
from dsp_slam_b200.optimizer import Optimizer
from oracle import dsp_oracle as O
G="/root/repo/tests/golden/"
cfg=load_config("config_kitti.json")
clss = ["cars", "chairs"] * 6
sizes = [700, 129, 2048, 64, 1000, 333, 128, 2047, 5, 900, 1500, 256]
objs = [synth.make_object(60 + i, m, cls=c) for i, (m, c) in enumerate(zip(sizes, clss))]
ins=[dict(t_cam_obj=o["t_cam_obj_init"], pts=o["pts"], class_id=(0 if c == "cars" else 1)) for o, c in zip(objs, clss)]
dws={c:O.DecoderWeights.from_npz(G+f"decoder_{c}.npz") for c in ("cars","chairs")}
ocfg=O.GNConfig.from_json_dict(cfg)
refs=[O.reconstruct_object(dws[c], ocfg, o["t_cam_obj_init"], o["pts"], None, None, sdf_only=True) for o,c in zip(objs,clss)]
def run(engine, mega):
    os.environ["DSPGN_MEGA"]="1" if mega else "0"
    opt=Optimizer(G+"decoder_cars.npz", cfg, sdf_only=True, engine=engine, extra_decoders=[G+"decoder_chairs.npz"])
    return opt.reconstruct_batch(ins)
res={"tc_mega":run("tc",True),"tc_iter":run("tc",False),"simt":run("simt",False)}
for i,(m,c) in enumerate(zip(sizes,clss)):
    line=f"{i:2d} {c:6s} M={m:5d} "
    for k,rs in res.items():
        line+=f"| {k} dT={np.abs(rs[i].t_cam_obj-refs[i]['t_cam_obj']).max():.2e} dc={np.abs(rs[i].code-refs[i]['code']).max():.2e} "
    print(line)
