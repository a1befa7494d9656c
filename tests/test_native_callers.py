"""Row f3 of SURVEY 8: compiled native callers of the boundary.

* tests/native/embed_caller.cpp -- a stand-in for DSP-SLAM's LocalMapping thread: pybind11::embed interpreter,
  `import reconstruct.optimizer` (the one-file drop-in under integration/), column-major float32 arrays as pybind11's
  Eigen casters deliver them, the attribute reads / casts of src/LocalMapping_util.cc:109-110,179-196, issued from a
  std::thread holding the GIL.
* tests/native/c_caller.c -- plain C against include/dspgn.h: the mono path's normal + flipped candidate pair
  (src/LocalMapping_util.cc:390-407) as ONE dspgn_reconstruct_batch call with Eigen strides.

CPU: both must compile and link.  GPU: both must reproduce the Python path bit for bit.
"""
import os
import struct
import subprocess
import sys
import sysconfig

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
NATIVE = os.path.join(ROOT, "tests", "native")


def _build(tmp):
    import pybind11
    inc = [f"-I{sysconfig.get_paths()['include']}", f"-I{pybind11.get_include()}"]
    libdir = sysconfig.get_config_var("LIBDIR") or "/usr/lib/x86_64-linux-gnu"
    cfgdir = sysconfig.get_config_var("LIBPL") or libdir
    ver = sysconfig.get_config_var("LDVERSION") or "3.12"
    emb = os.path.join(tmp, "embed_caller")
    subprocess.check_call(["g++", "-O1", "-std=c++17", *inc, os.path.join(NATIVE, "embed_caller.cpp"), "-o", emb,
                           f"-L{cfgdir}", f"-L{libdir}", f"-lpython{ver}", "-ldl", "-lm", f"-Wl,-rpath,{libdir}"])
    cc = os.path.join(tmp, "c_caller")
    libd = os.path.join(ROOT, "dsp_slam_b200")
    subprocess.check_call(["gcc", "-O1", "-std=c11", f"-I{os.path.join(ROOT, 'include')}", os.path.join(NATIVE, "c_caller.c"),
                           "-o", cc, f"-L{libd}", "-ldspgn", f"-Wl,-rpath,{libd}"])
    return emb, cc


def test_native_callers_compile_and_link(tmp_path):
    emb, cc = _build(str(tmp_path))
    assert os.path.isfile(emb) and os.path.isfile(cc)
    # the C caller resolves every dspgn_* symbol it uses at link time; without arguments it exits with usage code 2
    assert subprocess.run([cc]).returncode == 2


def _write_inputs(path, d, scale, code):
    T = np.asfortranarray(d["in_t_cam_obj"], dtype=np.float32)
    P = np.asfortranarray(d["in_pts"], dtype=np.float32)
    R = np.asfortranarray(d["in_rays"], dtype=np.float32)
    dep = np.ascontiguousarray(d["in_depth"], dtype=np.float32)
    with open(path, "wb") as f:
        f.write(struct.pack("<3i", P.shape[0], R.shape[0], dep.shape[0]))
        for a in (T, P, R):
            f.write(a.tobytes(order="F"))                    # column-major element order
        f.write(dep.tobytes())
        f.write(struct.pack("<f", scale))
        f.write(np.asarray(code, np.float32).tobytes())


@pytest.mark.gpu
def test_embedded_interpreter_caller_matches_python(tmp_path, golden_dir, cfg_kitti):
    from dsp_slam_b200.optimizer import Optimizer, MeshExtractor
    emb, _ = _build(str(tmp_path))
    d = np.load(os.path.join(golden_dir, "recon_kitti250.npz"))
    T = np.array(d["in_t_cam_obj"], dtype=np.float32)
    scale = float(np.cbrt(np.linalg.det(T[:3, :3].astype(np.float64))))
    code = (0.5 * d["gt_code"]).astype(np.float32)
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_inputs(inp, d, scale, code)
    env = dict(os.environ)
    site = [p for p in sys.path if p.endswith("site-packages")]
    env["PYTHONPATH"] = os.pathsep.join(site + [env.get("PYTHONPATH", "")])
    r = subprocess.run([emb, ROOT, inp, outp], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    raw = open(outp, "rb").read()
    is_good, = struct.unpack_from("<i", raw, 0)
    Tn = np.frombuffer(raw, np.float32, 16, 4).reshape(4, 4)
    zn = np.frombuffer(raw, np.float32, 64, 4 + 64)
    loss, nV, nF = struct.unpack_from("<fii", raw, 4 + 64 + 256)
    Tpo = np.frombuffer(raw, np.float32, 16, 4 + 64 + 256 + 12).reshape(4, 4)
    flags, = struct.unpack_from("<i", raw, 4 + 64 + 256 + 12 + 64)
    dec = os.path.join(golden_dir, "decoder_cars.npz")
    opt = Optimizer(dec, cfg_kitti)
    ref = opt.reconstruct_object(np.asfortranarray(d["in_t_cam_obj"]), np.asfortranarray(d["in_pts"]),
                                 np.asfortranarray(d["in_rays"]), d["in_depth"])
    assert is_good == 1 and ref.is_good
    np.testing.assert_array_equal(Tn, ref.t_cam_obj)
    np.testing.assert_array_equal(zn, ref.code)
    assert loss == np.float32(ref.loss)
    mesh = MeshExtractor(dec, 64, 16).extract_mesh_from_code(ref.code)
    assert (nV, nF) == (mesh.vertices.shape[0], mesh.faces.shape[0]) and nF > 50
    se3 = T.copy(); se3[:3, :3] /= np.float32(scale)
    np.testing.assert_array_equal(Tpo, opt.estimate_pose_cam_obj(np.asfortranarray(se3), np.float32(scale), np.asfortranarray(d["in_pts"]), code))
    assert flags & 1                                          # unusable detection came back as a soft failure


@pytest.mark.gpu
def test_plain_c_caller_batches_the_mono_candidate_pair(tmp_path, golden_dir, cfg_kitti):
    from dsp_slam_b200.optimizer import Optimizer
    from dsp_slam_b200.decoder import DecoderWeights
    _, cc = _build(str(tmp_path))
    d = np.load(os.path.join(golden_dir, "recon_kitti250.npz"))
    dec = os.path.join(golden_dir, "decoder_cars.npz")
    w = DecoderWeights.from_npz(dec)
    wp, inp, outp = str(tmp_path / "w.bin"), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(wp, "wb") as f:
        f.write(struct.pack("<3i", len(w.W), w.latent_size, w.latent_in_layer))
        for W, b in zip(w.W, w.b):
            f.write(struct.pack("<2i", *W.shape)); f.write(W.tobytes()); f.write(b.tobytes())
    _write_inputs(inp, d, 1.0, np.zeros(64, np.float32))
    r = subprocess.run([cc, wp, inp, outp], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "kernel launches" in r.stdout
    raw = np.frombuffer(open(outp, "rb").read(), np.float32).reshape(2, 82)
    opt = Optimizer(dec, cfg_kitti)
    T = np.array(d["in_t_cam_obj"], dtype=np.float32)
    Tf = T.copy(); Tf[:, 0] *= -1; Tf[:, 2] *= -1                # LocalMapping_util.cc:394-401: flipped candidate
    o = dict(pts=d["in_pts"], rays=d["in_rays"], depth=d["in_depth"])
    ref = opt.reconstruct_batch([dict(o, t_cam_obj=T), dict(o, t_cam_obj=Tf)])
    for i in range(2):
        st = int(raw[i].view(np.int32)[0])
        assert (st == 0) == ref[i].is_good
        if ref[i].is_good:
            np.testing.assert_array_equal(raw[i, 1:17].reshape(4, 4), ref[i].t_cam_obj)
            np.testing.assert_array_equal(raw[i, 17:81], ref[i].code)
        assert raw[i, 81] == np.float32(ref[i].loss)
    assert ref[0].is_good
