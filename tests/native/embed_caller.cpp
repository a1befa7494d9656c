// A stand-in for DSP-SLAM's LocalMapping thread (src/LocalMapping.cc:38-40, src/LocalMapping_util.cc:109-110,156-205):
// an embedded CPython interpreter (pybind11::embed), `import reconstruct.optimizer`, and the exact attribute calls /
// casts the C++ side performs, issued from a std::thread that takes the GIL -- against the one-file drop-in
// integration/reconstruct/optimizer.py.  Eigen is not in this image, so the arrays pybind11's Eigen casters would
// produce (float32, column-major strides, pybind11/eigen.h) are built explicitly as array_t<float, f_style>.
//
//   embed_caller <repo_root> <input.bin> <output.bin>
// input:  int32 M, N, Nfg | T[16] col-major | pts[M*3] col-major | rays[N*3] col-major | depth[Nfg] | scale | code[64]
// output: int32 is_good | T[16] row-major | code[64] | float loss | int32 nV, nF | pose-only T[16] | int32 flags
#include <pybind11/embed.h>
#include <pybind11/numpy.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace py = pybind11;
using farr = py::array_t<float, py::array::f_style>;

static std::vector<float> read_f(FILE* f, size_t n) {
  std::vector<float> v(n);
  if (n && fread(v.data(), 4, n, f) != n) { fprintf(stderr, "short input\n"); exit(2); }
  return v;
}

static farr col_major(const std::vector<float>& v, py::ssize_t rows, py::ssize_t cols) {
  farr a(std::vector<py::ssize_t>{rows, cols});           // strides (4, 4*rows): what Eigen::MatrixXf casts to
  std::memcpy(a.mutable_data(), v.data(), sizeof(float) * v.size());
  return a;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: embed_caller <repo_root> <input.bin> <output.bin>\n"); return 2; }
  const std::string root = argv[1];
  FILE* f = fopen(argv[2], "rb");
  if (!f) { perror("input"); return 2; }
  int32_t hdr[3];
  if (fread(hdr, 4, 3, f) != 3) return 2;
  const int M = hdr[0], N = hdr[1], Nfg = hdr[2];
  auto T = read_f(f, 16), pts = read_f(f, (size_t)M * 3), rays = read_f(f, (size_t)N * 3), depth = read_f(f, Nfg);
  auto scale = read_f(f, 1), code_in = read_f(f, 64);
  fclose(f);

  py::scoped_interpreter guard{};                          // src/System.cc:86-98 (Py_Initialize + imports)
  py::object pyOptimizer, pyMeshExtractor;
  {
    py::module_ sys = py::module_::import("sys");
    sys.attr("path").attr("insert")(0, root);              // dsp_slam_b200 on the path
    sys.attr("path").attr("insert")(0, root + "/integration");   // a DSP-SLAM checkout whose reconstruct/optimizer.py is the drop-in
    py::module_ optim = py::module_::import("reconstruct.optimizer");         // src/LocalMapping.cc:38
    py::module_ json = py::module_::import("json");
    py::object cfg = json.attr("load")(py::module_::import("builtins").attr("open")(root + "/dsp_slam_b200/configs/config_kitti.json"));
    py::object decoder = py::str(root + "/tests/golden/decoder_cars.npz");   // stands in for reconstruct.utils.get_decoder's module
    pyOptimizer = optim.attr("Optimizer")(decoder, cfg);                     // src/LocalMapping.cc:39
    pyMeshExtractor = optim.attr("MeshExtractor")(decoder, 64, 16);          // src/LocalMapping.cc:40
  }
  int rc = 0;
  std::vector<float> outT(16, 0.f), outCode(64, 0.f), outPose(16, 0.f);
  float loss = 0.f;
  int32_t is_good = 0, nV = 0, nF = 0, flags = 0;
  {
    py::gil_scoped_release release;                        // the main thread lets go of the GIL (src/System.cc:101)
    std::thread local_mapping([&]() {
      py::gil_scoped_acquire lock;                         // PyThreadStateLock (include/System.h:56-70)
      try {
        farr Sim3Tco = col_major(T, 4, 4), SurfacePoints = col_major(pts, M, 3), RayDirections = col_major(rays, N, 3);
        py::array_t<float> DepthObs(Nfg, depth.data());
        // src/LocalMapping_util.cc:179-192
        py::object pyMapObject = pyOptimizer.attr("reconstruct_object")(Sim3Tco, SurfacePoints, RayDirections, DepthObs);
        is_good = pyMapObject.attr("is_good").cast<bool>() ? 1 : 0;
        loss = pyMapObject.attr("loss").cast<float>();     // src/LocalMapping_util.cc:405
        if (is_good) {
          auto Tco = pyMapObject.attr("t_cam_obj").cast<py::array_t<float, py::array::c_style | py::array::forcecast>>();
          auto code = pyMapObject.attr("code").cast<py::array_t<float, py::array::c_style | py::array::forcecast>>();
          if (Tco.ndim() != 2 || Tco.shape(0) != 4 || Tco.shape(1) != 4 || code.size() != 64) { rc = 3; return; }
          std::memcpy(outT.data(), Tco.data(), 64);
          std::memcpy(outCode.data(), code.data(), 256);
          // src/LocalMapping_util.cc:194-196
          py::object pyMesh = pyMeshExtractor.attr("extract_mesh_from_code")(code);
          auto verts = pyMesh.attr("vertices").cast<py::array_t<float>>();
          auto faces = pyMesh.attr("faces").cast<py::array_t<int>>();
          nV = (int32_t)verts.shape(0); nF = (int32_t)faces.shape(0);
          if (verts.ndim() != 2 || verts.shape(1) != 3 || faces.ndim() != 2 || faces.shape(1) != 3) { rc = 4; return; }
        }
        if (pyOptimizer.attr("code_len").cast<int>() != 64) { rc = 5; return; }   // src/LocalMapping_util.cc:413
        // src/LocalMapping_util.cc:109-110: the return value is cast to Matrix4f unconditionally
        std::vector<float> se3 = T;
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) se3[c * 4 + r] /= scale[0];
        py::array_t<float> vcode(64, code_in.data());
        auto Tpo = pyOptimizer.attr("estimate_pose_cam_obj")(col_major(se3, 4, 4), scale[0], SurfacePoints, vcode)
                       .cast<py::array_t<float, py::array::c_style | py::array::forcecast>>();
        std::memcpy(outPose.data(), Tpo.data(), 64);
        // an unusable detection must come back as is_good=False, not as an exception (-> std::terminate here)
        farr empty(std::vector<py::ssize_t>{0, 3});
        py::object bad = pyOptimizer.attr("reconstruct_object")(Sim3Tco, empty, RayDirections, DepthObs);
        if (!bad.attr("is_good").cast<bool>() && bad.attr("t_cam_obj").is_none()) flags |= 1;
      } catch (py::error_already_set& e) {
        fprintf(stderr, "python exception reached C++: %s\n", e.what());
        rc = 10;
      }
    });
    local_mapping.join();
  }
  FILE* g = fopen(argv[3], "wb");
  fwrite(&is_good, 4, 1, g); fwrite(outT.data(), 4, 16, g); fwrite(outCode.data(), 4, 64, g); fwrite(&loss, 4, 1, g);
  fwrite(&nV, 4, 1, g); fwrite(&nF, 4, 1, g); fwrite(outPose.data(), 4, 16, g); fwrite(&flags, 4, 1, g);
  fclose(g);
  pyOptimizer = py::object(); pyMeshExtractor = py::object();
  return rc;
}
