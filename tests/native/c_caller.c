/* A plain-C caller of the C ABI (include/dspgn.h), the shape SURVEY 8(f3) asks for: all new detections of a keyframe
 * -- here the mono path's "normal + flipped pose, keep the lower loss" pair, src/LocalMapping_util.cc:390-407 -- go
 * into ONE dspgn_reconstruct_batch call, with the column-major (Eigen) strides the C++ side holds its matrices in
 * (row stride 1, column stride rows()).  No Python, no torch.
 *
 *   c_caller <weights.bin> <input.bin> <output.bin>
 * weights: int32 n_lin, latent, latent_in | per layer: int32 out, in | W[out*in] row-major | b[out]
 * input:   int32 M, N, Nfg | T[16] col-major | pts col-major | rays col-major | depth
 * output:  per object (2): int32 status | T[16] row-major | code[64] | loss
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dspgn.h"

static float* rd(FILE* f, size_t n) {
  float* p = (float*)malloc(4 * (n ? n : 1));
  if (n && fread(p, 4, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return p;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[3];
  if (fread(hdr, 4, 3, f) != 3) return 2;
  DspgnDecoderSpec spec;
  memset(&spec, 0, sizeof spec);
  spec.num_linear = hdr[0]; spec.latent_size = hdr[1]; spec.latent_in_layer = hdr[2];
  const float* W[DSPGN_MAX_LINEAR]; const float* B[DSPGN_MAX_LINEAR];
  for (int k = 0; k < spec.num_linear; ++k) {
    int d[2];
    if (fread(d, 4, 2, f) != 2) return 2;
    spec.out_dim[k] = d[0]; spec.in_dim[k] = d[1];
    W[k] = rd(f, (size_t)d[0] * d[1]); B[k] = rd(f, d[0]);
  }
  fclose(f);
  f = fopen(argv[2], "rb");
  if (!f || fread(hdr, 4, 3, f) != 3) return 2;
  const int M = hdr[0], N = hdr[1], Nfg = hdr[2];
  float* T = rd(f, 16); float* pts = rd(f, (size_t)M * 3); float* rays = rd(f, (size_t)N * 3); float* depth = rd(f, Nfg);
  fclose(f);

  DspgnDecoder* dec = NULL; DspgnSolver* sol = NULL;
  if (dspgn_decoder_create(&spec, W, B, 0, &dec)) { fprintf(stderr, "decoder: %s\n", dspgn_last_error()); return 3; }
  DspgnConfig cfg;
  memset(&cfg, 0, sizeof cfg);                       /* configs/config_kitti.json: optimizer block */
  cfg.k1 = 1.0f; cfg.k2 = 100.0f; cfg.k3 = 0.25f; cfg.k4 = 1e7f; cfg.b1 = 0.2f; cfg.b2 = 0.025f; cfg.lr = 1.0f; cfg.s_damp = 1.0f;
  cfg.num_iterations = 10; cfg.code_len = 64; cfg.num_depth_samples = 50; cfg.cut_off = 0.01f; cfg.pose_only_iterations = 5;
  cfg.sdf_only = 0; cfg.engine = DSPGN_ENGINE_AUTO;
  if (dspgn_solver_create(&cfg, &dec, 1, 0, &sol)) { fprintf(stderr, "solver: %s\n", dspgn_last_error()); return 3; }

  /* flipped pose: rotate the object 180 degrees about its y axis (x and z axes negated), LocalMapping_util.cc:394-401 */
  float Tf[16];
  memcpy(Tf, T, sizeof Tf);
  for (int r = 0; r < 4; ++r) { Tf[0 * 4 + r] = -T[0 * 4 + r]; Tf[2 * 4 + r] = -T[2 * 4 + r]; }   /* columns 0 and 2 (col-major) */
  DspgnObjectIn in[2];
  memset(in, 0, sizeof in);
  for (int i = 0; i < 2; ++i) {
    in[i].t_cam_obj = i ? Tf : T; in[i].t_rs = 1; in[i].t_cs = 4;
    in[i].pts = pts; in[i].n_pts = M; in[i].pts_rs = 1; in[i].pts_cs = M;
    in[i].rays = rays; in[i].n_rays = N; in[i].rays_rs = 1; in[i].rays_cs = N;
    in[i].depth = depth; in[i].n_depth = Nfg; in[i].code = NULL; in[i].scale = 1.f; in[i].class_id = 0;
  }
  DspgnObjectOut out[2];
  const int rc = dspgn_reconstruct_batch(sol, 2, in, out);      /* both candidates, all GN iterations, one call */
  if (rc) { fprintf(stderr, "reconstruct_batch: %s\n", dspgn_last_error()); return 4; }
  DspgnCounters c;
  dspgn_counters(sol, &c);
  f = fopen(argv[3], "wb");
  for (int i = 0; i < 2; ++i) {
    fwrite(&out[i].status, 4, 1, f); fwrite(out[i].t_cam_obj, 4, 16, f); fwrite(out[i].code, 4, 64, f); fwrite(&out[i].loss, 4, 1, f);
  }
  fclose(f);
  printf("c_caller: status %d/%d loss %.6f/%.6f kernel launches %lld\n", out[0].status, out[1].status, out[0].loss, out[1].loss,
         (long long)c.kernel_launches);
  dspgn_solver_destroy(sol);
  dspgn_decoder_destroy(dec);
  return 0;
}
