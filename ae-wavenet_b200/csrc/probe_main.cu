// Stand-alone probe for the tcgen05 engines (tgemm / wgrad): checks them against a CPU loop on small problems with
// TF32-exact (small integer) data, then times reference-sized problems.  Development tool, not part of libaewn.so.
//   build: see Makefile target `probe`;   run (GPU box): ./build/aewn_probe [quick]
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/aewn.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static uint32_t rng_state = 12345u;
static uint32_t rnd() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 8;
}
static float rnd_int(int lim) { return (float)((int)(rnd() % (2 * lim + 1)) - lim); }
static float rnd_f() { return (float)(rnd() & 0xFFFF) / 32768.0f - 1.0f; }

struct Act {
  int B, C, T, pitch;
  std::vector<float> h;
  float* d = nullptr;
  Act(int B_, int C_, int T_, bool ints, float fill_pad = 0.f) : B(B_), C(C_), T(T_) {
    pitch = (T + 31) / 32 * 32;
    h.assign((size_t)B * C * pitch, fill_pad);
    for (int b = 0; b < B; ++b)
      for (int c = 0; c < C; ++c)
        for (int t = 0; t < T; ++t) h[((size_t)b * C + c) * pitch + t] = ints ? rnd_int(3) : rnd_f();
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  }
  float at(int b, int c, int t) const {
    if (t < 0 || t >= T || c < 0 || c >= C) return 0.f;
    return h[((size_t)b * C + c) * pitch + t];
  }
  aewn_act desc() const {
    aewn_act a;
    a.ptr = d;
    a.t_extent = T;
    a.channels = C;
    a.batch = B;
    a.row_pitch = pitch;
    a.batch_stride = (long long)C * pitch;
    return a;
  }
  void download() { CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost)); }
  void upload() { CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice)); }
};

static int* g_err = nullptr;
static int read_err() {
  int e = 0;
  CK(cudaMemcpy(&e, g_err, 4, cudaMemcpyDeviceToHost));
  int z = 0;
  CK(cudaMemcpy(g_err, &z, 4, cudaMemcpyHostToDevice));
  return e;
}

static aewn_ntile mk_tile(int w_row, int n, int n_valid, int mode, int flags, int seg_mask, int t_lo, int t_hi,
                          float* out, long long bs, long long cs) {
  aewn_ntile nt;
  memset(&nt, 0, sizeof(nt));
  nt.w_row = w_row;
  nt.n = n;
  nt.n_valid = n_valid;
  nt.mode = mode;
  nt.flags = flags;
  nt.seg_mask = seg_mask;
  nt.t_lo = t_lo;
  nt.t_hi = t_hi;
  nt.out = out;
  nt.out_bs = bs;
  nt.out_cs = cs;
  return nt;
}

// ---------------------------------------------------------------- test 1: single segment, descriptor candidates
static int g_shift = 0;
static int test_basic(int lbo, int sbo, bool ints, double tol) {
  const int B = 2, C = 40, T = 300, N = 48, KP = 64;
  Act x(B, C, T, ints);
  std::vector<float> w((size_t)N * KP, 0.f);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < C; ++k) w[(size_t)n * KP + k] = ints ? rnd_int(2) : rnd_f();
  float* dw;
  CK(cudaMalloc(&dw, w.size() * 4));
  CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  Act out(B, N, T, true);
  CK(cudaMemset(out.d, 0, out.h.size() * 4));

  aewn_tgemm_desc d;
  memset(&d, 0, sizeof(d));
  d.acts[0] = x.desc();
  d.n_acts = 1;
  d.segs[0] = {0, g_shift, C, 0};
  d.n_segs = 1;
  d.w = dw;
  d.w_rows = N;
  d.w_kpad = KP;
  d.ntiles[0] = mk_tile(0, N, N, AEWN_EPI_LINEAR, 0, 1, 0, T, out.d, (long long)N * out.pitch, out.pitch);
  d.n_ntiles = 1;
  d.batch = B;
  d.t_begin = 0;
  d.t_end = T;
  d.err = g_err;
  d.dbg_lbo = lbo;
  d.dbg_sbo = sbo;
  int rc = aewn_tgemm(&d, 0);
  if (rc) {
    printf("  tgemm rc=%d (%s)\n", rc, aewn_last_error_string());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  sync error: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  int derr = read_err();
  out.download();
  double maxerr = 0;
  int bad = 0;
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < N; ++n)
      for (int t = 0; t < T; ++t) {
        double ref = 0;
        for (int k = 0; k < C; ++k) ref += (double)x.at(b, k, t + g_shift) * w[(size_t)n * KP + k];
        double err = fabs(ref - out.at(b, n, t));
        if (err > maxerr) maxerr = err;
        if (err > tol && bad < 4) {
          printf("    mismatch b=%d n=%d t=%d ref=%g got=%g\n", b, n, t, ref, out.at(b, n, t));
          ++bad;
        }
      }
  printf("  basic lbo=%d sbo=%d ints=%d: max|err|=%g device_err=%d -> %s\n", lbo, sbo, (int)ints, maxerr, derr,
         (maxerr <= tol && derr == 0) ? "PASS" : "FAIL");
  cudaFree(dw);
  cudaFree(x.d);
  cudaFree(out.d);
  return (maxerr <= tol && derr == 0) ? 0 : 1;
}

// ---------------------------------------------------------------- test 2: layer-shaped problem (3 segs, 2+ n-tiles)
static int test_layer(bool ints, double tol, int variant = 0) {
  const int B = 2, R = 72, Cc = 11, T = 700, d_ = 8, N = 368;
  const int KR = (R + 31) / 32 * 32, KC = (Cc + 31) / 32 * 32, KP = 2 * KR + KC;
  Act x(B, R, T, ints, NAN), cond(B, Cc, T, ints, NAN);
  std::vector<float> w((size_t)N * KP, 0.f);
  for (int n = 0; n < N; ++n) {
    for (int k = 0; k < R; ++k) {
      w[(size_t)n * KP + k] = ints ? rnd_int(2) : rnd_f();
      w[(size_t)n * KP + KR + k] = ints ? rnd_int(2) : rnd_f();
    }
    for (int k = 0; k < Cc; ++k) w[(size_t)n * KP + 2 * KR + k] = ints ? rnd_int(2) : rnd_f();
  }
  float* dw;
  CK(cudaMalloc(&dw, w.size() * 4));
  CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  Act add(B, N, T, ints), out(B, N, T, ints);  // out pre-filled: tile 1 accumulates into it
  std::vector<float> out0 = out.h;
  std::vector<float> bias(N);
  for (auto& v : bias) v = ints ? rnd_int(2) : rnd_f();
  float* dbias;
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMemcpy(dbias, bias.data(), N * 4, cudaMemcpyHostToDevice));

  const int t_lo = 37, t_hi = T;
  aewn_tgemm_desc d;
  memset(&d, 0, sizeof(d));
  d.acts[0] = x.desc();
  d.acts[1] = cond.desc();
  d.n_acts = 2;
  d.segs[0] = {0, -d_, R, 0};
  d.segs[1] = {0, 0, R, KR};
  d.segs[2] = {1, 0, Cc, 2 * KR};
  d.n_segs = 3;
  d.w = dw;
  d.w_rows = N;
  d.w_kpad = KP;
  // tile 0: rows [0,256): out = acc + bias + add, relu.  tile 1: rows [256,368): out += acc, only segs 0 and 2, t >= 200
  d.ntiles[0] = mk_tile(0, 256, 256, AEWN_EPI_LINEAR, AEWN_F_RELU, 7, t_lo, t_hi, out.d, (long long)N * out.pitch,
                        out.pitch);
  d.ntiles[0].add = add.d;
  d.ntiles[0].add_bs = (long long)N * add.pitch;
  d.ntiles[0].add_cs = add.pitch;
  d.ntiles[0].bias = dbias;
  d.ntiles[1] = mk_tile(256, 112, 112, AEWN_EPI_LINEAR, AEWN_F_ACCUM, 5, 200, t_hi, out.d + (size_t)256 * out.pitch,
                        (long long)N * out.pitch, out.pitch);
  d.n_ntiles = 2;
  if (variant == 1) { d.n_ntiles = 1; }                       // only the 256-wide tile
  if (variant == 2) { d.ntiles[0] = d.ntiles[1]; d.n_ntiles = 1; }  // only the 112-wide tile
  if (variant == 3) { d.n_segs = 1; d.ntiles[0].seg_mask = 1; d.ntiles[1].seg_mask = 1; }
  const bool tma = variant == 4;   // TMA-store epilogue: rows of the first active tile below t_lo receive zeros
  d.no_tma_store = tma ? 0 : 1;
  d.batch = B;
  d.t_begin = 32;
  d.t_end = T;
  d.err = g_err;
  int rc = aewn_tgemm(&d, 0);
  if (rc) {
    printf("  tgemm rc=%d (%s)\n", rc, aewn_last_error_string());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  sync error: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  int derr = read_err();
  out.download();
  double maxerr = 0;
  int bad = 0;
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < N; ++n)
      for (int t = 0; t < T; ++t) {
        double ref;
        const float prev = out0[((size_t)b * N + n) * out.pitch + t];
        if (n < 256) {
          if (t < t_lo) ref = (tma && t >= 32) ? 0.0 : prev;
          else {
            double acc = bias[n] + add.at(b, n, t);
            for (int k = 0; k < R; ++k)
              acc += (double)x.at(b, k, t - d_) * w[(size_t)n * KP + k] + (double)x.at(b, k, t) * w[(size_t)n * KP + KR + k];
            for (int k = 0; k < Cc; ++k) acc += (double)cond.at(b, k, t) * w[(size_t)n * KP + 2 * KR + k];
            ref = acc > 0 ? acc : 0;
          }
        } else {
          if (t < 200) ref = prev;
          else {
            double acc = prev;
            for (int k = 0; k < R; ++k) acc += (double)x.at(b, k, t - d_) * w[(size_t)n * KP + k];
            for (int k = 0; k < Cc; ++k) acc += (double)cond.at(b, k, t) * w[(size_t)n * KP + 2 * KR + k];
            ref = acc;
          }
        }
        double err = fabs(ref - out.at(b, n, t));
        if (!(err <= maxerr)) maxerr = err;
        if (!(err <= tol) && bad < 6) {
          printf("    mismatch b=%d n=%d t=%d ref=%g got=%g\n", b, n, t, ref, out.at(b, n, t));
          ++bad;
        }
      }
  printf("  layer ints=%d variant=%d: max|err|=%g device_err=%d -> %s\n", (int)ints, variant, maxerr, derr,
         (maxerr <= tol && derr == 0) ? "PASS" : "FAIL");
  return (maxerr <= tol && derr == 0) ? 0 : 1;
}

// ---------------------------------------------------------------- test 3: gate epilogues
static int test_gate() {
  const int B = 1, R = 32, T = 260, N = 256;
  Act x(B, R, T, false);
  std::vector<float> w((size_t)N * 32, 0.f);
  for (auto& v : w) v = rnd_f() * 0.3f;
  float* dw;
  CK(cudaMalloc(&dw, w.size() * 4));
  CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  std::vector<float> bias(N);
  for (auto& v : bias) v = rnd_f();
  float* dbias;
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMemcpy(dbias, bias.data(), N * 4, cudaMemcpyHostToDevice));
  Act th(B, 128, T, true), sg(B, 128, T, true), z(B, 128, T, true);
  aewn_tgemm_desc d;
  memset(&d, 0, sizeof(d));
  d.acts[0] = x.desc();
  d.n_acts = 1;
  d.segs[0] = {0, 0, R, 0};
  d.n_segs = 1;
  d.w = dw;
  d.w_rows = N;
  d.w_kpad = 32;
  d.ntiles[0] = mk_tile(0, 256, 128, AEWN_EPI_GATE_FWD, 0, 1, 0, T, th.d, (long long)128 * th.pitch, th.pitch);
  d.ntiles[0].out2 = sg.d;
  d.ntiles[0].out3 = z.d;
  d.ntiles[0].bias = dbias;
  d.n_ntiles = 1;
  d.batch = B;
  d.t_begin = 0;
  d.t_end = T;
  d.err = g_err;
  int rc = aewn_tgemm(&d, 0);
  if (rc) {
    printf("  tgemm(gate fwd) rc=%d (%s)\n", rc, aewn_last_error_string());
    return 1;
  }
  CK(cudaDeviceSynchronize());
  int derr = read_err();
  th.download();
  sg.download();
  z.download();
  double e1 = 0, e2 = 0, e3 = 0;
  for (int c = 0; c < 128; ++c)
    for (int t = 0; t < T; ++t) {
      double f = bias[c], g = bias[128 + c];
      for (int k = 0; k < R; ++k) {
        f += (double)x.at(0, k, t) * w[(size_t)c * 32 + k];
        g += (double)x.at(0, k, t) * w[(size_t)(128 + c) * 32 + k];
      }
      double rt = tanh(f), rs = 1.0 / (1.0 + exp(-g));
      e1 = fmax(e1, fabs(rt - th.at(0, c, t)));
      e2 = fmax(e2, fabs(rs - sg.at(0, c, t)));
      e3 = fmax(e3, fabs(rt * rs - z.at(0, c, t)));
    }
  printf("  gate_fwd: max|err| tanh=%g sigmoid=%g z=%g device_err=%d -> %s\n", e1, e2, e3, derr,
         (e3 < 5e-3 && derr == 0) ? "PASS" : "FAIL");
  int fails = (e3 < 5e-3 && derr == 0) ? 0 : 1;

  // gate bwd: acc = g_z from a GEMM over gin (R channels) -> 256 columns; th/sg = the tensors above (128 ch) tiled x2
  Act th2(B, 256, T, false), sg2(B, 256, T, false), gf(B, 256, T, true), gg(B, 256, T, true);
  memset(&d.ntiles[0], 0, sizeof(aewn_ntile));
  d.ntiles[0] = mk_tile(0, 256, 256, AEWN_EPI_GATE_BWD, 0, 1, 0, T, gf.d, (long long)256 * gf.pitch, gf.pitch);
  d.ntiles[0].out2 = gg.d;
  d.ntiles[0].add = th2.d;
  d.ntiles[0].add2 = sg2.d;
  d.ntiles[0].add_bs = (long long)256 * th2.pitch;
  d.ntiles[0].add_cs = th2.pitch;
  d.ntiles[0].t_zero_lo = 50;
  rc = aewn_tgemm(&d, 0);
  if (rc) {
    printf("  tgemm(gate bwd) rc=%d (%s)\n", rc, aewn_last_error_string());
    return 1;
  }
  CK(cudaDeviceSynchronize());
  derr = read_err();
  gf.download();
  gg.download();
  double e4 = 0, e5 = 0;
  for (int c = 0; c < 256; ++c)
    for (int t = 0; t < T; ++t) {
      double gz = 0;
      for (int k = 0; k < R; ++k) gz += (double)x.at(0, k, t) * w[(size_t)c * 32 + k];
      double tt = th2.at(0, c, t), ss = sg2.at(0, c, t);
      double rf = t < 50 ? 0 : gz * ss * (1 - tt * tt), rg = t < 50 ? 0 : gz * tt * ss * (1 - ss);
      e4 = fmax(e4, fabs(rf - gf.at(0, c, t)));
      e5 = fmax(e5, fabs(rg - gg.at(0, c, t)));
    }
  printf("  gate_bwd: max|err| g_f=%g g_g=%g device_err=%d -> %s\n", e4, e5, derr,
         (e4 < 5e-3 && e5 < 5e-3 && derr == 0) ? "PASS" : "FAIL");
  fails += (e4 < 5e-3 && e5 < 5e-3 && derr == 0) ? 0 : 1;
  return fails;
}

// ---------------------------------------------------------------- test 4: wgrad
static int test_wgrad(bool ints, double tol) {
  const int B = 2, M = 130, N = 70, T = 500, shift = -8, t_lo = 12;
  Act g(B, M, T, ints, NAN), x(B, N, T, ints, NAN);
  std::vector<float> out0((size_t)M * N * 2, 0.f);  // strided output: rs = 2N, cs = 2 (conv weight layout, tap 1)
  for (auto& v : out0) v = ints ? rnd_int(2) : rnd_f();
  float* dout;
  CK(cudaMalloc(&dout, out0.size() * 4));
  CK(cudaMemcpy(dout, out0.data(), out0.size() * 4, cudaMemcpyHostToDevice));
  aewn_wgrad_desc d;
  memset(&d, 0, sizeof(d));
  d.acts[0] = g.desc();
  d.acts[1] = x.desc();
  d.n_acts = 2;
  for (int mt = 0; mt < 2; ++mt) {
    aewn_wgrad_item& im = d.items[mt];
    im.g_act = 0;
    im.x_act = 1;
    im.g_row = mt * 128;
    im.x_row = 0;
    im.m_valid = mt == 0 ? 128 : M - 128;
    im.n = 80;
    im.n_valid = N;
    im.shift = shift;
    im.t_lo = t_lo;
    im.t_hi = T;
    im.n_split = 3;
    im.out = dout + (size_t)mt * 128 * 2 * N + 1;
    im.out_rs = 2 * N;
    im.out_cs = 2;
  }
  d.n_items = 2;
  d.batch = B;
  d.err = g_err;
  int rc = aewn_wgrad(&d, 0);
  if (rc) {
    printf("  wgrad rc=%d (%s)\n", rc, aewn_last_error_string());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  sync error: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  int derr = read_err();
  std::vector<float> out(out0.size());
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n)
      for (int tap = 0; tap < 2; ++tap) {
        double ref = out0[((size_t)m * N + n) * 2 + tap];
        if (tap == 1)
          for (int b = 0; b < B; ++b)
            for (int u = t_lo; u < T; ++u) ref += (double)g.at(b, m, u) * x.at(b, n, u + shift);
        double err = fabs(ref - out[((size_t)m * N + n) * 2 + tap]);
        if (!(err <= maxerr)) maxerr = err;
        if (!(err <= tol) && bad < 6) {
          printf("    mismatch m=%d n=%d tap=%d ref=%g got=%g\n", m, n, tap, ref, out[((size_t)m * N + n) * 2 + tap]);
          ++bad;
        }
      }
  printf("  wgrad ints=%d: max|err|=%g device_err=%d -> %s\n", (int)ints, maxerr, derr,
         (maxerr <= tol && derr == 0) ? "PASS" : "FAIL");
  return (maxerr <= tol && derr == 0) ? 0 : 1;
}

// ---------------------------------------------------------------- timing at reference size (cfg2 layer 0)
static void time_layer() {
  const int B = 8, R = 368, Cc = 139, D = 256, T = 18430, d_ = 4;
  const int pitch = (T + 31) / 32 * 32;
  const int KR = 384, KC = 160, KP = 2 * KR + KC;
  float *x, *cond, *w1, *th, *sg, *z, *w2, *sig, *skp;
  CK(cudaMalloc(&x, (size_t)B * R * pitch * 4));
  CK(cudaMalloc(&cond, (size_t)B * Cc * pitch * 4));
  CK(cudaMalloc(&w1, (size_t)512 * KP * 4));
  CK(cudaMalloc(&w2, (size_t)624 * 256 * 4));
  CK(cudaMalloc(&th, (size_t)B * D * pitch * 4));
  CK(cudaMalloc(&sg, (size_t)B * D * pitch * 4));
  CK(cudaMalloc(&z, (size_t)B * D * pitch * 4));
  CK(cudaMalloc(&sig, (size_t)B * R * pitch * 4));
  CK(cudaMalloc(&skp, (size_t)B * D * pitch * 4));
  CK(cudaMemset(x, 0, (size_t)B * R * pitch * 4));
  CK(cudaMemset(cond, 0, (size_t)B * Cc * pitch * 4));
  CK(cudaMemset(w1, 0, (size_t)512 * KP * 4));
  CK(cudaMemset(w2, 0, (size_t)624 * 256 * 4));
  CK(cudaMemset(skp, 0, (size_t)B * D * pitch * 4));
  CK(cudaMemset(z, 0, (size_t)B * D * pitch * 4));

  aewn_tgemm_desc g1;
  memset(&g1, 0, sizeof(g1));
  g1.acts[0] = {x, T, R, B, pitch, (long long)R * pitch};
  g1.acts[1] = {cond, T, Cc, B, pitch, (long long)Cc * pitch};
  g1.n_acts = 2;
  g1.segs[0] = {0, -d_, R, 0};
  g1.segs[1] = {0, 0, R, KR};
  g1.segs[2] = {1, 0, Cc, 2 * KR};
  g1.n_segs = 3;
  g1.w = w1;
  g1.w_rows = 512;
  g1.w_kpad = KP;
  for (int j = 0; j < 2; ++j) {
    g1.ntiles[j] = mk_tile(j * 256, 256, 128, AEWN_EPI_GATE_FWD, 0, 7, d_, T, th + (size_t)j * 128 * pitch,
                           (long long)D * pitch, pitch);
    g1.ntiles[j].out2 = sg + (size_t)j * 128 * pitch;
    g1.ntiles[j].out3 = z + (size_t)j * 128 * pitch;
  }
  g1.n_ntiles = 2;
  g1.batch = B;
  g1.t_begin = 0;
  g1.t_end = T;
  g1.err = g_err;

  aewn_tgemm_desc g2;
  memset(&g2, 0, sizeof(g2));
  g2.acts[0] = {z, T, D, B, pitch, (long long)D * pitch};
  g2.n_acts = 1;
  g2.segs[0] = {0, 0, D, 0};
  g2.n_segs = 1;
  g2.w = w2;
  g2.w_rows = 624;
  g2.w_kpad = 256;
  g2.ntiles[0] = mk_tile(0, 256, 256, AEWN_EPI_LINEAR, 0, 1, d_, T, sig, (long long)R * pitch, pitch);
  g2.ntiles[0].add = x;
  g2.ntiles[0].add_bs = (long long)R * pitch;
  g2.ntiles[0].add_cs = pitch;
  g2.ntiles[1] = mk_tile(256, 112, 112, AEWN_EPI_LINEAR, 0, 1, d_, T, sig + (size_t)256 * pitch, (long long)R * pitch, pitch);
  g2.ntiles[1].add = x + (size_t)256 * pitch;
  g2.ntiles[1].add_bs = (long long)R * pitch;
  g2.ntiles[1].add_cs = pitch;
  g2.ntiles[2] = mk_tile(368, 256, 256, AEWN_EPI_LINEAR, AEWN_F_ACCUM, 1, 2046, T, skp, (long long)D * pitch, pitch);
  g2.n_ntiles = 3;
  g2.batch = B;
  g2.t_begin = 0;
  g2.t_end = T;
  g2.err = g_err;

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int which = 0; which < 6; ++which) {
    aewn_tgemm_desc* g = (which & 1) == 0 ? &g1 : &g2;
    g->cluster = which < 2 ? 1 : (which < 4 ? 2 : 4);
    for (int i = 0; i < 3; ++i) {
      int rc = aewn_tgemm(g, 0);
      if (rc) printf("  tgemm rc=%d (%s)\n", rc, aewn_last_error_string());
    }
    CK(cudaDeviceSynchronize());
    const int reps = 10;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) aewn_tgemm(g, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    double flops = (which & 1) == 0 ? 2.0 * B * (T - d_) * 512.0 * (2 * R + Cc) : 2.0 * B * ((T - d_) * 368.0 + 16384.0 * 256) * 256;
    printf("  time %s cluster=%d: %.3f ms  -> %.1f TFLOP/s (useful)  device_err=%d\n",
           (which & 1) == 0 ? "GEMM1+gate" : "GEMM2+res+skip", g->cluster, ms, flops / ms * 1e-9, read_err());
  }

  // wgrad for the layer: G = g_fg (512 rows), X = x@-d, x@0, cond
  float* gfg;
  float* dwf;
  CK(cudaMalloc(&gfg, (size_t)B * 512 * pitch * 4));
  CK(cudaMemset(gfg, 0, (size_t)B * 512 * pitch * 4));
  CK(cudaMalloc(&dwf, (size_t)512 * (2 * R + Cc) * 4));
  CK(cudaMemset(dwf, 0, (size_t)512 * (2 * R + Cc) * 4));
  aewn_wgrad_desc wd;
  memset(&wd, 0, sizeof(wd));
  wd.acts[0] = {gfg, T, 512, B, pitch, (long long)512 * pitch};
  wd.acts[1] = {x, T, R, B, pitch, (long long)R * pitch};
  wd.acts[2] = {cond, T, Cc, B, pitch, (long long)Cc * pitch};
  wd.n_acts = 3;
  int ni = 0;
  for (int mt = 0; mt < 4; ++mt) {
    struct {
      int xa, xrow, n, nv, shift, col;
    } cols[5] = {{1, 0, 256, 256, -d_, 0}, {1, 256, 112, 112, -d_, 256}, {1, 0, 256, 256, 0, 368}, {1, 256, 112, 112, 0, 624},
                 {2, 0, 144, 139, 0, 736}};
    for (int c = 0; c < 5; ++c) {
      aewn_wgrad_item& im = wd.items[ni++];
      im.g_act = 0;
      im.x_act = cols[c].xa;
      im.g_row = mt * 128;
      im.x_row = cols[c].xrow;
      im.m_valid = 128;
      im.n = cols[c].n;
      im.n_valid = cols[c].nv;
      im.shift = cols[c].shift;
      im.t_lo = d_;
      im.t_hi = T;
      im.n_split = 14;
      im.out = dwf + (size_t)mt * 128 * (2 * R + Cc) + cols[c].col;
      im.out_rs = 2 * R + Cc;
      im.out_cs = 1;
    }
  }
  wd.n_items = ni;
  wd.batch = B;
  wd.err = g_err;
  for (int i = 0; i < 2; ++i) {
    int rc = aewn_wgrad(&wd, 0);
    if (rc) printf("  wgrad rc=%d (%s)\n", rc, aewn_last_error_string());
  }
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 5; ++i) aewn_wgrad(&wd, 0);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= 5;
  printf("  time wgrad(512 x 875, K=%d): %.3f ms -> %.1f TFLOP/s  device_err=%d\n", B * (T - d_), ms,
         2.0 * B * (T - d_) * 512.0 * (2 * R + Cc) / ms * 1e-9, read_err());
}

int main(int argc, char** argv) {
  bool quick = argc > 1 && !strcmp(argv[1], "quick");
  if (argc > 1 && strcmp(argv[1], "quick")) {
    CK(cudaMalloc(&g_err, 4));
    CK(cudaMemset(g_err, 0, 4));
    const char* t = argv[1];
    if (!strcmp(t, "shift")) { g_shift = atoi(argv[2]); return test_basic(0, 0, true, 1e-3); }
    if (!strcmp(t, "layer")) return test_layer(true, 1e-3) + test_layer(true, 1e-3, 4) + test_layer(false, 5e-2, 4);
    if (!strcmp(t, "layer1")) { test_layer(true, 1e9, 1); return 0; }
    if (!strcmp(t, "layer2")) { test_layer(true, 1e9, 2); return 0; }
    if (!strcmp(t, "layer3")) { test_layer(true, 1e9, 3); return 0; }
    if (!strcmp(t, "gate")) return test_gate();
    if (!strcmp(t, "wgrad")) return test_wgrad(true, 1e-3) + test_wgrad(false, 5e-2);
    if (!strcmp(t, "time")) { time_layer(); return 0; }
    printf("unknown test %s\n", t);
    return 2;
  }
  CK(cudaMalloc(&g_err, 4));
  CK(cudaMemset(g_err, 0, 4));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  int fails = 0;
  printf("[1] tgemm descriptor candidates (integer data, exact)\n");
  const int cand[][2] = {{4096, 512}, {512, 4096}, {4096, 1024}, {1024, 4096}, {4096, 128}, {128, 4096}};
  int first_ok = -1;
  for (int i = 0; i < 6; ++i) {
    int r = test_basic(cand[i][0], cand[i][1], true, 1e-3);
    if (r == 0 && first_ok < 0) first_ok = i;
  }
  if (first_ok != 0) {
    printf("  default descriptor (4096,512) did not pass; first passing candidate index = %d\n", first_ok);
    fails++;
  }
  printf("[1b] tgemm basic, random floats (TF32 rounding error expected ~1e-3)\n");
  test_basic(0, 0, false, 2e-2);
  printf("[2] tgemm layer-shaped (shifts, 3 segments, 2 n-tiles, bias/add/relu/accumulate, NaN padding)\n");
  fails += test_layer(true, 1e-3);
  fails += test_layer(false, 5e-2);
  printf("[3] gate epilogues\n");
  fails += test_gate();
  printf("[4] wgrad\n");
  fails += test_wgrad(true, 1e-3);
  fails += test_wgrad(false, 5e-2);
  if (!quick) {
    printf("[5] timing at cfg2 layer-0 size\n");
    time_layer();
  }
  printf("PROBE %s (%d failing groups)\n", fails ? "FAILED" : "OK", fails);
  return fails ? 1 : 0;
}
