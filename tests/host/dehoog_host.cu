// Host-side harness: runs the SAME De Hoog / head device functions the CUDA kernels use, compiled
// for the CPU, so the recurrence and its hand-written reverse sweep can be checked against the
// oracle without a GPU (tests/test_host_math.py).
#include <vector>

#include "../../neurallaplace_b200/csrc/nl_dehoog.cuh"
#include "../../neurallaplace_b200/csrc/nl_dehoog_sm.cuh"

using namespace nl;

extern "C" {

// a: [n] complex interleaved; returns w (2 doubles) and dw/da_k (n complex interleaved)
int nlh_dehoog(int n, double zr, double zi, const double* a, double* w_out, double* dw_out, double* w_fwd_only) {
  const int M = (n - 1) / 2;
  std::vector<Cx<double>> buf((size_t)dehoog_scratch_entries(n, true) + 8);
  Scratch<double> scr{buf.data(), 1, 0};
  auto acc = [&](int k) { return Cx<double>{a[2 * k], a[2 * k + 1]}; };
  auto emit = [&](int k, Cx<double> g) {
    dw_out[2 * k] = g.re;
    dw_out[2 * k + 1] = g.im;
  };
  Cx<double> w = dehoog_forward_backward<double>(M, Cx<double>{zr, zi}, acc, emit, scr);
  w_out[0] = w.re;
  w_out[1] = w.im;
  std::vector<Cx<double>> buf2((size_t)dehoog_scratch_entries(n, false) + 8);
  Scratch<double> scr2{buf2.data(), 1, 0};
  Cx<double> w2 = dehoog_forward<double>(M, Cx<double>{zr, zi}, acc, scr2);
  w_fwd_only[0] = w2.re;
  w_fwd_only[1] = w2.im;
  return 0;
}

int nlh_dehoog_f32(int n, float zr, float zi, const float* a, float* w_out) {
  const int M = (n - 1) / 2;
  std::vector<Cx<float>> buf((size_t)dehoog_scratch_entries(n, false) + 8);
  Scratch<float> scr{buf.data(), 1, 0};
  auto acc = [&](int k) { return Cx<float>{a[2 * k], a[2 * k + 1]}; };
  Cx<float> w = dehoog_forward<float>(M, Cx<float>{zr, zi}, acc, scr);
  w_out[0] = w.re;
  w_out[1] = w.im;
  return 0;
}

// shared-memory-strip variant (runtime M, rolled loops): the one the tensor-core De Hoog path launches
int nlh_dehoog_sm(int n, double zr, double zi, const double* a, double* w_out, double* dw_out, double* w_fwd_only) {
  const int M = (n - 1) / 2;
  std::vector<Cx<double>> colv(dehoog_sm_col_entries(n)), tabv(dehoog_sm_tab_entries(n));
  auto acc = [&](int k) { return Cx<double>{a[2 * k], a[2 * k + 1]}; };
  auto emit = [&](int k, Cx<double> g) {
    dw_out[2 * k] = g.re;
    dw_out[2 * k + 1] = g.im;
  };
  auto col = [&](int i) -> Cx<double>& { return colv.at(i); };
  auto tab = [&](int i) -> Cx<double>& { return tabv.at(i); };
  auto pf = [&](int i) { (void)tabv.at(i); };
  auto pf0 = [&](int) {};
  Cx<double> w = dehoog_sm<double, true>(M, Cx<double>{zr, zi}, acc, emit, col, tab, pf, pf0);
  w_out[0] = w.re;
  w_out[1] = w.im;
  Cx<double> w2 = dehoog_sm<double, false>(M, Cx<double>{zr, zi}, acc, emit, col, tab, pf, pf0);
  w_fwd_only[0] = w2.re;
  w_fwd_only[1] = w2.im;
  return 0;
}


// fixed contour (nl_dehoog_fixed.cu composes exactly these parts): a[n] -> table once; per time point z_j the Pade
// recurrence + its reverse; g_j-weighted sum of dw/dd_i over j; one reverse sweep through the table.
// out_w[2*nz] = w_j;  out_G[2n] = G_k = sum_j g_j dw_j/da_k (holomorphic);  w_fwd[2*nz]: forward-only parts.
int nlh_dehoog_fixed(int n, int nz, const double* z, const double* g, const double* a, double* out_w, double* out_G,
                     double* w_fwd) {
  const int M = (n - 1) / 2, LD = 2 * M + 2;
  auto acc = [&](int k) { return Cx<double>{a[2 * k], a[2 * k + 1]}; };
  std::vector<Cx<double>> tabv((size_t)(2 * M + 2) * LD), thr((size_t)3 * LD), adjv((size_t)5 * LD), fwd((size_t)3 * LD);
  Scratch<double> tab{tabv.data(), 1, 0}, dbar{thr.data(), 1, 0}, adj{adjv.data(), 1, 0}, fw{fwd.data(), 1, 0};
  dehoog_qd_store<double>(M, acc, tab);
  dehoog_qd_coeffs<double>(M, acc, fw);
  const Scratch<double> dd = tab.at((2 * M + 1) * LD), dd2 = fw.at(2 * LD);
  for (int i = 0; i <= 2 * M; ++i) adj[i] = Cx<double>{0.0, 0.0};
  for (int j = 0; j < nz; ++j) {
    const Cx<double> zj{z[2 * j], z[2 * j + 1]};
    Cx<double> w = dehoog_pade_fwd_bwd<double>(M, zj, [&](int i) { return dd[i]; }, dbar, dbar.at(LD));
    out_w[2 * j] = w.re;
    out_w[2 * j + 1] = w.im;
    Cx<double> w2 = dehoog_pade<double>(M, zj, [&](int i) { return dd2[i]; });
    w_fwd[2 * j] = w2.re;
    w_fwd[2 * j + 1] = w2.im;
    for (int i = 0; i <= 2 * M; ++i) adj[i] = cadd(adj[i], cscale(dbar[i], g[j]));
  }
  auto emit = [&](int k, Cx<double> G) {
    out_G[2 * k] = G.re;
    out_G[2 * k + 1] = G.im;
  };
  dehoog_qd_reverse<double>(M, acc, emit, tab, adj, adj.at(LD));
  return 0;
}

// head forward/backward:  in (oa, ob, gre, gim) -> out (fre, fim, goa, gob)
int nlh_head(int head, double oa, double ob, double gre, double gim, double* out) {
  HeadOut<double> h = head_forward<double>(head, oa, ob);
  out[0] = h.fre;
  out[1] = h.fim;
  head_backward<double>(head, h, gre, gim, &out[2], &out[3]);
  return 0;
}

// s point, features and reduction coefficient for (family, k, t)
int nlh_query(int family, int head, int n, double alpha, double log_tol, double scale, const double* beta,
              const double* eta, double t, int k, double* out) {
  IltParams<double> P{family, n, head, alpha, log_tol, scale, beta, eta};
  TimeCtx<double> c = make_time_ctx<double>(P, t, nullptr, 0);
  Cx<double> s = s_point(P, c, k);
  double f0, f1;
  s_features(P, s, &f0, &f1);
  Cx<double> w = reduce_coef(P, c, k);
  out[0] = s.re; out[1] = s.im; out[2] = f0; out[3] = f1; out[4] = w.re; out[5] = w.im; out[6] = c.pref;
  return 0;
}
}
