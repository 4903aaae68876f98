// TEST INFRASTRUCTURE: compiles csrc/pas_physics.cuh -- the host/device inline physics every kernel
// shares -- with a plain C++ compiler and exports it through a C interface, so that the table
// mappings the kernels use can be checked on the CPU against the oracle (tests/test_physics_header.py).
// Nothing here is part of the product.
#include "../../precomputed_atmospheric_scattering_b200/csrc/pas_physics.cuh"

using namespace pas;

extern "C" {

// layer k -> (r, rho); row j at that layer -> (mu, hit); column i -> mu_s; slab -> clamped nu
void emu_texel(const PasGeometry* g, int k, int j, int i_mu_s, int i_nu, double* out5) {
  double r, rho, mu, r_mu;
  bool hit;
  layer_radius(*g, (k + 0.5) / g->sz.r_n, g->sz.r_n, &r, &rho);
  scattering_row_mu(*g, r, rho, j, &mu, &r_mu, &hit);
  const double mu_s = scattering_col_mu_s(*g, i_mu_s);
  out5[0] = r;
  out5[1] = mu;
  out5[2] = mu_s;
  out5[3] = scattering_slab_nu(*g, i_nu, mu, mu_s);
  out5[4] = hit ? 1.0 : 0.0;
}

// forward maps in texel space (u * n - 0.5)
double emu_y_from_mu(const PasGeometry* g, double r, double mu, int hit) {
  const double rho = sqrt(d_pos(r * r - g->bottom * g->bottom));
  return scattering_y_from_mu(*g, r, rho, mu, hit != 0);
}
double emu_x_from_mu_s(const PasGeometry* g, double mu_s) { return scattering_x_from_mu_s(*g, mu_s); }
void emu_transmittance_xy(const PasGeometry* g, double r, double mu, double* xy) {
  transmittance_xy(*g, r, mu, &xy[0], &xy[1]);
}

// the fp32 inner-loop forms
float emu_f_mu_s_texel_x(const PasGeometry* g, float mu_s) {
  MuSMap m;
  m.H2 = (float)(g->H * g->H);
  m.d_min = (float)(g->top - g->bottom);
  m.inv_range = (float)(1.0 / (g->H - (g->top - g->bottom)));
  m.inv_A = (float)(1.0 / g->mus_A);
  m.scale = (float)(g->sz.mu_s_n - 1);
  return f_mu_s_texel_x(m, (float)g->bottom * mu_s);
}
// as the ray-march kernels call it: p = r mu and q = (top - r)(top + r) are formed in fp64 once per
// sample (kernel_raymarch.cu, fill_sample), the root is taken in fp32 per thread
float emu_f_dist_top(const PasGeometry* g, double r, double mu) {
  return f_dist_top((float)(r * mu), (float)((g->top - r) * (g->top + r)));
}
void emu_make_tap(double x, int n, int* i0, int* i1, float* w) {
  const Tap t = make_tap(x, n);
  *i0 = t.i0; *i1 = t.i1; *w = t.w;
}
void emu_make_tap_f(float x, int n, int* i0, int* i1, float* w) {
  const Tap t = make_tap_f(x, n);
  *i0 = t.i0; *i1 = t.i1; *w = t.w;
}
double emu_dist_top(const PasGeometry* g, double r, double mu) { return dist_top(*g, r, mu); }
double emu_dist_bottom(const PasGeometry* g, double r, double mu) { return dist_bottom(*g, r, mu); }
int emu_hits_ground(const PasGeometry* g, double r, double mu) { return hits_ground(*g, r, mu) ? 1 : 0; }
double emu_profile_density(const PasGeometry* g, int profile, double h) {
  return profile_density(g->profiles[profile], h);
}
int emu_sizeof_geometry() { return (int)sizeof(PasGeometry); }

}  // extern "C"
