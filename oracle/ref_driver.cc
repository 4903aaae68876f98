// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Thin extern "C" driver over the UNMODIFIED CPU reference model of
// ebruneton/precomputed_atmospheric_scattering. It is compiled by
// oracle/Makefile together with the reference's own translation unit
// /root/reference/atmosphere/reference/functions.cc (which #includes
// atmosphere/functions.glsl as C++, functions.cc:43-54); no reference source
// is copied into this repository. The result, oracle/_ref/libpas_ref.so, is
//   (1) the ground truth the fp64 restatement in oracle/pas_oracle.c is pinned
//       against (tests/test_oracle_vs_reference.py, oracle/gen_golden.py), and
//   (2) the "reference" CPU arm timed by bench.py (--impl reference and the
//       cpu_baseline leg).
//
// The phase sequence and the per-texel calls follow
// atmosphere/reference/model.cc:140-237 (Model::Init). The reference keeps its
// intermediate textures private, so this driver calls the same public
// Compute*Texture functions (atmosphere/reference/functions.h:53-244) itself
// and lets the caller read every intermediate back. Spectra are 47 independent
// lanes (atmosphere/reference/definitions.h:110-112); callers pack the C
// channels they care about into lanes 0..C-1 exactly like the reference's unit
// tests do (atmosphere/reference/functions_test.cc:301-311).
//
// Work distribution: the reference's RunJobs hard-codes 8 threads
// (external/progress_bar/util/progress_bar.cc:75). Here the thread count is a
// parameter so that the CPU baseline can use every host core.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "atmosphere/reference/definitions.h"
#include "atmosphere/reference/functions.h"

// Defined (with external linkage) by functions.glsl:905-926 but not declared
// in atmosphere/reference/functions.h.
namespace atmosphere {
namespace reference {
void GetRMuMuSNuFromScatteringTextureFragCoord(
    const AtmosphereParameters& atmosphere, const vec3& frag_coord,
    Length& r, Number& mu, Number& mu_s, Number& nu,
    bool& ray_r_mu_intersects_ground);
}  // namespace reference
}  // namespace atmosphere

namespace {

using namespace atmosphere;             // NOLINT
using namespace atmosphere::reference;  // NOLINT

constexpr int kLanes = 47;
constexpr int kNT = TRANSMITTANCE_TEXTURE_WIDTH * TRANSMITTANCE_TEXTURE_HEIGHT;
constexpr int kNE = IRRADIANCE_TEXTURE_WIDTH * IRRADIANCE_TEXTURE_HEIGHT;
constexpr int kNS = SCATTERING_TEXTURE_WIDTH * SCATTERING_TEXTURE_HEIGHT *
    SCATTERING_TEXTURE_DEPTH;

struct State {
  AtmosphereParameters atmosphere;
  std::unique_ptr<TransmittanceTexture> transmittance;
  std::unique_ptr<IrradianceTexture> delta_irradiance;
  std::unique_ptr<IrradianceTexture> irradiance;
  std::unique_ptr<ReducedScatteringTexture> delta_rayleigh;
  std::unique_ptr<ReducedScatteringTexture> delta_mie;
  std::unique_ptr<ScatteringDensityTexture> delta_density;
  std::unique_ptr<ScatteringTexture> delta_multiple;
  std::unique_ptr<ReducedScatteringTexture> scattering;
};

DensityProfileLayer MakeLayer(const double* v) {
  return DensityProfileLayer(v[0] * m, v[1], v[2] / m, v[3] / m, v[4]);
}

template <class Job>
double RunRows(int row_count, int stride, int nthreads, Job job) {
  auto t0 = std::chrono::steady_clock::now();
  std::atomic<int> next(0);
  std::vector<std::thread> threads;
  if (nthreads < 1) nthreads = 1;
  for (int t = 0; t < nthreads; ++t) {
    threads.emplace_back([&]() {
      for (;;) {
        int row = next.fetch_add(1) * stride;
        if (row >= row_count) return;
        job(row);
      }
    });
  }
  for (auto& t : threads) t.join();
  return std::chrono::duration<double>(
      std::chrono::steady_clock::now() - t0).count();
}

template <class Spectrum>
void ZeroSpectrum(Spectrum* s) {
  for (int i = 0; i < kLanes; ++i) (*s)[i] = (*s)[i] * 0.0;
}

}  // namespace

extern "C" {

// out[0..7] = T width, T height, R, MU, MU_S, NU, E width, E height: the
// compile-time sizes of atmosphere/constants.h:47-61.
void pasref_sizes(int* out) {
  out[0] = TRANSMITTANCE_TEXTURE_WIDTH;
  out[1] = TRANSMITTANCE_TEXTURE_HEIGHT;
  out[2] = SCATTERING_TEXTURE_R_SIZE;
  out[3] = SCATTERING_TEXTURE_MU_SIZE;
  out[4] = SCATTERING_TEXTURE_MU_S_SIZE;
  out[5] = SCATTERING_TEXTURE_NU_SIZE;
  out[6] = IRRADIANCE_TEXTURE_WIDTH;
  out[7] = IRRADIANCE_TEXTURE_HEIGHT;
}

// spectra: 6 arrays of nlanes values, in this order: solar_irradiance,
// rayleigh_scattering, mie_scattering, mie_extinction, absorption_extinction,
// ground_albedo. scalars: sun_angular_radius, bottom_radius, top_radius,
// mie_phase_function_g, mu_s_min. profiles: rayleigh, mie, absorption, each
// 2 layers x (width, exp_term, exp_scale, linear_term, constant_term).
// All lengths in one consistent unit of the caller's choice (the dimensional
// types only check homogeneity; definitions.h:74-101).
void* pasref_create(int nlanes, const double* spectra, const double* scalars,
                    const double* profiles) {
  if (nlanes < 1 || nlanes > kLanes) return nullptr;
  State* s = new State();
  AtmosphereParameters& a = s->atmosphere;
  for (int i = 0; i < kLanes; ++i) {
    int l = i < nlanes ? i : 0;  // unused lanes replicate lane 0 (stay finite)
    a.solar_irradiance[i] =
        spectra[0 * nlanes + l] * watt_per_square_meter_per_nm;
    a.rayleigh_scattering[i] = spectra[1 * nlanes + l] / m;
    a.mie_scattering[i] = spectra[2 * nlanes + l] / m;
    a.mie_extinction[i] = spectra[3 * nlanes + l] / m;
    a.absorption_extinction[i] = spectra[4 * nlanes + l] / m;
    a.ground_albedo[i] = spectra[5 * nlanes + l];
  }
  a.sun_angular_radius = scalars[0] * rad;
  a.bottom_radius = scalars[1] * m;
  a.top_radius = scalars[2] * m;
  a.mie_phase_function_g = scalars[3];
  a.mu_s_min = scalars[4];
  for (int l = 0; l < 2; ++l) {
    a.rayleigh_density.layers[l] = MakeLayer(profiles + 0 + 5 * l);
    a.mie_density.layers[l] = MakeLayer(profiles + 10 + 5 * l);
    a.absorption_density.layers[l] = MakeLayer(profiles + 20 + 5 * l);
  }
  s->transmittance.reset(new TransmittanceTexture());
  s->delta_irradiance.reset(new IrradianceTexture());
  s->irradiance.reset(new IrradianceTexture());
  s->delta_rayleigh.reset(new ReducedScatteringTexture());
  s->delta_mie.reset(new ReducedScatteringTexture());
  s->delta_density.reset(new ScatteringDensityTexture());
  s->delta_multiple.reset(new ScatteringTexture());
  s->scattering.reset(new ReducedScatteringTexture());
  // Deterministic contents for partially computed (benchmark) runs.
  IrradianceSpectrum zero_e(0.0 * watt_per_square_meter_per_nm);
  s->irradiance->Set(zero_e);
  s->delta_irradiance->Set(zero_e);
  s->delta_rayleigh->Set(zero_e);
  s->delta_mie->Set(zero_e);
  s->scattering->Set(zero_e);
  s->delta_density->Set(
      RadianceDensitySpectrum(0.0 * watt_per_cubic_meter_per_sr_per_nm));
  s->delta_multiple->Set(
      RadianceSpectrum(0.0 * watt_per_square_meter_per_sr_per_nm));
  s->transmittance->Set(DimensionlessSpectrum(1.0));
  return s;
}

void pasref_destroy(void* handle) { delete static_cast<State*>(handle); }

// Runs one phase of atmosphere/reference/model.cc:140-237 over the rows
// {0, stride, 2*stride, ...}. A row is one j for the 2D tables and one (k, j)
// pair (row = k * MU + j) for the 3D tables. stride == 1 is the full phase
// and also performs the phase's accumulation into the final tables.
//   phase 0: transmittance                      (model.cc:140-147)
//   phase 1: direct irradiance, irradiance = 0  (model.cc:152-161)
//   phase 2: single scattering                  (model.cc:166-179)
//   phase 3: scattering density of `order`      (model.cc:187-200)
//   phase 4: indirect irradiance from order-1   (model.cc:204-215)
//   phase 5: multiple scattering                (model.cc:220-237)
// Returns the wall-clock seconds spent, or -1 on a bad argument.
double pasref_phase(void* handle, int phase, int order, int stride,
                    int nthreads) {
  State* s = static_cast<State*>(handle);
  if (s == nullptr || stride < 1) return -1.0;
  const AtmosphereParameters& atm = s->atmosphere;
  const int W = SCATTERING_TEXTURE_WIDTH;
  const int H = SCATTERING_TEXTURE_HEIGHT;
  const int D = SCATTERING_TEXTURE_DEPTH;
  switch (phase) {
    case 0:
      return RunRows(TRANSMITTANCE_TEXTURE_HEIGHT, stride, nthreads,
          [&](int j) {
            for (int i = 0; i < TRANSMITTANCE_TEXTURE_WIDTH; ++i) {
              s->transmittance->Set(i, j,
                  ComputeTransmittanceToTopAtmosphereBoundaryTexture(
                      atm, vec2(i + 0.5, j + 0.5)));
            }
          });
    case 1:
      return RunRows(IRRADIANCE_TEXTURE_HEIGHT, stride, nthreads,
          [&](int j) {
            for (int i = 0; i < IRRADIANCE_TEXTURE_WIDTH; ++i) {
              s->delta_irradiance->Set(i, j, ComputeDirectIrradianceTexture(
                  atm, *s->transmittance, vec2(i + 0.5, j + 0.5)));
              s->irradiance->Set(i, j,
                  IrradianceSpectrum(0.0 * watt_per_square_meter_per_nm));
            }
          });
    case 2:
      return RunRows(H * D, stride, nthreads, [&](int row) {
        int k = row / H, j = row % H;
        for (int i = 0; i < W; ++i) {
          IrradianceSpectrum rayleigh, mie;
          ComputeSingleScatteringTexture(atm, *s->transmittance,
              vec3(i + 0.5, j + 0.5, k + 0.5), rayleigh, mie);
          s->delta_rayleigh->Set(i, j, k, rayleigh);
          s->delta_mie->Set(i, j, k, mie);
          s->scattering->Set(i, j, k, rayleigh);
        }
      });
    case 3:
      if (order < 2) return -1.0;
      return RunRows(H * D, stride, nthreads, [&](int row) {
        int k = row / H, j = row % H;
        for (int i = 0; i < W; ++i) {
          s->delta_density->Set(i, j, k, ComputeScatteringDensityTexture(
              atm, *s->transmittance, *s->delta_rayleigh, *s->delta_mie,
              *s->delta_multiple, *s->delta_irradiance,
              vec3(i + 0.5, j + 0.5, k + 0.5), order));
        }
      });
    case 4: {
      if (order < 2) return -1.0;
      double t = RunRows(IRRADIANCE_TEXTURE_HEIGHT, stride, nthreads,
          [&](int j) {
            for (int i = 0; i < IRRADIANCE_TEXTURE_WIDTH; ++i) {
              s->delta_irradiance->Set(i, j, ComputeIndirectIrradianceTexture(
                  atm, *s->delta_rayleigh, *s->delta_mie, *s->delta_multiple,
                  vec2(i + 0.5, j + 0.5), order - 1));
            }
          });
      if (stride == 1) (*s->irradiance) += *s->delta_irradiance;
      return t;
    }
    case 5:
      return RunRows(H * D, stride, nthreads, [&](int row) {
        int k = row / H, j = row % H;
        for (int i = 0; i < W; ++i) {
          Number nu;
          RadianceSpectrum delta = ComputeMultipleScatteringTexture(
              atm, *s->transmittance, *s->delta_density,
              vec3(i + 0.5, j + 0.5, k + 0.5), nu);
          s->delta_multiple->Set(i, j, k, delta);
          if (stride == 1) {
            s->scattering->Set(i, j, k, s->scattering->Get(i, j, k) +
                delta * (1.0 / RayleighPhaseFunction(nu)));
          }
        }
      });
    default:
      return -1.0;
  }
}

// Copies lanes 0..nlanes-1 of one table into dst as planar doubles,
// dst[lane * texels + texel], texel = i + j*NX (+ k*NX*NY) like
// external/dimensional_types/math/binary_function.h:73-83 and
// ternary_function.h:74-79. table: 0 transmittance, 1 delta_irradiance,
// 2 irradiance, 3 delta_rayleigh, 4 delta_mie, 5 delta_density,
// 6 delta_multiple, 7 scattering. Returns the number of texels, or -1.
int pasref_read(void* handle, int table, int nlanes, double* dst) {
  State* s = static_cast<State*>(handle);
  if (s == nullptr || nlanes < 1 || nlanes > kLanes) return -1;
  const int W = SCATTERING_TEXTURE_WIDTH, H = SCATTERING_TEXTURE_HEIGHT;
  auto copy2 = [&](auto& tex, auto unit, int nx, int ny) {
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      const auto& v = tex.Get(i, j);
      for (int l = 0; l < nlanes; ++l) {
        dst[static_cast<size_t>(l) * nx * ny + i + j * nx] = v[l].to(unit);
      }
    }
    return nx * ny;
  };
  auto copy3 = [&](auto& tex, auto unit) {
    for (int k = 0; k < SCATTERING_TEXTURE_DEPTH; ++k)
      for (int j = 0; j < H; ++j) for (int i = 0; i < W; ++i) {
        const auto& v = tex.Get(i, j, k);
        size_t t = i + static_cast<size_t>(W) * (j + static_cast<size_t>(H) * k);
        for (int l = 0; l < nlanes; ++l) {
          dst[static_cast<size_t>(l) * kNS + t] = v[l].to(unit);
        }
      }
    return kNS;
  };
  switch (table) {
    case 0: return copy2(*s->transmittance, Number(1.0),
        TRANSMITTANCE_TEXTURE_WIDTH, TRANSMITTANCE_TEXTURE_HEIGHT);
    case 1: return copy2(*s->delta_irradiance, watt_per_square_meter_per_nm,
        IRRADIANCE_TEXTURE_WIDTH, IRRADIANCE_TEXTURE_HEIGHT);
    case 2: return copy2(*s->irradiance, watt_per_square_meter_per_nm,
        IRRADIANCE_TEXTURE_WIDTH, IRRADIANCE_TEXTURE_HEIGHT);
    case 3: return copy3(*s->delta_rayleigh, watt_per_square_meter_per_nm);
    case 4: return copy3(*s->delta_mie, watt_per_square_meter_per_nm);
    case 5: return copy3(*s->delta_density,
        watt_per_cubic_meter_per_sr_per_nm);
    case 6: return copy3(*s->delta_multiple,
        watt_per_square_meter_per_sr_per_nm);
    case 7: return copy3(*s->scattering, watt_per_square_meter_per_nm);
    default: return -1;
  }
}

// Inverse of pasref_read: loads lanes 0..nlanes-1 of one table from planar
// doubles (same layout); the remaining lanes replicate lane 0 so that they stay
// finite. Lets the reference's render functions run on tables produced
// elsewhere (e.g. by the CUDA path), which isolates render parity from
// precompute parity. Returns the number of texels, or -1.
int pasref_write(void* handle, int table, int nlanes, const double* src) {
  State* s = static_cast<State*>(handle);
  if (s == nullptr || nlanes < 1 || nlanes > kLanes) return -1;
  const int W = SCATTERING_TEXTURE_WIDTH, H = SCATTERING_TEXTURE_HEIGHT;
  auto load2 = [&](auto& tex, auto unit, int nx, int ny) {
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      auto v = tex.Get(i, j);
      for (int l = 0; l < kLanes; ++l) {
        int ls = l < nlanes ? l : 0;
        v[l] = src[static_cast<size_t>(ls) * nx * ny + i + j * nx] * unit;
      }
      tex.Set(i, j, v);
    }
    return nx * ny;
  };
  auto load3 = [&](auto& tex, auto unit) {
    for (int k = 0; k < SCATTERING_TEXTURE_DEPTH; ++k)
      for (int j = 0; j < H; ++j) for (int i = 0; i < W; ++i) {
        auto v = tex.Get(i, j, k);
        size_t t = i + static_cast<size_t>(W) * (j + static_cast<size_t>(H) * k);
        for (int l = 0; l < kLanes; ++l) {
          int ls = l < nlanes ? l : 0;
          v[l] = src[static_cast<size_t>(ls) * kNS + t] * unit;
        }
        tex.Set(i, j, k, v);
      }
    return kNS;
  };
  switch (table) {
    case 0: return load2(*s->transmittance, Number(1.0),
        TRANSMITTANCE_TEXTURE_WIDTH, TRANSMITTANCE_TEXTURE_HEIGHT);
    case 2: return load2(*s->irradiance, watt_per_square_meter_per_nm,
        IRRADIANCE_TEXTURE_WIDTH, IRRADIANCE_TEXTURE_HEIGHT);
    case 4: return load3(*s->delta_mie, watt_per_square_meter_per_nm);
    case 7: return load3(*s->scattering, watt_per_square_meter_per_nm);
    default: return -1;
  }
}

}  // extern "C"

// ---- render-time functions and the test scene -------------------------------
// The reference's own CPU renderer (atmosphere/reference/model_test.cc:632-738)
// views atmosphere/reference/model_test.glsl as C++ inside its test fixture,
// the "uniforms" being fields of the fixture. Same construction here: the GLSL
// file is included where it lies, unmodified; the four model entry points call
// the reference's GetSkyRadiance / GetSkyRadianceToPoint /
// GetSunAndSkyIrradiance (atmosphere/reference/functions.h:223-244) on this
// driver's tables exactly like reference::Model does
// (atmosphere/reference/model.cc:255-283; single Mie = delta_mie).
namespace atmosphere {
namespace reference {
namespace {

using std::max;  // as atmosphere/reference/functions.cc:49-50 does for the GLSL text
using std::min;

struct SceneRenderer {
  const State* state;
  Position kSphereCenter;
  Length kSphereRadius;
  Position camera_;
  Position earth_center_;
  Direction sun_direction_;
  dimensional::vec2 sun_size_;
  DimensionlessSpectrum ground_albedo_;
  DimensionlessSpectrum sphere_albedo_;

  RadianceSpectrum GetSolarRadiance() {
    // atmosphere/reference/model.cc:255-259
    SolidAngle sun_solid_angle = 2.0 * PI *
        (1.0 - cos(state->atmosphere.sun_angular_radius)) * sr;
    return state->atmosphere.solar_irradiance * (1.0 / sun_solid_angle);
  }
  RadianceSpectrum GetSkyRadiance(Position camera, Direction view_ray,
      Length shadow_length, Direction sun_direction,
      DimensionlessSpectrum& transmittance) {
    return reference::GetSkyRadiance(state->atmosphere, *state->transmittance,
        *state->scattering, *state->delta_mie, camera, view_ray, shadow_length,
        sun_direction, transmittance);
  }
  RadianceSpectrum GetSkyRadianceToPoint(Position camera, Position point,
      Length shadow_length, Direction sun_direction,
      DimensionlessSpectrum& transmittance) {
    return reference::GetSkyRadianceToPoint(state->atmosphere,
        *state->transmittance, *state->scattering, *state->delta_mie, camera,
        point, shadow_length, sun_direction, transmittance);
  }
  IrradianceSpectrum GetSunAndSkyIrradiance(Position point, Direction normal,
      Direction sun_direction, IrradianceSpectrum& sky_irradiance) {
    return reference::GetSunAndSkyIrradiance(state->atmosphere,
        *state->transmittance, *state->irradiance, point, normal,
        sun_direction, sky_irradiance);
  }

#define OUT(x) x&
#include "atmosphere/reference/model_test.glsl"
#undef OUT
};

}  // namespace
}  // namespace reference
}  // namespace atmosphere

extern "C" {

// Renders the test scene of atmosphere/reference/model_test.glsl with the
// reference's CPU functions on the tables currently held by `handle`
// (transmittance, scattering, delta_mie as single Mie, irradiance).
// scene: camera[3], earth_center[3], sun_direction[3], sun_size[2] (tan, cos),
// sphere_center[3], sphere_radius, model_from_clip[9] (row major) = 24 doubles;
// albedos: ground_albedo[nlanes] then sphere_albedo[nlanes].
// out[(j*width + i)*nlanes + l] = spectral radiance of lane l at pixel (i, j),
// j = 0 at the top, view rays as in model_test.cc:688-711.
// Returns wall-clock seconds, or -1.
double pasref_render_scene(void* handle, int nlanes, const double* scene,
                           const double* albedos, int width, int height,
                           int row_stride, int nthreads, double* out) {
  State* s = static_cast<State*>(handle);
  if (s == nullptr || nlanes < 1 || nlanes > kLanes || width < 1 ||
      height < 1 || row_stride < 1) {
    return -1.0;
  }
  atmosphere::reference::SceneRenderer proto;
  proto.state = s;
  proto.camera_ = Position(scene[0] * m, scene[1] * m, scene[2] * m);
  proto.earth_center_ = Position(scene[3] * m, scene[4] * m, scene[5] * m);
  proto.sun_direction_ = Direction(scene[6], scene[7], scene[8]);
  proto.sun_size_ = dimensional::vec2(scene[9], scene[10]);
  proto.kSphereCenter = Position(scene[11] * m, scene[12] * m, scene[13] * m);
  proto.kSphereRadius = scene[14] * m;
  for (int l = 0; l < kLanes; ++l) {
    int ls = l < nlanes ? l : 0;
    proto.ground_albedo_[l] = albedos[ls];
    proto.sphere_albedo_[l] = albedos[nlanes + ls];
  }
  const double* M = scene + 15;
  return RunRows(height, row_stride, nthreads, [&](int j) {
    atmosphere::reference::SceneRenderer r = proto;
    double y = 1.0 - 2.0 * (j + 0.5) / height;
    double dy = -2.0 / height;
    for (int i = 0; i < width; ++i) {
      double x = 2.0 * (i + 0.5) / width - 1.0;
      double dx = 2.0 / width;
      Direction view_ray(M[0] * x + M[1] * y + M[2], M[3] * x + M[4] * y + M[5],
                         M[6] * x + M[7] * y + M[8]);
      Direction view_ray_diff(M[0] * dx + M[1] * dy, M[3] * dx + M[4] * dy,
                              M[6] * dx + M[7] * dy);
      RadianceSpectrum radiance = r.GetViewRayRadiance(view_ray, view_ray_diff);
      double* o = out + (static_cast<size_t>(j) * width + i) * nlanes;
      for (int l = 0; l < nlanes; ++l) {
        o[l] = radiance[l].to(watt_per_square_meter_per_sr_per_nm);
      }
    }
  });
}

// Point forms of the three render-time lookups. vectors are 3 doubles each.
// (atmosphere/reference/functions.h:223-244)
void pasref_sky_radiance(void* handle, int nlanes, const double* camera,
                         const double* view_ray, double shadow_length,
                         const double* sun_direction, double* radiance,
                         double* transmittance) {
  State* s = static_cast<State*>(handle);
  DimensionlessSpectrum t;
  RadianceSpectrum L = GetSkyRadiance(s->atmosphere, *s->transmittance,
      *s->scattering, *s->delta_mie,
      Position(camera[0] * m, camera[1] * m, camera[2] * m),
      Direction(view_ray[0], view_ray[1], view_ray[2]), shadow_length * m,
      Direction(sun_direction[0], sun_direction[1], sun_direction[2]), t);
  for (int l = 0; l < nlanes; ++l) {
    radiance[l] = L[l].to(watt_per_square_meter_per_sr_per_nm);
    transmittance[l] = t[l]();
  }
}

void pasref_sky_radiance_to_point(void* handle, int nlanes,
                                  const double* camera, const double* point,
                                  double shadow_length,
                                  const double* sun_direction, double* radiance,
                                  double* transmittance) {
  State* s = static_cast<State*>(handle);
  DimensionlessSpectrum t;
  RadianceSpectrum L = GetSkyRadianceToPoint(s->atmosphere, *s->transmittance,
      *s->scattering, *s->delta_mie,
      Position(camera[0] * m, camera[1] * m, camera[2] * m),
      Position(point[0] * m, point[1] * m, point[2] * m), shadow_length * m,
      Direction(sun_direction[0], sun_direction[1], sun_direction[2]), t);
  for (int l = 0; l < nlanes; ++l) {
    radiance[l] = L[l].to(watt_per_square_meter_per_sr_per_nm);
    transmittance[l] = t[l]();
  }
}

void pasref_sun_and_sky_irradiance(void* handle, int nlanes,
                                   const double* point, const double* normal,
                                   const double* sun_direction,
                                   double* sun_irradiance,
                                   double* sky_irradiance) {
  State* s = static_cast<State*>(handle);
  IrradianceSpectrum sky;
  IrradianceSpectrum sun = GetSunAndSkyIrradiance(s->atmosphere,
      *s->transmittance, *s->irradiance,
      Position(point[0] * m, point[1] * m, point[2] * m),
      Direction(normal[0], normal[1], normal[2]),
      Direction(sun_direction[0], sun_direction[1], sun_direction[2]), sky);
  for (int l = 0; l < nlanes; ++l) {
    sun_irradiance[l] = sun[l].to(watt_per_square_meter_per_nm);
    sky_irradiance[l] = sky[l].to(watt_per_square_meter_per_nm);
  }
}

// Point evaluations of reference functions, used to pin the restatement's
// mappings texel by texel without running a whole phase.
// (functions.glsl:773-831, 837-890 via functions.h)
void pasref_uvwz_from_rmumusnu(void* handle, double r, double mu, double mu_s,
                               double nu, int hits_ground, double* uvwz) {
  State* s = static_cast<State*>(handle);
  vec4 v = GetScatteringTextureUvwzFromRMuMuSNu(
      s->atmosphere, r * m, mu, mu_s, nu, hits_ground != 0);
  uvwz[0] = v.x(); uvwz[1] = v.y(); uvwz[2] = v.z(); uvwz[3] = v.w();
}

void pasref_rmumusnu_from_frag_coord(void* handle, double x, double y,
                                     double z, double* out) {
  State* s = static_cast<State*>(handle);
  Length r; Number mu, mu_s, nu; bool hit;
  GetRMuMuSNuFromScatteringTextureFragCoord(
      s->atmosphere, vec3(x, y, z), r, mu, mu_s, nu, hit);
  out[0] = r.to(m); out[1] = mu(); out[2] = mu_s(); out[3] = nu();
  out[4] = hit ? 1.0 : 0.0;
}

}  // extern "C"
