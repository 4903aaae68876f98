// Host side of the C ABI (include/pas_b200.h): parameter conversion that mirrors
// atmosphere::Model::Model (atmosphere/model.cc:613-795), the pass schedule that mirrors
// Model::Init / Model::Precompute (atmosphere/model.cc:866-975, 1048-1215) with CUDA kernels in
// place of the GL draws, table readback, and the r-slab multi-GPU exchange over NCCL.
//
// There is no CPU implementation of any pass in this library: every table is produced by the
// kernels in kernels_setup.cu, kernel_raymarch.cu, kernel_density.cu and kernel_irradiance.cu.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pas_b200.h"
#include "cie1931.h"
#include "kernel_render.cuh"
#include "pas_kernels.h"
#include "pas_physics.cuh"

namespace {

thread_local std::string g_last_error;

pas_status fail(pas_status code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define PAS_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      return fail(PAS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
    }                                                                                       \
  } while (0)

// ---- NCCL, bound lazily so that the single-GPU path has no NCCL dependency -------------------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi& nccl() {
  static NcclApi api;
  if (api.lib != nullptr || api.ok) return api;
  api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (api.lib == nullptr) return api;
#define PAS_SYM(field, name) *reinterpret_cast<void**>(&api.field) = dlsym(api.lib, name)
  PAS_SYM(GetUniqueId, "ncclGetUniqueId");
  PAS_SYM(CommInitRank, "ncclCommInitRank");
  PAS_SYM(CommDestroy, "ncclCommDestroy");
  PAS_SYM(AllGather, "ncclAllGather");
  PAS_SYM(AllReduce, "ncclAllReduce");
  PAS_SYM(GroupStart, "ncclGroupStart");
  PAS_SYM(GroupEnd, "ncclGroupEnd");
  PAS_SYM(GetErrorString, "ncclGetErrorString");
#undef PAS_SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather &&
           api.AllReduce && api.GroupStart && api.GroupEnd && api.GetErrorString;
  return api;
}
#define PAS_NCCL(expr)                                                                      \
  do {                                                                                      \
    ncclResult_t r__ = (expr);                                                              \
    if (r__ != ncclSuccess) {                                                               \
      return fail(PAS_ERR_NCCL, std::string(#expr) + ": " + nccl().GetErrorString(r__));    \
    }                                                                                       \
  } while (0)

// ---- spectra helpers (atmosphere/model.cc:521-595) -------------------------------------------
// CieColorMatchingFunctionTableValue (model.cc:521-533): linear in the 5 nm table, 0 outside.
double cie_value(double wavelength, int column) {
  if (wavelength <= pas::kCieLambdaMin || wavelength >= pas::kCieLambdaMax) return 0.0;
  double u = (wavelength - pas::kCieLambdaMin) / pas::kCieStep;
  int row = static_cast<int>(std::floor(u));
  u -= row;
  const double* t = column == 1 ? pas::kCieXBar : (column == 2 ? pas::kCieYBar : pas::kCieZBar);
  return t[row] * (1.0 - u) + t[row + 1] * u;
}

// Interpolate (model.cc:535-552): piecewise linear, clamped at both ends.
double interpolate(const std::vector<double>& wl, const std::vector<double>& v, double wavelength) {
  if (wavelength < wl[0]) return v[0];
  for (size_t i = 0; i + 1 < wl.size(); ++i) {
    if (wavelength < wl[i + 1]) {
      double u = (wavelength - wl[i]) / (wl[i + 1] - wl[i]);
      return v[i] * (1.0 - u) + v[i + 1] * u;
    }
  }
  return v.back();
}

// ComputeSpectralRadianceToLuminanceFactors (model.cc:562-595), lumen.nm/watt.
void luminance_factors(const std::vector<double>& wl, const std::vector<double>& solar,
                       double lambda_power, double* k) {
  k[0] = k[1] = k[2] = 0.0;
  const double lam_rgb[3] = {680.0, 550.0, 440.0};
  double solar_rgb[3];
  for (int a = 0; a < 3; ++a) solar_rgb[a] = interpolate(wl, solar, lam_rgb[a]);
  for (int lambda = 360; lambda < 830; ++lambda) {
    const double xyz[3] = {cie_value(lambda, 1), cie_value(lambda, 2), cie_value(lambda, 3)};
    const double irradiance = interpolate(wl, solar, lambda);
    for (int a = 0; a < 3; ++a) {
      double bar = pas::kXyzToSrgb[a][0] * xyz[0] + pas::kXyzToSrgb[a][1] * xyz[1] +
                   pas::kXyzToSrgb[a][2] * xyz[2];
      k[a] += bar * irradiance / solar_rgb[a] * std::pow(lambda / lam_rgb[a], lambda_power);
    }
  }
  for (int a = 0; a < 3; ++a) k[a] *= pas::kMaxLuminousEfficacy;
}

// Wavelengths and luminance_from_radiance matrices of Model::Init (model.cc:907-949), all batches
// side by side: lum is row-major [3][C].
void spectral_channels(unsigned num_precomputed_wavelengths, std::vector<double>* lambdas,
                       std::vector<float>* lum) {
  if (num_precomputed_wavelengths <= 3) {
    *lambdas = {680.0, 550.0, 440.0};
    *lum = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    return;
  }
  const int iters = (int)(num_precomputed_wavelengths + 2) / 3;
  const int C = 3 * iters;
  const double dl = (pas::kCieLambdaMax - pas::kCieLambdaMin) / C;
  lambdas->resize(C);
  lum->assign(3 * (size_t)C, 0.f);
  for (int j = 0; j < C; ++j) {
    const double l = pas::kCieLambdaMin + (j + 0.5) * dl;
    (*lambdas)[j] = l;
    const double xyz[3] = {cie_value(l, 1), cie_value(l, 2), cie_value(l, 3)};
    for (int a = 0; a < 3; ++a) {
      // MAX_LUMINOUS_EFFICACY deliberately omitted here (model.cc:926-930)
      (*lum)[(size_t)a * C + j] = static_cast<float>(
          (pas::kXyzToSrgb[a][0] * xyz[0] + pas::kXyzToSrgb[a][1] * xyz[1] +
           pas::kXyzToSrgb[a][2] * xyz[2]) * dl);
    }
  }
}

// Device allocations are recycled through a process-wide pool keyed by (device, size): the demo
// re-creates its Model on every settings change (atmosphere/demo/demo.cc:446-494), and
// cudaMalloc/cudaFree of ~0.5 GB of tables would otherwise dominate the constructor.
struct BufferPool {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> free_list;
  size_t pooled = 0;           // bytes parked in the pool
  // PAS_POOL_MAX_MB (default 16 GiB): what the pool may keep; a buffer that does not fit is freed at once
  // (a process that walks through many table sizes would otherwise pin every size it has ever used)
  size_t cap = [] {
    const char* e = getenv("PAS_POOL_MAX_MB");
    const long long mb = e != nullptr ? atoll(e) : 0;
    return (size_t)(mb > 0 ? mb : 16384) << 20;
  }();
  void* take(int device, size_t bytes) {
    std::lock_guard<std::mutex> lock(mu);
    auto it = free_list.find({device, bytes});
    if (it == free_list.end()) return nullptr;
    void* p = it->second;
    free_list.erase(it);
    pooled -= bytes;
    return p;
  }
  void give(int device, size_t bytes, void* p) {
    {
      std::lock_guard<std::mutex> lock(mu);
      if (pooled + bytes <= cap) {
        free_list.insert({{device, bytes}, p});
        pooled += bytes;
        return;
      }
    }
    int current = 0;
    cudaGetDevice(&current);
    cudaSetDevice(device);
    cudaFree(p);   // synchronises with the device: the buffer's last user has finished
    cudaSetDevice(current);
  }
  void release_all() {
    std::lock_guard<std::mutex> lock(mu);
    int current = 0;
    cudaGetDevice(&current);
    for (auto& kv : free_list) {
      cudaSetDevice(kv.first.first);
      cudaFree(kv.second);
    }
    free_list.clear();
    pooled = 0;
    cudaSetDevice(current);
  }
};
// Streams and events are recycled the same way (creating and destroying two streams and two events
// per Model costs more than the pool lookup of all its buffers).
struct StreamSet {
  cudaStream_t stream = nullptr, aux = nullptr, copy = nullptr;
  cudaEvent_t ev_main = nullptr, ev_aux = nullptr, ev_copy = nullptr;
};
struct StreamPool {
  std::mutex mu;
  std::multimap<int, StreamSet> free_list;
  bool take(int device, StreamSet* out) {
    std::lock_guard<std::mutex> lock(mu);
    auto it = free_list.find(device);
    if (it == free_list.end()) return false;
    *out = it->second;
    free_list.erase(it);
    return true;
  }
  void give(int device, const StreamSet& s) {
    std::lock_guard<std::mutex> lock(mu);
    free_list.insert({device, s});
  }
};
StreamPool& stream_pool() {
  static StreamPool* p = new StreamPool();  // intentionally leaked, like the buffer pool
  return *p;
}

BufferPool& pool() {
  static BufferPool* p = new BufferPool();  // intentionally leaked: outlives the CUDA context teardown
  return *p;
}

struct DeviceBuffer {
  void* p = nullptr;
  size_t bytes = 0;
  int device = 0;
  bool borrowed = false;   // points into memory owned by somebody else (a symmetric arena)
  ~DeviceBuffer() { release(); }
  void release() {
    if (p && !borrowed) pool().give(device, bytes, p);
    p = nullptr;
    bytes = 0;
    borrowed = false;
  }
  void adopt(void* ptr, size_t n) {
    release();
    p = ptr;
    bytes = n;
    borrowed = true;
  }
  cudaError_t ensure(size_t n) {
    if (n <= bytes && !borrowed) return cudaSuccess;
    release();
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    p = pool().take(device, n);
    if (p == nullptr) {
      e = cudaMalloc(&p, n);
      if (e != cudaSuccess) {
        // make room and retry once
        pool().release_all();
        e = cudaMalloc(&p, n);
      }
    }
    if (e == cudaSuccess) bytes = n; else p = nullptr;
    return e;
  }
  float* f() const { return static_cast<float*>(p); }
};

// NCCL communicators are cached per (device, rank, world size) for the life of the process, so
// that re-creating a Model does not pay the communicator bootstrap again.
struct CommCache {
  std::mutex mu;
  std::map<std::tuple<int, int, int>, ncclComm_t> comms;
};
CommCache& comm_cache() {
  static CommCache* c = new CommCache();
  return *c;
}

// Peer-memory worlds: per (device, rank, world size) one flag array (the other ranks signal their
// barrier epochs into it; the epoch counters of the two barrier channels live beside the flags and
// are advanced by the barrier kernels, pas_kernels.h) and the error word of the barrier kernel; plus
// the IPC mappings already opened by this process (buffers come back from the pool with the same
// handles, so re-created models find their peers' tables mapped).
// The flag array is ONE barrier sequence per channel for the whole process: two models of the same
// world in flight at once would draw interleaved epochs and release each other's barriers early, so
// pas_model_init_async admits one model per PeerWorld at a time (`in_flight`).
struct PeerWorld {
  unsigned* flags = nullptr;   // device memory, PAS_FLAG_CHANNELS x PAS_FLAG_WORDS words, zero at creation
  int* error_host = nullptr;   // pinned, mapped
  int* error_dev = nullptr;
  const pas_model* in_flight = nullptr;  // the model whose Init is enqueued and not yet waited for
  bool broken = false;         // a barrier timed out: epochs and tables of the ranks are out of step
};
struct PeerCache {
  std::mutex mu;
  std::map<std::tuple<int, int, int>, PeerWorld> worlds;
  std::map<std::string, void*> opened;
};
PeerCache& peer_cache() {
  static PeerCache* c = new PeerCache();
  return *c;
}

// What a rank publishes to the others (pas_model_ipc_export).
struct PasIpcExport {
  cudaIpcMemHandle_t T, dJ[2], S, M, xE, flags;
  int has_M;
  int pad[3];
};
static_assert(sizeof(PasIpcExport) == PAS_IPC_EXPORT_BYTES, "PAS_IPC_EXPORT_BYTES");

}  // namespace

struct PhaseMarks {
  std::vector<std::pair<std::string, cudaEvent_t>> marks;
};

struct pas_model {
  // ---- what the constructor was given (SI units) ----
  std::vector<double> wavelengths, solar, rayleigh, mie_sca, mie_ext, absorption, albedo;
  std::vector<pas_density_layer> profile_layers[3];
  double sun_angular_radius = 0, bottom = 0, top = 0, mie_g = 0, max_sun_zenith = 0, unit = 1;
  unsigned num_precomputed_wavelengths = 3;
  bool combined = true, half = false;
  // ---- derived ----
  PasGeometry geom{};
  std::vector<double> lambdas;      // channel wavelengths (model.cc:907-924)
  std::vector<float> lum;           // [3][C] luminance_from_radiance (model.cc:909, 925-943)
  std::vector<PasSpectrum> groups;  // channel groups, <= PAS_MAX_CH each
  std::vector<int> group_offset;
  PasSpectrum rgb_spectrum{};       // 680/550/440 nm, for the final transmittance (model.cc:951-963)
  double sky_k[3] = {0, 0, 0}, sun_k[3] = {0, 0, 0};
  // ---- device state ----
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t aux = nullptr;       // side stream of Init: irradiance passes, final RGB transmittance
  cudaEvent_t ev_main = nullptr, ev_aux = nullptr;
  // host destinations registered with pas_model_set_host_outputs: Init copies every product table
  // out as soon as it is final, on a copy stream beside the passes still running
  cudaStream_t copy = nullptr;
  cudaEvent_t ev_copy = nullptr;
  void* host_out[4] = {nullptr, nullptr, nullptr, nullptr};  // indexed by pas_texture
  void* host_S_dev = nullptr;       // device-visible address of host_out[SCATTERING] when that is pinned memory
  void* host_S_now = nullptr;       // = host_S_dev while the pass that makes S final is being enqueued
  bool host_own_layers = false;     // multi-GPU: copy out only the layers this rank computed (shared host buffers)
  DeviceBuffer T, dE, dR, dM, dJ, dS, dirs, G, cR, cM, T_rgb, scratch;
  DeviceBuffer ray_setup;           // per-ray tables of the ray-march passes (pas::launch_ray_setup), or empty
  DeviceBuffer S, M, E, T_rgba;
  DeviceBuffer render_in[4], render_out[2];  // staging of host-pointer render queries
  float last_render_ms = 0.f;
  bool initialised = false;
  bool capture = false;
  std::map<std::string, std::unique_ptr<DeviceBuffer>> captured;
  std::vector<std::pair<std::string, float>> timings;
  PhaseMarks pending;               // events of an Init that has been enqueued but not waited for
  bool in_flight = false;
  bool fuse_rgb_transmittance = false;  // set by Init: the first transmittance launch also fills T_rgba
  int launches = 0;
  // ---- multi-GPU ----
  int rank = 0, world = 1;
  int pending_rank = 0, pending_world = 0;  // of pas_model_ipc_export, until pas_model_attach_peers succeeds
  ncclComm_t comm = nullptr;
  // peer-memory exchange (pas_model_attach_peers): tables of the other ranks, indexed by rank
  bool peer = false;
  PeerWorld* pw = nullptr;
  DeviceBuffer dJ2, xE;            // second density buffer, irradiance partials [2][world][n_e * nc]
  float* peer_T[PAS_MAX_PEERS + 1] = {};
  float* peer_dJ[2][PAS_MAX_PEERS + 1] = {};
  void* peer_S[PAS_MAX_PEERS + 1] = {};
  void* peer_M[PAS_MAX_PEERS + 1] = {};
  float* peer_xE[PAS_MAX_PEERS + 1] = {};
  unsigned* peer_flags[PAS_MAX_PEERS + 1] = {};
  unsigned exchanges = 0;          // orders exchanged so far: its parity selects dJ / dJ2 and the xE half
  // symmetric-arena worlds (pas_model_attach_symmetric): T, dJ, dJ2, xE live in the arena; the final
  // scattering slabs are exchanged through arena staging areas (the product tables stay the model's)
  bool symm = false;
  char* arena[PAS_MAX_PEERS + 1] = {};
  char* arena_mc = nullptr;        // multicast address of the arenas, or nullptr
  size_t off_T = 0, off_dJ[2] = {0, 0}, off_xE = 0, off_Sx = 0, off_Mx = 0, off_Tx = 0;
  // destinations of this rank's stores for an exchanged table at arena offset `off`: the one multicast
  // address when there is one (it includes the local copy), else the table in every other rank's arena
  pas::PeerTables mirrors_at(size_t off) const {
    pas::PeerTables t{};
    if (arena_mc != nullptr) {
      t.tab[t.n++] = reinterpret_cast<float*>(arena_mc + off);
      t.multicast = 1;
    } else {
      for (int r = 0; r < world; ++r) {
        if (r != rank) t.tab[t.n++] = reinterpret_cast<float*>(arena[r] + off);
      }
    }
    return t;
  }
  pas::PeerTargets targets_at(size_t off) const {
    pas::PeerTargets t{};
    if (arena_mc != nullptr) {
      t.dst[t.n++] = arena_mc + off;
    } else {
      for (int r = 0; r < world; ++r) {
        if (r != rank) t.dst[t.n++] = arena[r] + off;
      }
    }
    return t;
  }

  size_t n_t() const { return (size_t)geom.sz.t_w * geom.sz.t_h; }
  size_t n_e() const { return (size_t)geom.sz.e_w * geom.sz.e_h; }
  size_t n_s() const { return (size_t)geom.sz.nu_n * geom.sz.mu_s_n * geom.sz.mu_n * geom.sz.r_n; }
  size_t layer_texels() const { return (size_t)geom.sz.nu_n * geom.sz.mu_s_n * geom.sz.mu_n; }
  int total_channels() const { return (int)lambdas.size(); }
  size_t s_texel_bytes() const { return half ? 8 : 16; }
  int max_nc() const {
    int n = 0;
    for (const auto& g : groups) n = std::max(n, g.nc);
    return n;
  }
  size_t xe_stride() const { return n_e() * (size_t)max_nc(); }   // floats per rank slot
  float* cur_dJ() const { return (peer && (exchanges & 1)) ? dJ2.f() : dJ.f(); }
  // Scattering layers owned by this rank: all of them on one GPU, else a contiguous r-slab (the
  // first r_n % world ranks get one layer more). The kernels take any strided LayerSet; dealing the
  // layers round-robin (to even out the altitude-dependent cost of the density pass) was measured
  // slower than slabs on 4 and 8 B200 (2.34 vs 2.31 ms, 1.69 vs 1.54 ms).
  pas::LayerSet layers() const {
    if (world == 1) return pas::LayerSet{0, geom.sz.r_n, 1};
    // PAS_LAYER_DEAL=rr (peer / symmetric worlds): rank r takes layers r, r + world, ... -- the density
    // pass costs more at high altitude (wider ground windows), dealing the layers evens the ranks out
    static const bool deal = getenv("PAS_LAYER_DEAL") != nullptr && std::string(getenv("PAS_LAYER_DEAL")) == "rr";
    if (deal && peer) return pas::LayerSet{rank, geom.sz.r_n, world};
    const int base = geom.sz.r_n / world, extra = geom.sz.r_n % world;
    const int k_begin = rank * base + std::min(rank, extra);
    return pas::LayerSet{k_begin, k_begin + base + (rank < extra ? 1 : 0), 1};
  }
};

namespace {

PasSpectrum make_spectrum(const pas_model& m, const double* lambdas, int n, const float* lum3xC,
                          int lum_stride, int lum_offset) {
  PasSpectrum s{};
  s.nc = n;
  const double u = m.unit;
  for (int c = 0; c < n; ++c) {
    const double l = lambdas[c];
    // scales as in the GLSL header (model.cc:718-734): coefficients are per length unit
    s.solar[c] = interpolate(m.wavelengths, m.solar, l);
    s.beta_r[c] = interpolate(m.wavelengths, m.rayleigh, l) * u;
    s.beta_m_sca[c] = interpolate(m.wavelengths, m.mie_sca, l) * u;
    s.beta_m_ext[c] = interpolate(m.wavelengths, m.mie_ext, l) * u;
    s.beta_abs[c] = interpolate(m.wavelengths, m.absorption, l) * u;
    s.albedo[c] = interpolate(m.wavelengths, m.albedo, l);
    for (int a = 0; a < 3; ++a) {
      s.lum[a][c] = lum3xC ? lum3xC[a * lum_stride + lum_offset + c] : (a == c ? 1.f : 0.f);
    }
  }
  return s;
}

pas_status validate(const pas_model_params* p) {
  if (p == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "params is NULL");
  if (p->num_wavelengths < 1 || !p->wavelengths || !p->solar_irradiance || !p->rayleigh_scattering ||
      !p->mie_scattering || !p->mie_extinction || !p->absorption_extinction || !p->ground_albedo) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "missing spectrum arrays");
  }
  for (size_t i = 0; i + 1 < p->num_wavelengths; ++i) {
    if (!(p->wavelengths[i] < p->wavelengths[i + 1])) {
      return fail(PAS_ERR_INVALID_ARGUMENT, "wavelengths must be strictly increasing");
    }
  }
  if (!(p->bottom_radius > 0.0) || !(p->top_radius > p->bottom_radius)) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "need 0 < bottom_radius < top_radius");
  }
  if (!(p->length_unit_in_meters > 0.0)) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "length_unit_in_meters must be positive");
  }
  if (!(p->sun_angular_radius > 0.0) || p->sun_angular_radius >= 0.1) {
    // documented validity limit of the sun-disc approximations (atmosphere/model.h:194-195)
    return fail(PAS_ERR_INVALID_ARGUMENT, "sun_angular_radius must be in (0, 0.1) rad");
  }
  if (p->num_rayleigh_layers > 2 || p->num_mie_layers > 2 || p->num_absorption_layers > 2) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "at most 2 density layers per profile");
  }
  if ((p->num_rayleigh_layers && !p->rayleigh_density) || (p->num_mie_layers && !p->mie_density) ||
      (p->num_absorption_layers && !p->absorption_density)) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "missing density layer arrays");
  }
  if (!(std::fabs(p->mie_phase_function_g) < 1.0)) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "|mie_phase_function_g| must be < 1");
  }
  if (p->num_precomputed_wavelengths < 1 || p->num_precomputed_wavelengths > 240) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "num_precomputed_wavelengths must be in [1, 240]");
  }
  return PAS_OK;
}

// Splits the channels (always a multiple of 3, model.cc:917) into the fewest launch groups of a
// size the kernels are instantiated for.
void split_channels(int total, std::vector<int>* sizes) {
  static const int kSupported[] = {16, 15, 8, 4, 3};
  std::vector<int> best(total + 1, 1 << 20), pick(total + 1, 0);
  best[0] = 0;
  for (int n = 1; n <= total; ++n) {
    for (int s : kSupported) {
      if (s <= n && pas::channel_count_supported(s) && best[n - s] + 1 < best[n]) {
        best[n] = best[n - s] + 1;
        pick[n] = s;
      }
    }
  }
  for (int n = total; n > 0 && pick[n] > 0; n -= pick[n]) sizes->push_back(pick[n]);
}

struct nullptr_model_tag {};
struct PhaseTimer {
  pas_model* m;
  explicit PhaseTimer(pas_model* model) : m(model) {
    for (auto& mk : m->pending.marks) cudaEventDestroy(mk.second);
    m->pending.marks.clear();
  }
  PhaseTimer(nullptr_model_tag, pas_model* model) : m(model) {}  // picks up the marks of an Init in flight
  void mark(const std::string& name) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, m->stream);
    m->pending.marks.emplace_back(name, e);
  }
  void finish() {
    auto& marks = m->pending.marks;
    m->timings.clear();
    for (size_t i = 1; i < marks.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
      m->timings.emplace_back(marks[i].first, ms);
    }
    for (auto& mk : marks) cudaEventDestroy(mk.second);
    marks.clear();
  }
};

// keeps a planar copy [c][texel] of an intermediate (all channels of all groups) when capture is on
pas_status capture_copy(pas_model* m, const std::string& name, const float* src, size_t texels,
                        int nc, int channel_offset, bool interleaved) {
  if (!m->capture) return PAS_OK;
  auto& slot = m->captured[name];
  if (!slot) slot.reset(new DeviceBuffer());
  PAS_CUDA(slot->ensure(texels * m->total_channels() * sizeof(float)));
  float* dst = slot->f() + (size_t)channel_offset * texels;
  if (interleaved) {
    PAS_CUDA(pas::launch_interleaved_to_planar(src, texels, nc, dst, m->stream));
  } else {
    PAS_CUDA(cudaMemcpyAsync(dst, src, texels * nc * sizeof(float), cudaMemcpyDeviceToDevice, m->stream));
  }
  return PAS_OK;
}

__global__ void accumulate_irradiance_kernel(const float* __restrict__ dE, int n, int nc,
                                             PasSpectrum sp, float* __restrict__ E) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float4 e = reinterpret_cast<float4*>(E)[t];
  for (int c = 0; c < nc; ++c) {
    const float v = dE[(size_t)c * n + t];
    e.x = fmaf(sp.lum[0][c], v, e.x);
    e.y = fmaf(sp.lum[1][c], v, e.y);
    e.z = fmaf(sp.lum[2][c], v, e.z);
  }
  reinterpret_cast<float4*>(E)[t] = e;
}

__global__ void half_to_float_kernel(const __half* __restrict__ src, size_t n, float* __restrict__ dst) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[t] = __half2float(src[t]);
}

pas_status texture_lookup(const pas_model* m, pas_texture which, const DeviceBuffer** buf,
                          pas_texture_info* info) {
  pas_texture_info i{};
  i.channels = 4;
  i.present = 1;
  const PasSizes& z = m->geom.sz;
  switch (which) {
    case PAS_TEXTURE_TRANSMITTANCE:
      i.width = z.t_w; i.height = z.t_h; i.depth = 1; i.bytes_per_channel = 4;
      if (buf) *buf = &m->T_rgba;
      break;
    case PAS_TEXTURE_IRRADIANCE:
      i.width = z.e_w; i.height = z.e_h; i.depth = 1; i.bytes_per_channel = 4;
      if (buf) *buf = &m->E;
      break;
    case PAS_TEXTURE_SCATTERING:
      i.width = z.nu_n * z.mu_s_n; i.height = z.mu_n; i.depth = z.r_n;
      i.bytes_per_channel = m->half ? 2 : 4;
      if (buf) *buf = &m->S;
      break;
    case PAS_TEXTURE_SINGLE_MIE:
      i.width = z.nu_n * z.mu_s_n; i.height = z.mu_n; i.depth = z.r_n;
      i.bytes_per_channel = m->half ? 2 : 4;
      i.present = m->combined ? 0 : 1;
      if (buf) *buf = &m->M;
      break;
    default:
      return fail(PAS_ERR_INVALID_ARGUMENT, "unknown texture id");
  }
  if (info) *info = i;
  return PAS_OK;
}

// Live intermediate buffers by name (teacher-forced test hooks).
bool live_buffer(pas_model* m, const std::string& name, float** p, size_t* texels, bool* interleaved) {
  *interleaved = true;
  if (name == "transmittance") { *p = m->T.f(); *texels = m->n_t(); return true; }
  if (name == "delta_irradiance") { *p = m->dE.f(); *texels = m->n_e(); *interleaved = false; return true; }
  if (name == "delta_rayleigh") { *p = m->dR.f(); *texels = m->n_s(); return true; }
  if (name == "delta_mie") { *p = m->dM.f(); *texels = m->n_s(); return true; }
  if (name == "delta_density") { *p = m->dJ.f(); *texels = m->n_s(); return true; }
  if (name == "delta_multiple") { *p = m->dS.f(); *texels = m->n_s(); return true; }
  return false;
}

pas::FinalTables final_tables(pas_model* m, bool accumulate) {
  pas::FinalTables f{};
  f.scattering = m->S.p;
  f.single_mie = m->combined ? nullptr : m->M.p;
  f.irradiance = m->E.f();
  f.half_precision = m->half ? 1 : 0;
  f.accumulate = accumulate ? 1 : 0;
  f.host_scattering = m->host_S_now;
  return f;
}

pas_status allocate(pas_model* m) {
  int max_nc = 0;
  for (const auto& g : m->groups) max_nc = std::max(max_nc, g.nc);
  const PasSizes& z = m->geom.sz;
  const size_t cp = PAS_CHANNEL_PITCH(max_nc);  // interleaved tables: cp floats per texel
  PAS_CUDA(m->T.ensure(m->n_t() * cp * sizeof(float)));
  PAS_CUDA(m->T_rgb.ensure(m->n_t() * PAS_CHANNEL_PITCH(3) * sizeof(float)));
  PAS_CUDA(m->dE.ensure(m->n_e() * max_nc * sizeof(float)));
  PAS_CUDA(m->dR.ensure(m->n_s() * cp * sizeof(float)));
  PAS_CUDA(m->dM.ensure(m->n_s() * cp * sizeof(float)));
  PAS_CUDA(m->dJ.ensure(m->n_s() * cp * sizeof(float)));
  PAS_CUDA(m->dS.ensure(m->n_s() * cp * sizeof(float)));
  PAS_CUDA(m->dirs.ensure((size_t)z.r_n * PAS_DIR_THETA * sizeof(PasDensityDir)));
  PAS_CUDA(m->G.ensure((size_t)z.r_n * PAS_DIR_THETA * PAS_MAX_CH * sizeof(float)));
  PAS_CUDA(m->cR.ensure((size_t)z.r_n * PAS_MAX_CH * sizeof(float)));
  PAS_CUDA(m->cM.ensure((size_t)z.r_n * PAS_MAX_CH * sizeof(float)));
  if (pas::ray_setup_bytes(m->geom, max_nc) > 0) PAS_CUDA(m->ray_setup.ensure(pas::ray_setup_bytes(m->geom, max_nc)));
  PAS_CUDA(m->S.ensure(m->n_s() * m->s_texel_bytes()));
  if (!m->combined) PAS_CUDA(m->M.ensure(m->n_s() * m->s_texel_bytes()));
  PAS_CUDA(m->E.ensure(m->n_e() * 16));
  PAS_CUDA(m->T_rgba.ensure(m->n_t() * 16));
  return PAS_OK;
}

// A model leaving a symmetric arena (re-attached to another kind of world) gets buffers of its own back.
pas_status leave_arena(pas_model* m) {
  m->symm = false;
  m->arena_mc = nullptr;
  m->T.release();
  m->dJ.release();
  m->dJ2.release();
  m->xE.release();
  return allocate(m);
}

// One phase of Precompute (model.cc:1048-1215) for channel group `gi`.
// Cross-GPU barrier on `stream`; channel 0 belongs to the main stream of Init, 1 to the side stream.
// Entry points that read results first wait for an Init still in flight (pas_model_init_async).
pas_status settle(const pas_model* m) {
  return (m != nullptr && m->in_flight) ? pas_model_wait(const_cast<pas_model*>(m)) : PAS_OK;
}

pas_status peer_barrier(pas_model* m, int channel, cudaStream_t stream) {
  pas::PeerFlags f{};
  f.rank = m->rank;
  f.world = m->world;
  for (int r = 0; r < m->world; ++r) {
    f.flags[r] = m->peer_flags[r] + channel * PAS_FLAG_WORDS;
    f.poison[r] = m->peer_flags[r] + PAS_FLAG_POISON;
  }
  f.error = m->pw->error_dev;
  PAS_CUDA(pas::launch_peer_barrier(f, stream));
  m->launches += 1;
  return PAS_OK;
}

// `stream`: where the phase is enqueued. `ds_in`: the table holding the previous order's multiple
// scattering (read by the density and irradiance passes); `ds_out`: where the multiple-scattering
// pass writes. The reference aliases both to one texture (model.cc:897); Init alternates two buffers
// so that the irradiance pass of order n can overlap the multiple-scattering pass of order n.
pas_status run_phase(pas_model* m, int gi, int phase, int order, bool accumulate, cudaStream_t stream,
                     float* ds_in, float* ds_out, int channel = 0, const pas::LayerSet* only = nullptr) {
  const PasSpectrum& sp = m->groups[gi];
  const PasGeometry& g = m->geom;
  const pas::LayerSet ks = only != nullptr ? *only : m->layers();
  const int k0 = ks.begin;
  pas::FinalTables fin = final_tables(m, accumulate);
  switch (phase) {
    case 0:
      if (m->peer) {
        // a band of transmittance rows per rank, stored to every rank; then all meet
        pas::PeerTables mirrors{};
        if (m->symm) {
          mirrors = m->mirrors_at(m->off_T);
        } else {
          for (int r = 0; r < m->world; ++r) {
            if (r != m->rank) mirrors.tab[mirrors.n++] = m->peer_T[r];
          }
        }
        const int base = g.sz.t_h / m->world, extra = g.sz.t_h % m->world;
        const int j0 = m->rank * base + std::min(m->rank, extra);
        const int j1 = j0 + base + (m->rank < extra ? 1 : 0);
        if (gi > 0) {
          // the rows go straight into the other ranks' T, which their ray marches of the previous
          // channel group may still be reading (the last barrier of a group comes before its last
          // multiple-scattering pass, and Init(1) has none after phase 0): all ranks meet first
          pas_status st = peer_barrier(m, channel, stream);
          if (st != PAS_OK) return st;
        }
        if (m->fuse_rgb_transmittance && gi == 0) {
          // symmetric worlds: the same launch fills this rank's rows of the final RGBA table, in the
          // staging area of every rank's arena; copied into the model's own table after the barrier
          const pas::PeerTables rgba_mirrors = m->mirrors_at(m->off_Tx);
          PAS_CUDA(pas::launch_transmittance_rows(g, sp, m->T.f(), mirrors, j0, j1, stream, &m->rgb_spectrum,
                                                  reinterpret_cast<float*>(m->arena[m->rank] + m->off_Tx),
                                                  &rgba_mirrors));
        } else {
          PAS_CUDA(pas::launch_transmittance_rows(g, sp, m->T.f(), mirrors, j0, j1, stream));
        }
        pas_status st = peer_barrier(m, channel, stream);
        if (st != PAS_OK) return st;
        if (m->fuse_rgb_transmittance && gi == 0) {
          PAS_CUDA(cudaMemcpyAsync(m->T_rgba.p, m->arena[m->rank] + m->off_Tx, m->n_t() * 16,
                                   cudaMemcpyDeviceToDevice, stream));
        }
      } else if (m->fuse_rgb_transmittance && gi == 0) {
        PAS_CUDA(pas::launch_transmittance(g, sp, m->T.f(), stream, &m->rgb_spectrum, m->T_rgba.f()));
      } else {
        PAS_CUDA(pas::launch_transmittance(g, sp, m->T.f(), stream));
      }
      PAS_CUDA(pas::launch_density_setup(g, sp, m->T.f(), static_cast<PasDensityDir*>(m->dirs.p),
                                         m->G.f(), m->cR.f(), m->cM.f(), stream));
      m->launches += 2;
      if (m->ray_setup.p != nullptr) {
        // sample records / path transmittances / permutations of this rank's rays, once for the single-
        // scattering pass and all the multiple-scattering passes of this channel group
        PAS_CUDA(pas::launch_ray_setup(g, sp, m->T.f(), m->ray_setup.p, ks, stream));
        m->launches += 1;
      }
      break;
    case 1:
      PAS_CUDA(pas::launch_direct_irradiance(g, sp, m->T.f(), m->dE.f(), fin, stream));
      m->launches += 1;
      break;
    case 2:
      PAS_CUDA(pas::launch_single_scattering(g, sp, m->T.f(), m->dR.f(), m->dM.f(), fin, ks, stream, m->ray_setup.p));
      m->launches += 1;
      break;
    case 3:
    {
      // peer mode: the kernel stores its slab to every rank (mirrors), into the density buffer of
      // this order's parity; the barrier comes with the irradiance partial sums (phase 4)
      pas::PeerTables mirrors{};
      const int par = (int)(m->exchanges & 1);
      if (m->symm) {
        mirrors = m->mirrors_at(m->off_dJ[par]);
      } else if (m->peer) {
        for (int r = 0; r < m->world; ++r) {
          if (r != m->rank) mirrors.tab[mirrors.n++] = m->peer_dJ[par][r];
        }
      }
      PAS_CUDA(pas::launch_scattering_density(
          g, sp, static_cast<const PasDensityDir*>(m->dirs.p), m->G.f(), m->cR.f(), m->cM.f(),
          m->dR.f(), m->dM.f(), ds_in, m->dE.f(), order, m->cur_dJ(), mirrors, ks, stream));
      m->launches += 1;
      if (m->world > 1 && !m->peer) {
        // all-gather of the density r-slabs over NVLink: every rank needs every layer its rays
        // cross in the multiple-scattering pass
        // interleaved layout: one layer holds every channel, so each rank's slab is one
        // contiguous run and the whole exchange is a single all-gather
        const size_t lt = m->layer_texels() * PAS_CHANNEL_PITCH(sp.nc);
        const int base = g.sz.r_n / m->world;
        if (g.sz.r_n % m->world != 0) {
          return fail(PAS_ERR_UNSUPPORTED, "scattering_r must be divisible by the world size");
        }
        PAS_NCCL(nccl().AllGather(m->dJ.f() + (size_t)k0 * lt, m->dJ.f(), (size_t)base * lt, ncclFloat,
                                  m->comm, stream));
      }
      break;
    }
    case 4: {
      if (m->world > 1) fin.irradiance = nullptr;  // partial sums: accumulate after the all-reduce
      if (m->peer) {
        // partial sums of this rank's layers go to slot [parity][rank] of every rank's xE; the
        // barrier that follows also completes the density slabs stored by phase 3
        const int par = (int)(m->exchanges & 1);
        const size_t stride = m->xe_stride();
        const size_t slot = ((size_t)par * m->world + m->rank) * stride;
        PAS_CUDA(pas::launch_indirect_irradiance(g, sp, m->dR.f(), m->dM.f(), ds_in, order,
                                                 m->xE.f() + slot, fin, 0, g.sz.e_h, ks, stream));
        pas::PeerTargets t{};
        if (m->symm) {
          t = m->targets_at(m->off_xE);
        } else {
          for (int r = 0; r < m->world; ++r) {
            if (r != m->rank) t.dst[t.n++] = m->peer_xE[r];
          }
        }
        PAS_CUDA(pas::launch_peer_push(m->xE.f(), m->n_e() * sp.nc * sizeof(float), slot * sizeof(float), t, stream));
        if (channel != 0) {
          // the density slabs stored by phase 3 are completed by a barrier of the main stream
          pas_status st = peer_barrier(m, 0, m->stream);
          if (st != PAS_OK) return st;
        }
        pas_status st = peer_barrier(m, channel, stream);
        if (st != PAS_OK) return st;
        PAS_CUDA(pas::launch_sum_partials(m->xE.f() + (size_t)par * m->world * stride, m->world, stride,
                                          (int)(m->n_e() * sp.nc), m->dE.f(), stream));
        const int n = (int)m->n_e();
        accumulate_irradiance_kernel<<<(n + 127) / 128, 128, 0, stream>>>(m->dE.f(), n, sp.nc, sp, m->E.f());
        PAS_CUDA(cudaGetLastError());
        m->launches += 4;
        break;
      }
      PAS_CUDA(pas::launch_indirect_irradiance(g, sp, m->dR.f(), m->dM.f(), ds_in, order,
                                               m->dE.f(), fin, 0, g.sz.e_h, ks, stream));
      m->launches += 1;
      if (m->world > 1) {
        PAS_NCCL(nccl().AllReduce(m->dE.f(), m->dE.f(), m->n_e() * sp.nc, ncclFloat, ncclSum,
                                  m->comm, stream));
        const int n = (int)m->n_e();
        accumulate_irradiance_kernel<<<(n + 127) / 128, 128, 0, stream>>>(m->dE.f(), n, sp.nc, sp,
                                                                              m->E.f());
        PAS_CUDA(cudaGetLastError());
        m->launches += 1;
      }
      break;
    }
    case 5:
      PAS_CUDA(pas::launch_multiple_scattering(g, sp, m->T.f(), m->cur_dJ(), ds_out, fin, ks, stream, m->ray_setup.p));
      m->launches += 1;
      break;
    case 6:
      // test hook: the per-(layer, direction) tables of the density pass alone, from whatever the
      // transmittance buffer holds (device-side known-answer tests write their own transmittance)
      PAS_CUDA(pas::launch_density_setup(g, sp, m->T.f(), static_cast<PasDensityDir*>(m->dirs.p),
                                         m->G.f(), m->cR.f(), m->cM.f(), stream));
      m->launches += 1;
      break;
    default:
      return fail(PAS_ERR_INVALID_ARGUMENT, "unknown phase");
  }
  return PAS_OK;
}

}  // namespace

extern "C" {

const char* pas_last_error(void) { return g_last_error.c_str(); }
int pas_abi_version(void) { return PAS_B200_ABI_VERSION; }

}  // extern "C"

namespace {
// The host half of atmosphere::Model::Model (model.cc:613-690): parameters, units, channel grid,
// luminance matrices and factors. Needs no device: the GLSL source of a model is a function of this alone.
pas_status build_host_model(const pas_model_params* p, std::unique_ptr<pas_model>* out) {
  pas_status st = validate(p);
  if (st != PAS_OK) return st;
  std::unique_ptr<pas_model> m(new pas_model());
  const size_t n = p->num_wavelengths;
  m->wavelengths.assign(p->wavelengths, p->wavelengths + n);
  m->solar.assign(p->solar_irradiance, p->solar_irradiance + n);
  m->rayleigh.assign(p->rayleigh_scattering, p->rayleigh_scattering + n);
  m->mie_sca.assign(p->mie_scattering, p->mie_scattering + n);
  m->mie_ext.assign(p->mie_extinction, p->mie_extinction + n);
  m->absorption.assign(p->absorption_extinction, p->absorption_extinction + n);
  m->albedo.assign(p->ground_albedo, p->ground_albedo + n);
  m->profile_layers[0].assign(p->rayleigh_density, p->rayleigh_density + p->num_rayleigh_layers);
  m->profile_layers[1].assign(p->mie_density, p->mie_density + p->num_mie_layers);
  m->profile_layers[2].assign(p->absorption_density, p->absorption_density + p->num_absorption_layers);
  m->sun_angular_radius = p->sun_angular_radius;
  m->bottom = p->bottom_radius;
  m->top = p->top_radius;
  m->mie_g = p->mie_phase_function_g;
  m->max_sun_zenith = p->max_sun_zenith_angle;
  m->unit = p->length_unit_in_meters;
  m->num_precomputed_wavelengths = p->num_precomputed_wavelengths;
  m->combined = p->combine_scattering_textures != 0;
  m->half = p->half_precision != 0;

  // ---- geometry block, in length units (model.cc:718-734) ----
  PasGeometry& g = m->geom;
  auto pick = [](int v, int dflt) { return v > 0 ? v : dflt; };
  g.sz.t_w = pick(p->sizes.transmittance_width, 256);
  g.sz.t_h = pick(p->sizes.transmittance_height, 64);
  g.sz.r_n = pick(p->sizes.scattering_r, 32);
  g.sz.mu_n = pick(p->sizes.scattering_mu, 128);
  g.sz.mu_s_n = pick(p->sizes.scattering_mu_s, 32);
  g.sz.nu_n = pick(p->sizes.scattering_nu, 8);
  g.sz.e_w = pick(p->sizes.irradiance_width, 64);
  g.sz.e_h = pick(p->sizes.irradiance_height, 16);
  if (g.sz.t_w < 2 || g.sz.t_h < 2 || g.sz.r_n < 2 || g.sz.mu_n < 4 || (g.sz.mu_n & 1) ||
      g.sz.mu_s_n < 2 || g.sz.nu_n < 2 || g.sz.e_w < 2 || g.sz.e_h < 2) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "table sizes must be >= 2 (scattering_mu even, >= 4)");
  }
  if (g.sz.nu_n > PAS_MAX_NU || g.sz.nu_n * g.sz.mu_s_n > 1024 || g.sz.e_w > 1024) {
    return fail(PAS_ERR_UNSUPPORTED,
                "supported: scattering_nu <= 16, scattering_nu * scattering_mu_s <= 1024, "
                "irradiance_width <= 1024");
  }
  g.bottom = m->bottom / m->unit;
  g.top = m->top / m->unit;
  g.H = std::sqrt(g.top * g.top - g.bottom * g.bottom);
  g.mu_s_min = std::cos(m->max_sun_zenith);
  g.sun_angular_radius = m->sun_angular_radius;
  g.mie_g = m->mie_g;
  {
    // "A" of the mu_s mapping (functions.glsl:819-821)
    const double d_min = g.top - g.bottom, d_max = g.H;
    const double D = pas::dist_top(g, g.bottom, g.mu_s_min);
    g.mus_A = (D - d_min) / (d_max - d_min);
  }
  for (int pr = 0; pr < 3; ++pr) {
    // missing layers are zero layers inserted at the front (model.cc:653-666); lengths are divided,
    // inverse lengths multiplied by the length unit (model.cc:641-650)
    std::vector<pas_density_layer> ls = m->profile_layers[pr];
    while (ls.size() < 2) ls.insert(ls.begin(), pas_density_layer{0, 0, 0, 0, 0});
    for (int l = 0; l < 2; ++l) {
      g.profiles[pr][l][0] = ls[l].width / m->unit;
      g.profiles[pr][l][1] = ls[l].exp_term;
      g.profiles[pr][l][2] = ls[l].exp_scale * m->unit;
      g.profiles[pr][l][3] = ls[l].linear_term * m->unit;
      g.profiles[pr][l][4] = ls[l].constant_term;
    }
  }

  // ---- channels and luminance matrices (model.cc:907-949) ----
  const double lam_rgb[3] = {680.0, 550.0, 440.0};
  spectral_channels(m->num_precomputed_wavelengths, &m->lambdas, &m->lum);
  const int C = m->total_channels();
  std::vector<int> sizes;
  split_channels(C, &sizes);
  int off = 0;
  for (int s : sizes) {
    m->groups.push_back(make_spectrum(*m, m->lambdas.data() + off, s, m->lum.data(), C, off));
    m->group_offset.push_back(off);
    off += s;
  }
  m->rgb_spectrum = make_spectrum(*m, lam_rgb, 3, nullptr, 0, 0);
  // SKY / SUN_SPECTRAL_RADIANCE_TO_LUMINANCE (model.cc:668-686)
  if (m->num_precomputed_wavelengths > 3) {
    m->sky_k[0] = m->sky_k[1] = m->sky_k[2] = pas::kMaxLuminousEfficacy;
  } else {
    luminance_factors(m->wavelengths, m->solar, -3.0, m->sky_k);
  }
  luminance_factors(m->wavelengths, m->solar, 0.0, m->sun_k);
  *out = std::move(m);
  return PAS_OK;
}
}  // namespace

extern "C" {

pas_status pas_model_create(const pas_model_params* p, pas_model** out) {
  if (out == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "out_model is NULL");
  *out = nullptr;
  std::unique_ptr<pas_model> m;
  pas_status st = build_host_model(p, &m);
  if (st != PAS_OK) return st;

  // ---- device ----
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    return fail(PAS_ERR_CUDA, std::string("no CUDA device available: ") +
                                  (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
  }
  if (p->device > 0) {
    m->device = p->device - 1;
    PAS_CUDA(cudaSetDevice(m->device));
  } else {
    PAS_CUDA(cudaGetDevice(&m->device));
  }
  {
    StreamSet ss;
    if (!stream_pool().take(m->device, &ss)) {
      // the side stream gets the higher priority: its small kernels (1024 blocks) are scheduled as
      // soon as the big pass running beside them frees a slot
      int lo = 0, hi = 0;
      PAS_CUDA(cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking));
      PAS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      PAS_CUDA(cudaStreamCreateWithPriority(&ss.aux, cudaStreamNonBlocking, hi));
      PAS_CUDA(cudaEventCreateWithFlags(&ss.ev_main, cudaEventDisableTiming));
      PAS_CUDA(cudaEventCreateWithFlags(&ss.ev_aux, cudaEventDisableTiming));
      PAS_CUDA(cudaStreamCreateWithFlags(&ss.copy, cudaStreamNonBlocking));
      PAS_CUDA(cudaEventCreateWithFlags(&ss.ev_copy, cudaEventDisableTiming));
    }
    m->stream = ss.stream;
    m->copy = ss.copy;
    m->ev_copy = ss.ev_copy;
    m->aux = ss.aux;
    m->ev_main = ss.ev_main;
    m->ev_aux = ss.ev_aux;
  }
  st = allocate(m.get());
  if (st != PAS_OK) {
    pas_model_destroy(m.release());  // returns the streams and whatever was allocated to the pools
    return st;
  }
  *out = m.release();
  return PAS_OK;
}

void pas_model_destroy(pas_model* m) {
  if (m == nullptr) return;
  cudaSetDevice(m->device);
  for (auto& mk : m->pending.marks) cudaEventDestroy(mk.second);
  if (m->stream) {
    // everything enqueued by this model has to finish before its buffers go back to the pool
    cudaStreamSynchronize(m->stream);
    if (m->aux) cudaStreamSynchronize(m->aux);
    if (m->copy) cudaStreamSynchronize(m->copy);
    if (m->pw != nullptr) {
      PeerCache& cache = peer_cache();
      std::lock_guard<std::mutex> lock(cache.mu);
      if (m->pw->in_flight == m) m->pw->in_flight = nullptr;
    }
    StreamSet ss;
    ss.copy = m->copy;
    ss.ev_copy = m->ev_copy;
    ss.stream = m->stream;
    ss.aux = m->aux;
    ss.ev_main = m->ev_main;
    ss.ev_aux = m->ev_aux;
    stream_pool().give(m->device, ss);
  }
  delete m;
}

pas_status pas_model_init(pas_model* m, unsigned int num_scattering_orders) {
  pas_status st = pas_model_init_async(m, num_scattering_orders);
  return st != PAS_OK ? st : pas_model_wait(m);
}

pas_status pas_model_set_host_outputs(pas_model* m, void* transmittance, void* scattering, void* single_mie,
                                      void* irradiance) {
  if (m == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "model is NULL");
  { pas_status settled = settle(m); if (settled != PAS_OK) return settled; }
  if (single_mie != nullptr && m->combined) {
    return fail(PAS_ERR_STATE, "this model has no single-Mie table (combined scattering textures)");
  }
  m->host_out[PAS_TEXTURE_TRANSMITTANCE] = transmittance;
  m->host_out[PAS_TEXTURE_SCATTERING] = scattering;
  m->host_out[PAS_TEXTURE_SINGLE_MIE] = single_mie;
  m->host_out[PAS_TEXTURE_IRRADIANCE] = irradiance;
  // Pinned (cudaHostAlloc / cudaHostRegister) destinations are visible to the device: the last
  // multiple-scattering pass then writes the scattering table there itself. Pageable destinations are
  // filled by copies behind the pass.
  m->host_S_dev = nullptr;
  static const bool no_fused = getenv("PAS_NO_FUSED_READBACK") != nullptr;
  if (scattering != nullptr && !no_fused) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, scattering) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
      m->host_S_dev = attr.devicePointer;
    }
    cudaGetLastError();
  }
  return PAS_OK;
}

pas_status pas_model_set_host_output_mode(pas_model* m, int own_layers_only) {
  if (m == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "model is NULL");
  { pas_status settled = settle(m); if (settled != PAS_OK) return settled; }
  m->host_own_layers = own_layers_only != 0;
  return PAS_OK;
}

pas_status pas_model_wait(pas_model* m) {
  if (m == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "model is NULL");
  if (!m->in_flight) return PAS_OK;
  PAS_CUDA(cudaSetDevice(m->device));
  m->in_flight = false;
  PAS_CUDA(cudaStreamSynchronize(m->stream));
  PhaseTimer(nullptr_model_tag(), m).finish();
  if (m->peer) {
    PeerCache& cache = peer_cache();
    std::lock_guard<std::mutex> lock(cache.mu);
    if (m->pw->in_flight == m) m->pw->in_flight = nullptr;
    const int code = *m->pw->error_host;
    if (code != 0) {
      // the barrier epochs and the tables of the ranks are out of step from here on: the peer world
      // of this (device, rank, world) stays unusable (attach the models to an NCCL world instead)
      *m->pw->error_host = 0;
      m->pw->broken = true;
      return fail(PAS_ERR_NCCL, code == 1 ? "peer barrier timed out: another rank failed or is out of step "
                                            "(PAS_PEER_TIMEOUT_MS sets the wait)"
                                          : "another rank's peer barrier timed out: this Init used stale tables");
    }
  }
  m->initialised = true;
  return PAS_OK;
}

pas_status pas_model_init_async(pas_model* m, unsigned int num_scattering_orders) {
  if (m == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "model is NULL");
  if (num_scattering_orders < 1) return fail(PAS_ERR_INVALID_ARGUMENT, "need >= 1 scattering order");
  PAS_CUDA(cudaSetDevice(m->device));
  if (m->in_flight) {
    pas_status st = pas_model_wait(m);
    if (st != PAS_OK) return st;
  }
  if (m->peer) {
    PeerCache& cache = peer_cache();
    std::lock_guard<std::mutex> lock(cache.mu);
    if (m->pw->broken) {
      return fail(PAS_ERR_NCCL, "the peer world of this rank is out of step after a barrier timeout");
    }
    if (m->pw->in_flight != nullptr && m->pw->in_flight != m) {
      return fail(PAS_ERR_STATE, "another model of this peer world has an Init in flight: wait for it first "
                                 "(the ranks share one barrier sequence per world)");
    }
    m->pw->in_flight = m;
  }
  m->launches = 0;
  // Overlapped schedule (single GPU, no captures): the irradiance pass of order n depends on the
  // radiance of order n - 1 only, so it runs on the side stream beside the multiple-scattering pass
  // of order n; so does the final RGB transmittance, beside everything. With captures (tests) or in
  // a multi-GPU world every pass is enqueued on the one stream, in the reference's order.
  static const bool no_overlap = getenv("PAS_NO_OVERLAP") != nullptr;
  const bool overlap = !m->capture && (m->world == 1 || m->peer) && !no_overlap;
  cudaStream_t main = m->stream, side = overlap ? m->aux : m->stream;
  auto side_after_main = [&]() -> cudaError_t {
    if (!overlap) return cudaSuccess;
    cudaError_t e = cudaEventRecord(m->ev_main, main);
    return e != cudaSuccess ? e : cudaStreamWaitEvent(side, m->ev_main, 0);
  };
  auto main_after_side = [&]() -> cudaError_t {
    if (!overlap) return cudaSuccess;
    cudaError_t e = cudaEventRecord(m->ev_aux, side);
    return e != cudaSuccess ? e : cudaStreamWaitEvent(main, m->ev_aux, 0);
  };
  // Pipelined read-back (pas_model_set_host_outputs; one GPU, no captures): a table is copied to the
  // host as soon as its last writer is done -- T after the first pass, the single-Mie table after
  // single scattering, E after the last irradiance pass. S is written to the host by the last
  // multiple-scattering pass itself when the destination is pinned memory (FinalTables::host_scattering);
  // a pageable destination is filled in four bands of layers behind the four launches that pass is
  // then split into.
  // Peer worlds with pas_model_set_host_output_mode(own layers only): every rank copies the layers of
  // the 3-D tables IT computed (rank 0 also T and E) into host buffers the ranks share, and the last
  // barrier of Init comes after those copies: when Init returns on any rank, the whole table is there.
  const bool pipe_out = !m->capture && (m->world == 1 || (m->peer && m->host_own_layers)) &&
                        (m->host_out[0] || m->host_out[1] || m->host_out[2] || m->host_out[3]);
  const pas::LayerSet own = m->layers();
  const bool lead = m->world == 1 || m->rank == 0;   // copies the 2-D tables
  auto copy_after = [&](cudaStream_t producer, int which, size_t offset, size_t bytes) -> cudaError_t {
    if (m->host_out[which] == nullptr || bytes == 0) return cudaSuccess;
    const DeviceBuffer* buf = nullptr;
    pas_texture_info info;
    if (texture_lookup(m, (pas_texture)which, &buf, &info) != PAS_OK || !info.present) return cudaSuccess;
    cudaError_t e = cudaEventRecord(m->ev_copy, producer);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(m->copy, m->ev_copy, 0);
    if (e == cudaSuccess) {
      e = cudaMemcpyAsync(static_cast<char*>(m->host_out[which]) + offset, static_cast<const char*>(buf->p) + offset,
                          bytes, cudaMemcpyDeviceToHost, m->copy);
    }
    return e;
  };
  // layers `ls` of a 3-D product table (a strided set when the layers are dealt round-robin)
  auto copy_layers_after = [&](cudaStream_t producer, int which, const pas::LayerSet& ls) -> cudaError_t {
    if (m->host_out[which] == nullptr || ls.count() == 0) return cudaSuccess;
    const size_t layer_bytes = m->layer_texels() * m->s_texel_bytes();
    if (ls.stride == 1) return copy_after(producer, which, (size_t)ls.begin * layer_bytes, (size_t)ls.count() * layer_bytes);
    const DeviceBuffer* buf = nullptr;
    pas_texture_info info;
    if (texture_lookup(m, (pas_texture)which, &buf, &info) != PAS_OK || !info.present) return cudaSuccess;
    cudaError_t e = cudaEventRecord(m->ev_copy, producer);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(m->copy, m->ev_copy, 0);
    if (e == cudaSuccess) {
      const size_t off = (size_t)ls.begin * layer_bytes, pitch = (size_t)ls.stride * layer_bytes;
      e = cudaMemcpy2DAsync(static_cast<char*>(m->host_out[which]) + off, pitch, static_cast<const char*>(buf->p) + off,
                            pitch, layer_bytes, (size_t)ls.count(), cudaMemcpyDeviceToHost, m->copy);
    }
    return e;
  };
  // elements [a, b) of a layer set
  auto sub_set = [](const pas::LayerSet& ls, int a, int b) {
    return pas::LayerSet{ls.begin + a * ls.stride, std::min(ls.end, ls.begin + b * ls.stride), ls.stride};
  };
  PhaseTimer timer(m);
  timer.mark("start");
  // final transmittance at 680/550/440 nm (model.cc:951-963). One GPU: filled by the transmittance
  // launch of the first channel group (same optical lengths). Peer worlds compute the transmittance
  // in bands of rows, so there the RGB table is a launch of its own, beside everything else.
  const bool fused_rgb = m->num_precomputed_wavelengths > 3 && (!m->peer || m->symm);
  if (m->num_precomputed_wavelengths > 3 && !fused_rgb) {
    PAS_CUDA(side_after_main());
    PAS_CUDA(pas::launch_transmittance(m->geom, m->rgb_spectrum, m->T_rgb.f(), side));
    PAS_CUDA(pas::launch_pack_rgba(m->T_rgb.f(), (int)m->n_t(), 3, m->T_rgba.f(), side));
    m->launches += 2;
    if (!overlap) timer.mark("final_transmittance");
  }
  m->fuse_rgb_transmittance = fused_rgb;
  for (size_t gi = 0; gi < m->groups.size(); ++gi) {
    const bool blend = gi > 0;  // additive blending for batches after the first (model.cc:946-948)
    const int nc = m->groups[gi].nc, off = m->group_offset[gi];
    pas_status st;
    float* const dS = m->dS.f();
    if ((st = run_phase(m, (int)gi, 0, 0, blend, main, dS, dS)) != PAS_OK) return st;
    if (gi == 0 && m->num_precomputed_wavelengths <= 3) {
      PAS_CUDA(pas::launch_pack_rgba(m->T.f(), (int)m->n_t(), 3, m->T_rgba.f(), main));
      m->launches += 1;
    }
    if (pipe_out && lead && gi == 0 && (fused_rgb || m->num_precomputed_wavelengths <= 3)) {
      PAS_CUDA(copy_after(main, PAS_TEXTURE_TRANSMITTANCE, 0, m->n_t() * 16));
    }
    timer.mark("transmittance");
    if ((st = capture_copy(m, "transmittance", m->T.f(), m->n_t(), nc, off, true)) != PAS_OK) return st;
    if ((st = run_phase(m, (int)gi, 1, 0, blend, main, dS, dS)) != PAS_OK) return st;
    timer.mark("direct_irradiance");
    if ((st = capture_copy(m, "delta_irradiance_1", m->dE.f(), m->n_e(), nc, off, false)) != PAS_OK) return st;
    if ((st = run_phase(m, (int)gi, 2, 0, blend, main, dS, dS)) != PAS_OK) return st;
    const bool last_group = gi + 1 == m->groups.size();
    if (pipe_out && last_group) {
      // the single-Mie table is written by single scattering only (model.cc:151-156)
      PAS_CUDA(copy_layers_after(main, PAS_TEXTURE_SINGLE_MIE, own));
      if (num_scattering_orders == 1) PAS_CUDA(copy_layers_after(main, PAS_TEXTURE_SCATTERING, own));
    }
    timer.mark("single_scattering");
    if ((st = capture_copy(m, "delta_rayleigh", m->dR.f(), m->n_s(), nc, off, true)) != PAS_OK) return st;
    if ((st = capture_copy(m, "delta_mie", m->dM.f(), m->n_s(), nc, off, true)) != PAS_OK) return st;
    // multiple scattering of order n is written to dS (n even) or, overlapped, to the then unused dR
    // (n odd): the irradiance pass running beside it still reads the previous order
    float* ds_in = dS;
    for (unsigned order = 2; order <= num_scattering_orders; ++order) {
      const std::string tag = std::to_string(order);
      float* ds_out = (overlap && (order & 1)) ? m->dR.f() : dS;
      // the density pass reads the irradiance of the previous order: join the side stream
      PAS_CUDA(main_after_side());
      if ((st = run_phase(m, (int)gi, 3, (int)order, blend, main, ds_in, ds_out)) != PAS_OK) return st;
      timer.mark("scattering_density_" + tag);
      // irradiance from the radiance of the previous order (model.cc:1187-1188); it overwrites the
      // irradiance table the density pass has just read
      PAS_CUDA(side_after_main());
      if ((st = run_phase(m, (int)gi, 4, (int)order - 1, blend, side, ds_in, ds_out, overlap ? 1 : 0)) != PAS_OK) return st;
      if (!overlap) timer.mark("indirect_irradiance_" + tag);
      // peer worlds: phase 4 has enqueued the barrier that completes the density slabs on the main
      // stream; what a rank waits there (the slowest rank's density pass + the drain of its stores)
      // is timed apart from the multiple-scattering kernel that follows
      if (overlap && m->peer) timer.mark("exchange_wait_" + tag);
      if ((st = capture_copy(m, "delta_irradiance_" + tag, m->dE.f(), m->n_e(), nc, off, false)) != PAS_OK) return st;
      // (multi-GPU: the density table is complete once the exchange of phase 3 / 4 is done)
      if ((st = capture_copy(m, "delta_density_" + tag, m->cur_dJ(), m->n_s(), nc, off, true)) != PAS_OK) return st;
      if (pipe_out && last_group && order == num_scattering_orders) {
        // E is final (side stream)
        if (lead) PAS_CUDA(copy_after(side, PAS_TEXTURE_IRRADIANCE, 0, m->n_e() * 16));
        if (m->host_out[PAS_TEXTURE_SCATTERING] != nullptr && m->host_S_dev != nullptr) {
          // S becomes final in this pass, and its host destination is pinned memory: the kernel writes
          // every row it finishes to the host table as well -- no copy behind the pass
          m->host_S_now = m->host_S_dev;
          st = run_phase(m, (int)gi, 5, (int)order, blend, main, ds_in, ds_out);
          m->host_S_now = nullptr;
          if (st != PAS_OK) return st;
        } else {
          // pageable destination: S becomes final band by band, each band copied behind its launch
          const int n_own = own.count(), parts = n_own >= 8 ? 4 : (n_own >= 2 ? 2 : 1);
          for (int part = 0; part < parts; ++part) {
            const pas::LayerSet band = sub_set(own, part * n_own / parts, (part + 1) * n_own / parts);
            if ((st = run_phase(m, (int)gi, 5, (int)order, blend, main, ds_in, ds_out, 0, &band)) != PAS_OK) return st;
            PAS_CUDA(copy_layers_after(main, PAS_TEXTURE_SCATTERING, band));
          }
        }
      } else if ((st = run_phase(m, (int)gi, 5, (int)order, blend, main, ds_in, ds_out)) != PAS_OK) {
        return st;
      }
      if (m->peer) m->exchanges += 1;  // the next order uses the other density buffer / xE half
      timer.mark("multiple_scattering_" + tag);
      if ((st = capture_copy(m, "delta_multiple_" + tag, ds_out, m->n_s(), nc, off, true)) != PAS_OK) return st;
      ds_in = ds_out;
    }
    // the next group restarts with the transmittance pass, which the side stream may still read
    PAS_CUDA(main_after_side());
  }
  if (pipe_out && lead) {
    if (num_scattering_orders == 1) PAS_CUDA(copy_after(main, PAS_TEXTURE_IRRADIANCE, 0, m->n_e() * 16));
    if (m->num_precomputed_wavelengths > 3 && !fused_rgb) {
      PAS_CUDA(copy_after(main, PAS_TEXTURE_TRANSMITTANCE, 0, m->n_t() * 16));
    }
  }
  auto join_copies = [&]() -> cudaError_t {
    // the main stream ends after the last copy: one synchronisation covers everything
    if (!pipe_out) return cudaSuccess;
    cudaError_t e = cudaEventRecord(m->ev_copy, m->copy);
    return e != cudaSuccess ? e : cudaStreamWaitEvent(main, m->ev_copy, 0);
  };
  if (!(m->world > 1 && m->peer)) PAS_CUDA(join_copies());
  if (m->world > 1 && m->peer) {
    // every rank ends with the complete scattering table(s): push this rank's layers, then barrier
    const pas::LayerSet ks = m->layers();
    const size_t layer_bytes = m->layer_texels() * m->s_texel_bytes();
    pas::PeerTargets ts{}, tm{};
    if (m->symm) {
      // the product tables are the model's own: the slabs travel through staging areas of the arena
      ts = m->targets_at(m->off_Sx);
      tm = m->targets_at(m->off_Mx);
    } else {
      for (int r = 0; r < m->world; ++r) {
        if (r == m->rank) continue;
        ts.dst[ts.n++] = m->peer_S[r];
        tm.dst[tm.n++] = m->peer_M[r];
      }
    }
    PAS_CUDA(pas::launch_peer_push(m->S.p, layer_bytes, (size_t)ks.begin * layer_bytes, ts, main, ks.count(),
                                   (size_t)ks.stride * layer_bytes));
    if (!m->combined) {
      PAS_CUDA(pas::launch_peer_push(m->M.p, layer_bytes, (size_t)ks.begin * layer_bytes, tm, main, ks.count(),
                                     (size_t)ks.stride * layer_bytes));
    }
    // (host copies of this rank's layers, if any, end before the barrier: passing it means that every
    // rank's part of the shared host tables is written)
    PAS_CUDA(join_copies());
    pas_status st = peer_barrier(m, 0, main);
    if (st != PAS_OK) return st;
    m->launches += m->combined ? 1 : 2;
    if (m->symm) {
      // the other ranks' layers: staging area -> product table (the own slab is already there)
      const size_t lo = (size_t)ks.begin * layer_bytes, hi = (size_t)ks.end * layer_bytes;
      const size_t all = (size_t)m->geom.sz.r_n * layer_bytes;
      auto gather = [&](void* table, size_t off) -> cudaError_t {
        const char* src = m->arena[m->rank] + off;
        cudaError_t e = cudaSuccess;
        if (ks.stride > 1) {
          // dealt layers: the layers of rank r are r, r + world, ...
          const size_t pitch = (size_t)ks.stride * layer_bytes;
          for (int r = 0; r < m->world && e == cudaSuccess; ++r) {
            if (r == m->rank) continue;
            const int n = (m->geom.sz.r_n - r + m->world - 1) / m->world;
            e = cudaMemcpy2DAsync(static_cast<char*>(table) + (size_t)r * layer_bytes, pitch, src + (size_t)r * layer_bytes,
                                  pitch, layer_bytes, (size_t)n, cudaMemcpyDeviceToDevice, main);
          }
          return e;
        }
        if (lo > 0) e = cudaMemcpyAsync(table, src, lo, cudaMemcpyDeviceToDevice, main);
        if (e == cudaSuccess && hi < all) {
          e = cudaMemcpyAsync(static_cast<char*>(table) + hi, src + hi, all - hi, cudaMemcpyDeviceToDevice, main);
        }
        return e;
      };
      PAS_CUDA(gather(m->S.p, m->off_Sx));
      if (!m->combined) PAS_CUDA(gather(m->M.p, m->off_Mx));
      // the staging areas are rewritten by the next Init's final push: every rank must have read them
      // first, which the first barrier of that Init (transmittance rows) guarantees
    }
  } else if (m->world > 1) {
    // every rank ends with the complete scattering table(s)
    const pas::LayerSet ks = m->layers();
    const int k0 = ks.begin, k1 = ks.end;
    const size_t slab_bytes = (size_t)(k1 - k0) * m->layer_texels() * m->s_texel_bytes();
    PAS_NCCL(nccl().GroupStart());
    PAS_NCCL(nccl().AllGather(static_cast<char*>(m->S.p) + (size_t)k0 * m->layer_texels() * m->s_texel_bytes(),
                              m->S.p, slab_bytes, ncclChar, m->comm, m->stream));
    if (!m->combined) {
      PAS_NCCL(nccl().AllGather(static_cast<char*>(m->M.p) + (size_t)k0 * m->layer_texels() * m->s_texel_bytes(),
                                m->M.p, slab_bytes, ncclChar, m->comm, m->stream));
    }
    PAS_NCCL(nccl().GroupEnd());
  }
  if (!pipe_out) {
    // captures or a multi-GPU world: registered host outputs are filled by plain copies at the end
    for (int which = 0; which < 4; ++which) {
      if (m->host_out[which] == nullptr) continue;
      const DeviceBuffer* buf = nullptr;
      pas_texture_info info;
      if (texture_lookup(m, (pas_texture)which, &buf, &info) != PAS_OK || !info.present) continue;
      const size_t bytes = (size_t)info.width * info.height * info.depth * 4 * info.bytes_per_channel;
      PAS_CUDA(cudaMemcpyAsync(m->host_out[which], buf->p, bytes, cudaMemcpyDeviceToHost, main));
    }
  }
  timer.mark("finalize");
  m->in_flight = true;
  return PAS_OK;
}

pas_status pas_model_texture_info(const pas_model* m, pas_texture which, pas_texture_info* info) {
  if (m == nullptr || info == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  return texture_lookup(m, which, nullptr, info);
}

pas_status pas_model_texture_device_ptr(const pas_model* m, pas_texture which, const void** ptr) {
  { pas_status settled = settle(m); if (settled != PAS_OK) return settled; }
  if (m == nullptr || ptr == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  const DeviceBuffer* buf = nullptr;
  pas_texture_info info;
  pas_status st = texture_lookup(m, which, &buf, &info);
  if (st != PAS_OK) return st;
  if (!info.present) return fail(PAS_ERR_STATE, "this model has no such table");
  *ptr = buf->p;
  return PAS_OK;
}

pas_status pas_model_read_texture(pas_model* m, pas_texture which, int as_float32, void* dst,
                                  size_t dst_bytes) {
  { pas_status settled = settle(m); if (settled != PAS_OK) return settled; }
  if (m == nullptr || dst == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!m->initialised) return fail(PAS_ERR_STATE, "pas_model_init has not been called");
  const DeviceBuffer* buf = nullptr;
  pas_texture_info info;
  pas_status st = texture_lookup(m, which, &buf, &info);
  if (st != PAS_OK) return st;
  if (!info.present) return fail(PAS_ERR_STATE, "this model has no such table");
  PAS_CUDA(cudaSetDevice(m->device));
  const size_t values = (size_t)info.width * info.height * info.depth * 4;
  const bool convert = as_float32 && info.bytes_per_channel == 2;
  const size_t need = values * (convert || info.bytes_per_channel == 4 ? 4 : 2);
  if (dst_bytes != need) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "dst_bytes is " + std::to_string(dst_bytes) + ", table needs " +
                                              std::to_string(need));
  }
  const void* src = buf->p;
  if (convert) {
    PAS_CUDA(m->scratch.ensure(values * sizeof(float)));
    half_to_float_kernel<<<(unsigned)((values + 255) / 256), 256, 0, m->stream>>>(
        static_cast<const __half*>(buf->p), values, m->scratch.f());
    PAS_CUDA(cudaGetLastError());
    src = m->scratch.p;
  }
  PAS_CUDA(cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, m->stream));
  PAS_CUDA(cudaStreamSynchronize(m->stream));
  return PAS_OK;
}

pas_status pas_model_save_dat(pas_model* m, const char* directory) {
  { pas_status settled = settle(m); if (settled != PAS_OK) return settled; }
  if (m == nullptr || directory == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  static const struct { pas_texture id; const char* file; } kFiles[] = {
      {PAS_TEXTURE_TRANSMITTANCE, "transmittance.dat"},
      {PAS_TEXTURE_SCATTERING, "scattering.dat"},
      {PAS_TEXTURE_IRRADIANCE, "irradiance.dat"},
      {PAS_TEXTURE_SINGLE_MIE, "single_mie_scattering.dat"}};
  for (const auto& f : kFiles) {
    pas_texture_info info;
    pas_status st = texture_lookup(m, f.id, nullptr, &info);
    if (st != PAS_OK) return st;
    if (!info.present) continue;
    std::vector<float> host((size_t)info.width * info.height * info.depth * 4);
    st = pas_model_read_texture(m, f.id, 1, host.data(), host.size() * sizeof(float));
    if (st != PAS_OK) return st;
    const std::string path = std::string(directory) + "/" + f.file;
    std::ofstream out(path, std::ios::binary);
    out.write(reinterpret_cast<const char*>(host.data()), host.size() * sizeof(float));
    if (!out) return fail(PAS_ERR_IO, "cannot write " + path);
  }
  return PAS_OK;
}

pas_status pas_model_luminance_factors(const pas_model* m, double* out6) {
  if (m == nullptr || out6 == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  for (int a = 0; a < 3; ++a) {
    out6[a] = m->sky_k[a];
    out6[3 + a] = m->sun_k[a];
  }
  return PAS_OK;
}

pas_status pas_convert_spectrum_to_linear_srgb(size_t n, const double* wavelengths,
                                               const double* spectrum, double* r, double* g,
                                               double* b) {
  if (n < 1 || !wavelengths || !spectrum || !r || !g || !b) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  }
  const std::vector<double> wl(wavelengths, wavelengths + n), sp(spectrum, spectrum + n);
  double xyz[3] = {0, 0, 0};
  for (int lambda = 360; lambda < 830; ++lambda) {
    const double v = interpolate(wl, sp, lambda);
    for (int a = 0; a < 3; ++a) xyz[a] += cie_value(lambda, a + 1) * v;
  }
  double* out[3] = {r, g, b};
  for (int a = 0; a < 3; ++a) {
    *out[a] = pas::kMaxLuminousEfficacy * (pas::kXyzToSrgb[a][0] * xyz[0] + pas::kXyzToSrgb[a][1] * xyz[1] +
                                           pas::kXyzToSrgb[a][2] * xyz[2]);
  }
  return PAS_OK;
}

pas_status pas_model_channels(const pas_model* m, int* num_channels, double* lambdas) {
  if (m == nullptr || num_channels == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  *num_channels = m->total_channels();
  if (lambdas) std::copy(m->lambdas.begin(), m->lambdas.end(), lambdas);
  return PAS_OK;
}

pas_status pas_model_luminance_matrix(const pas_model* m, float* out) {
  if (m == nullptr || out == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  std::copy(m->lum.begin(), m->lum.end(), out);
  return PAS_OK;
}

pas_status pas_spectral_channels(unsigned int num_precomputed_wavelengths, int* num_channels,
                                 double* lambdas, float* luminance_from_radiance) {
  if (num_channels == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (num_precomputed_wavelengths < 1 || num_precomputed_wavelengths > 240) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "num_precomputed_wavelengths must be in [1, 240]");
  }
  std::vector<double> lam;
  std::vector<float> lum;
  spectral_channels(num_precomputed_wavelengths, &lam, &lum);
  *num_channels = (int)lam.size();
  if (lambdas) std::copy(lam.begin(), lam.end(), lambdas);
  if (luminance_from_radiance) std::copy(lum.begin(), lum.end(), luminance_from_radiance);
  return PAS_OK;
}

pas_status pas_model_set_capture(pas_model* m, int enabled) {
  if (m == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "model is NULL");
  m->capture = enabled != 0;
  if (!m->capture) m->captured.clear();
  return PAS_OK;
}

pas_status pas_model_read_intermediate(pas_model* m, const char* name, float* dst, size_t* num_floats) {
  { pas_status settled = settle(m); if (settled != PAS_OK) return settled; }
  if (m == nullptr || name == nullptr || num_floats == nullptr) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  }
  PAS_CUDA(cudaSetDevice(m->device));
  const float* src = nullptr;
  size_t count = 0;
  auto it = m->captured.find(name);
  float* live = nullptr;
  size_t texels = 0;
  bool interleaved = false;
  const int nc0 = m->groups[0].nc;
  if (it != m->captured.end()) {
    // captured copies hold every channel of every group, planar
    src = it->second->f();
    count = it->second->bytes / sizeof(float);
  } else if (live_buffer(m, name, &live, &texels, &interleaved)) {
    src = live;
    count = texels * nc0;
  } else {
    return fail(PAS_ERR_STATE, std::string("no intermediate named '") + name + "' (capture enabled?)");
  }
  if (dst == nullptr) {
    *num_floats = count;
    return PAS_OK;
  }
  if (*num_floats < count) return fail(PAS_ERR_INVALID_ARGUMENT, "destination too small");
  if (src == live && interleaved) {
    PAS_CUDA(m->scratch.ensure(count * sizeof(float)));
    PAS_CUDA(pas::launch_interleaved_to_planar(live, texels, nc0, m->scratch.f(), m->stream));
    src = m->scratch.f();
  }
  PAS_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  PAS_CUDA(cudaStreamSynchronize(m->stream));
  *num_floats = count;
  return PAS_OK;
}

pas_status pas_model_write_intermediate(pas_model* m, const char* name, const float* src,
                                        size_t num_floats) {
  { pas_status settled = settle(m); if (settled != PAS_OK) return settled; }
  if (m == nullptr || name == nullptr || src == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  PAS_CUDA(cudaSetDevice(m->device));
  float* live = nullptr;
  size_t texels = 0;
  bool interleaved = false;
  if (!live_buffer(m, name, &live, &texels, &interleaved)) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "unknown live buffer");
  }
  const int nc0 = m->groups[0].nc;
  if (num_floats != texels * nc0) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "expected " + std::to_string(texels * nc0) + " floats");
  }
  if (interleaved) {
    PAS_CUDA(m->scratch.ensure(num_floats * sizeof(float)));
    PAS_CUDA(cudaMemcpyAsync(m->scratch.f(), src, num_floats * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    PAS_CUDA(pas::launch_planar_to_interleaved(m->scratch.f(), texels, nc0, live, m->stream));
  } else {
    PAS_CUDA(cudaMemcpyAsync(live, src, num_floats * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  }
  PAS_CUDA(cudaStreamSynchronize(m->stream));
  return PAS_OK;
}

pas_status pas_model_run_phase(pas_model* m, int phase, int order) {
  { pas_status settled = settle(m); if (settled != PAS_OK) return settled; }
  if (m == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "model is NULL");
  if (m->groups.size() != 1) return fail(PAS_ERR_UNSUPPORTED, "single passes need <= 16 channels");
  PAS_CUDA(cudaSetDevice(m->device));
  if ((phase == 2 || phase == 5) && m->ray_setup.p != nullptr) {
    // single passes may follow a pas_model_write_intermediate("transmittance"): the ray tables are rebuilt
    // from the transmittance buffer as it stands
    PAS_CUDA(pas::launch_ray_setup(m->geom, m->groups[0], m->T.f(), m->ray_setup.p, m->layers(), m->stream));
  }
  pas_status st = run_phase(m, 0, phase, order, false, m->stream, m->dS.f(), m->dS.f());
  if (st != PAS_OK) return st;
  PAS_CUDA(cudaStreamSynchronize(m->stream));
  return PAS_OK;
}

pas_status pas_model_last_timings(const pas_model* m, int* count, const char** names, float* ms) {
  { pas_status settled = settle(m); if (settled != PAS_OK) return settled; }
  if (m == nullptr || count == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  const int cap = *count;
  *count = (int)m->timings.size();
  for (int i = 0; i < cap && i < (int)m->timings.size(); ++i) {
    if (names) names[i] = m->timings[i].first.c_str();
    if (ms) ms[i] = m->timings[i].second;
  }
  return PAS_OK;
}

pas_status pas_model_last_launch_count(const pas_model* m, int* launches) {
  if (m == nullptr || launches == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  *launches = m->launches;
  return PAS_OK;
}

int pas_world_is_cached(int device, int rank, int world_size) {
  CommCache& cache = comm_cache();
  std::lock_guard<std::mutex> lock(cache.mu);
  return cache.comms.count(std::make_tuple(device, rank, world_size)) ? 1 : 0;
}

void pas_release_cached_memory(void) { pool().release_all(); }

pas_status pas_nccl_unique_id(void* id_bytes) {
  if (id_bytes == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!nccl().ok) return fail(PAS_ERR_NCCL, "libnccl.so.2 could not be loaded");
  static_assert(sizeof(ncclUniqueId) == PAS_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  PAS_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(id_bytes, &id, sizeof id);
  return PAS_OK;
}

pas_status pas_model_attach_world(pas_model* m, int rank, int world_size, const void* id_bytes) {
  if (m == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "model is NULL");
  if (world_size < 1 || rank < 0 || rank >= world_size) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "bad rank / world size");
  }
  if (world_size == 1) {
    if (m->symm) {
      pas_status st = leave_arena(m);
      if (st != PAS_OK) return st;
    }
    m->rank = 0;
    m->world = 1;
    m->peer = false;
    return PAS_OK;
  }
  if (m->geom.sz.r_n % world_size != 0) {
    return fail(PAS_ERR_UNSUPPORTED, "scattering_r must be divisible by the world size");
  }
  if (!nccl().ok) return fail(PAS_ERR_NCCL, "libnccl.so.2 could not be loaded");
  PAS_CUDA(cudaSetDevice(m->device));
  {
    CommCache& cache = comm_cache();
    std::lock_guard<std::mutex> lock(cache.mu);
    const auto key = std::make_tuple(m->device, rank, world_size);
    auto it = cache.comms.find(key);
    if (it != cache.comms.end()) {
      m->comm = it->second;  // communicator of an earlier model of this process
    } else {
      if (id_bytes == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL unique id");
      ncclUniqueId id;
      std::memcpy(&id, id_bytes, sizeof id);
      PAS_NCCL(nccl().CommInitRank(&m->comm, world_size, id, rank));
      cache.comms[key] = m->comm;
    }
  }
  if (m->symm) {
    // back to buffers of its own
    pas_status st = leave_arena(m);
    if (st != PAS_OK) return st;
  }
  m->rank = rank;
  m->world = world_size;
  m->peer = false;  // a model attached to an NCCL world no longer uses a peer mapping it may have had
  return PAS_OK;
}

pas_status pas_model_ipc_export(pas_model* m, int rank, int world_size, void* out, size_t* bytes) {
  if (m == nullptr || bytes == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (out == nullptr || *bytes < sizeof(PasIpcExport)) {
    *bytes = sizeof(PasIpcExport);
    return out == nullptr ? PAS_OK : fail(PAS_ERR_INVALID_ARGUMENT, "export buffer too small");
  }
  if (world_size < 2 || world_size > PAS_MAX_PEERS + 1 || rank < 0 || rank >= world_size) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "peer worlds have 2..8 ranks");
  }
  if (m->geom.sz.r_n % world_size != 0) {
    return fail(PAS_ERR_UNSUPPORTED, "scattering_r must be divisible by the world size");
  }
  PAS_CUDA(cudaSetDevice(m->device));
  if (m->symm) {
    pas_status st = leave_arena(m);
    if (st != PAS_OK) return st;
    m->peer = false;
    m->world = 1;
    m->rank = 0;
  }
  // rank and world take effect in pas_model_attach_peers, once every mapping has succeeded
  m->pending_rank = rank;
  m->pending_world = world_size;
  const size_t cp = PAS_CHANNEL_PITCH(m->max_nc());
  PAS_CUDA(m->dJ2.ensure(m->n_s() * cp * sizeof(float)));
  PAS_CUDA(m->xE.ensure((size_t)2 * world_size * m->xe_stride() * sizeof(float)));
  if (!m->combined) PAS_CUDA(m->M.ensure(m->n_s() * m->s_texel_bytes()));
  {
    PeerCache& cache = peer_cache();
    std::lock_guard<std::mutex> lock(cache.mu);
    PeerWorld& w = cache.worlds[std::make_tuple(m->device, rank, world_size)];
    if (w.flags == nullptr) {
      PAS_CUDA(cudaMalloc(&w.flags, 64 * sizeof(unsigned)));
      PAS_CUDA(cudaMemset(w.flags, 0, 64 * sizeof(unsigned)));
      PAS_CUDA(cudaHostAlloc(&w.error_host, sizeof(int), cudaHostAllocMapped));
      *w.error_host = 0;
      PAS_CUDA(cudaHostGetDevicePointer(&w.error_dev, w.error_host, 0));
      PAS_CUDA(cudaDeviceSynchronize());
    }
    m->pw = &w;
  }
  PasIpcExport e{};
  PAS_CUDA(cudaIpcGetMemHandle(&e.T, m->T.p));
  PAS_CUDA(cudaIpcGetMemHandle(&e.dJ[0], m->dJ.p));
  PAS_CUDA(cudaIpcGetMemHandle(&e.dJ[1], m->dJ2.p));
  PAS_CUDA(cudaIpcGetMemHandle(&e.S, m->S.p));
  e.has_M = m->combined ? 0 : 1;
  if (e.has_M) PAS_CUDA(cudaIpcGetMemHandle(&e.M, m->M.p));
  PAS_CUDA(cudaIpcGetMemHandle(&e.xE, m->xE.p));
  PAS_CUDA(cudaIpcGetMemHandle(&e.flags, m->pw->flags));
  std::memcpy(out, &e, sizeof e);
  *bytes = sizeof e;
  return PAS_OK;
}

namespace {
pas_status open_peer(const cudaIpcMemHandle_t& h, void** ptr) {
  PeerCache& cache = peer_cache();
  std::lock_guard<std::mutex> lock(cache.mu);
  const std::string key(reinterpret_cast<const char*>(&h), sizeof h);
  auto it = cache.opened.find(key);
  if (it != cache.opened.end()) {
    *ptr = it->second;
    return PAS_OK;
  }
  PAS_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  cache.opened[key] = *ptr;
  return PAS_OK;
}
}  // namespace

pas_status pas_model_attach_peers(pas_model* m, const void* exports, size_t bytes_per_rank) {
  if (m == nullptr || exports == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (m->pw == nullptr || m->pending_world < 2) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "call pas_model_ipc_export first");
  }
  if (bytes_per_rank != sizeof(PasIpcExport)) return fail(PAS_ERR_INVALID_ARGUMENT, "bad export size");
  PAS_CUDA(cudaSetDevice(m->device));
  const int rank = m->pending_rank, world = m->pending_world;
  m->peer = false;  // until every peer is mapped
  for (int r = 0; r < world; ++r) {
    PasIpcExport e;
    std::memcpy(&e, static_cast<const char*>(exports) + (size_t)r * bytes_per_rank, sizeof e);
    if (r == rank) {
      m->peer_T[r] = m->T.f();
      m->peer_dJ[0][r] = m->dJ.f();
      m->peer_dJ[1][r] = m->dJ2.f();
      m->peer_S[r] = m->S.p;
      m->peer_M[r] = m->M.p;
      m->peer_xE[r] = m->xE.f();
      m->peer_flags[r] = m->pw->flags;
      continue;
    }
    if ((e.has_M != 0) == m->combined) return fail(PAS_ERR_INVALID_ARGUMENT, "ranks disagree on the model");
    pas_status st;
    void* p = nullptr;
    if ((st = open_peer(e.T, &p)) != PAS_OK) return st;
    m->peer_T[r] = static_cast<float*>(p);
    if ((st = open_peer(e.dJ[0], &p)) != PAS_OK) return st;
    m->peer_dJ[0][r] = static_cast<float*>(p);
    if ((st = open_peer(e.dJ[1], &p)) != PAS_OK) return st;
    m->peer_dJ[1][r] = static_cast<float*>(p);
    if ((st = open_peer(e.S, &p)) != PAS_OK) return st;
    m->peer_S[r] = p;
    if (e.has_M) {
      if ((st = open_peer(e.M, &p)) != PAS_OK) return st;
      m->peer_M[r] = p;
    }
    if ((st = open_peer(e.xE, &p)) != PAS_OK) return st;
    m->peer_xE[r] = static_cast<float*>(p);
    if ((st = open_peer(e.flags, &p)) != PAS_OK) return st;
    m->peer_flags[r] = static_cast<unsigned*>(p);
  }
  m->rank = rank;
  m->world = world;
  m->peer = true;
  m->exchanges = 0;
  return PAS_OK;
}

namespace {
// Layout of a model's exchange tables inside a symmetric arena: flag words first (fixed place, whatever
// the model), then the tables, each aligned to 1 KiB.
struct ArenaLayout {
  size_t flags = 0, T = 0, dJ[2] = {0, 0}, xE = 0, Sx = 0, Mx = 0, Tx = 0, bytes = 0;
};
ArenaLayout arena_layout(const pas_model* m, int world) {
  ArenaLayout a;
  size_t at = 4096;  // PAS_FLAG_CHANNELS x PAS_FLAG_WORDS words, generously
  auto take = [&](size_t n) {
    const size_t here = at;
    at += (n + 1023) / 1024 * 1024;
    return here;
  };
  const size_t cp = PAS_CHANNEL_PITCH(m->max_nc());
  a.T = take(m->n_t() * cp * sizeof(float));
  a.dJ[0] = take(m->n_s() * cp * sizeof(float));
  a.dJ[1] = take(m->n_s() * cp * sizeof(float));
  a.xE = take((size_t)2 * world * m->xe_stride() * sizeof(float));
  a.Sx = take(m->n_s() * m->s_texel_bytes());
  a.Mx = m->combined ? a.Sx : take(m->n_s() * m->s_texel_bytes());
  a.Tx = take(m->n_t() * 16);   // staging of the final RGBA transmittance (rows computed by different ranks)
  a.bytes = at;
  return a;
}
}  // namespace

pas_status pas_model_exchange_bytes(const pas_model* m, int world_size, size_t* bytes) {
  if (m == nullptr || bytes == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (world_size < 2 || world_size > PAS_MAX_PEERS + 1) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "symmetric worlds have 2..8 ranks");
  }
  *bytes = arena_layout(m, world_size).bytes;
  return PAS_OK;
}

pas_status pas_model_attach_symmetric(pas_model* m, int rank, int world_size, void* const* arena_bases,
                                      void* multicast_base, size_t bytes) {
  if (m == nullptr || arena_bases == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (world_size < 2 || world_size > PAS_MAX_PEERS + 1 || rank < 0 || rank >= world_size) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "symmetric worlds have 2..8 ranks");
  }
  if (m->geom.sz.r_n % world_size != 0) {
    return fail(PAS_ERR_UNSUPPORTED, "scattering_r must be divisible by the world size");
  }
  { pas_status settled = settle(m); if (settled != PAS_OK) return settled; }
  const ArenaLayout a = arena_layout(m, world_size);
  if (bytes < a.bytes) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "arena of " + std::to_string(bytes) + " bytes, this model needs " +
                                              std::to_string(a.bytes));
  }
  for (int r = 0; r < world_size; ++r) {
    if (arena_bases[r] == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL arena base");
  }
  PAS_CUDA(cudaSetDevice(m->device));
  {
    // one barrier sequence per (device, rank, world) and kind of world; the flag words (and with them
    // the epoch counters) live in the arena, so a new, zeroed arena restarts the sequence on every rank
    PeerCache& cache = peer_cache();
    std::lock_guard<std::mutex> lock(cache.mu);
    PeerWorld& w = cache.worlds[std::make_tuple(m->device, rank, world_size | 0x100)];
    if (w.error_host == nullptr) {
      PAS_CUDA(cudaHostAlloc(&w.error_host, sizeof(int), cudaHostAllocMapped));
      *w.error_host = 0;
      PAS_CUDA(cudaHostGetDevicePointer(&w.error_dev, w.error_host, 0));
    }
    if (w.in_flight != nullptr && w.in_flight != m) {
      return fail(PAS_ERR_STATE, "another model of this world has an Init in flight: wait for it first");
    }
    if (w.flags != static_cast<unsigned*>(arena_bases[rank])) {
      w.flags = static_cast<unsigned*>(arena_bases[rank]);  // a new arena: a fresh barrier sequence
      w.broken = false;
    }
    m->pw = &w;
  }
  for (int r = 0; r < world_size; ++r) {
    m->arena[r] = static_cast<char*>(arena_bases[r]);
    m->peer_flags[r] = reinterpret_cast<unsigned*>(m->arena[r] + a.flags);
  }
  m->arena_mc = static_cast<char*>(multicast_base);
  m->off_T = a.T;
  m->off_dJ[0] = a.dJ[0];
  m->off_dJ[1] = a.dJ[1];
  m->off_xE = a.xE;
  m->off_Sx = a.Sx;
  m->off_Mx = a.Mx;
  m->off_Tx = a.Tx;
  const size_t cp = PAS_CHANNEL_PITCH(m->max_nc());
  char* mine = m->arena[rank];
  m->T.adopt(mine + a.T, m->n_t() * cp * sizeof(float));
  m->dJ.adopt(mine + a.dJ[0], m->n_s() * cp * sizeof(float));
  m->dJ2.adopt(mine + a.dJ[1], m->n_s() * cp * sizeof(float));
  m->xE.adopt(mine + a.xE, (size_t)2 * world_size * m->xe_stride() * sizeof(float));
  m->rank = rank;
  m->world = world_size;
  m->peer = true;
  m->symm = true;
  m->comm = nullptr;
  m->exchanges = 0;
  return PAS_OK;
}

}  // extern "C"

// ---- render-time use of the tables (kernel_render.cu) ------------------------------------------
namespace {

bool is_device_pointer(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

pas_status render_ready(const pas_model* m, int use_luminance) {
  { pas_status settled = settle(m); if (settled != PAS_OK) return settled; }
  if (m == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "model is NULL");
  if (!m->initialised) return fail(PAS_ERR_STATE, "pas_model_init has not been called");
  if (!use_luminance && m->num_precomputed_wavelengths > 3) {
    // RADIANCE_API_ENABLED is only defined for <= 3 wavelengths (model.cc:771)
    return fail(PAS_ERR_STATE, "the radiance API needs num_precomputed_wavelengths <= 3 (use luminance)");
  }
  return PAS_OK;
}

pas::RenderTables render_tables(const pas_model* m) {
  pas::RenderTables t{};
  t.transmittance = static_cast<const float4*>(m->T_rgba.p);
  t.scattering = m->S.p;
  t.single_mie = m->combined ? nullptr : m->M.p;
  t.irradiance = static_cast<const float4*>(m->E.p);
  t.half_precision = m->half ? 1 : 0;
  return t;
}

pas::RenderConstants render_constants(const pas_model* m, int use_luminance) {
  pas::RenderConstants c{};
  for (int a = 0; a < 3; ++a) {
    c.solar[a] = m->rgb_spectrum.solar[a];
    c.rayleigh[a] = m->rgb_spectrum.beta_r[a];
    c.mie_sca[a] = m->rgb_spectrum.beta_m_sca[a];
    c.sky_k[a] = use_luminance ? m->sky_k[a] : 1.0;
    c.sun_k[a] = use_luminance ? m->sun_k[a] : 1.0;
  }
  return c;
}

// Device view of an input array: the pointer itself if it is device memory, else a staged copy.
pas_status stage_in(pas_model* m, int slot, const void* src, size_t bytes, const void** dev) {
  if (src == nullptr) {
    *dev = nullptr;
    return PAS_OK;
  }
  if (is_device_pointer(src)) {
    *dev = src;
    return PAS_OK;
  }
  PAS_CUDA(m->render_in[slot].ensure(bytes));
  PAS_CUDA(cudaMemcpyAsync(m->render_in[slot].p, src, bytes, cudaMemcpyHostToDevice, m->stream));
  *dev = m->render_in[slot].p;
  return PAS_OK;
}
pas_status stage_out(pas_model* m, int slot, void* dst, size_t bytes, void** dev) {
  if (dst == nullptr) {
    *dev = nullptr;
    return PAS_OK;
  }
  if (is_device_pointer(dst)) {
    *dev = dst;
    return PAS_OK;
  }
  PAS_CUDA(m->render_out[slot].ensure(bytes));
  *dev = m->render_out[slot].p;
  return PAS_OK;
}
pas_status unstage_out(pas_model* m, void* dst, const void* dev, size_t bytes) {
  if (dst != nullptr && dst != dev) {
    PAS_CUDA(cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, m->stream));
  }
  return PAS_OK;
}

struct RenderTimer {
  pas_model* m;
  cudaEvent_t a = nullptr, b = nullptr;
  explicit RenderTimer(pas_model* model) : m(model) {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, m->stream);
  }
  void stop() { cudaEventRecord(b, m->stream); }
  ~RenderTimer() {
    if (cudaEventSynchronize(b) == cudaSuccess) cudaEventElapsedTime(&m->last_render_ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  }
};

pas_status sky_radiance_common(pas_model* m, int use_luminance, bool to_point, size_t n,
                               const double* camera, const double* target, const double* shadow_length,
                               const double* sun_direction, float* radiance, float* transmittance) {
  pas_status st = render_ready(m, use_luminance);
  if (st != PAS_OK) return st;
  if (n == 0) return PAS_OK;
  if (!camera || !target || !sun_direction || !radiance) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  PAS_CUDA(cudaSetDevice(m->device));
  const void *d_cam, *d_tgt, *d_sl, *d_sun;
  void *d_rad, *d_tr;
  if ((st = stage_in(m, 0, camera, n * 3 * sizeof(double), &d_cam)) != PAS_OK) return st;
  if ((st = stage_in(m, 1, target, n * 3 * sizeof(double), &d_tgt)) != PAS_OK) return st;
  if ((st = stage_in(m, 2, shadow_length, n * sizeof(double), &d_sl)) != PAS_OK) return st;
  if ((st = stage_in(m, 3, sun_direction, n * 3 * sizeof(double), &d_sun)) != PAS_OK) return st;
  if ((st = stage_out(m, 0, radiance, n * 3 * sizeof(float), &d_rad)) != PAS_OK) return st;
  if ((st = stage_out(m, 1, transmittance, n * 3 * sizeof(float), &d_tr)) != PAS_OK) return st;
  {
    RenderTimer timer(m);
    cudaError_t e = pas::launch_sky_radiance(
        m->geom, render_tables(m), render_constants(m, use_luminance), n, to_point,
        static_cast<const double*>(d_cam), static_cast<const double*>(d_tgt), static_cast<const double*>(d_sl),
        static_cast<const double*>(d_sun), static_cast<float*>(d_rad), static_cast<float*>(d_tr), m->stream);
    timer.stop();
    PAS_CUDA(e);
  }
  if ((st = unstage_out(m, radiance, d_rad, n * 3 * sizeof(float))) != PAS_OK) return st;
  if ((st = unstage_out(m, transmittance, d_tr, n * 3 * sizeof(float))) != PAS_OK) return st;
  PAS_CUDA(cudaStreamSynchronize(m->stream));
  return PAS_OK;
}

}  // namespace

extern "C" {

pas_status pas_model_get_solar_radiance(const pas_model* m, int use_luminance, double* rgb) {
  if (m == nullptr || rgb == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!use_luminance && m->num_precomputed_wavelengths > 3) {
    return fail(PAS_ERR_STATE, "the radiance API needs num_precomputed_wavelengths <= 3 (use luminance)");
  }
  // model.cc:228-231, 254-258
  const double a = m->sun_angular_radius;
  for (int c = 0; c < 3; ++c) {
    rgb[c] = m->rgb_spectrum.solar[c] / (pas::kPi * a * a) * (use_luminance ? m->sun_k[c] : 1.0);
  }
  return PAS_OK;
}

pas_status pas_model_get_sky_radiance(pas_model* m, int use_luminance, size_t n, const double* camera,
                                      const double* view_ray, const double* shadow_length,
                                      const double* sun_direction, float* radiance, float* transmittance) {
  return sky_radiance_common(m, use_luminance, false, n, camera, view_ray, shadow_length, sun_direction,
                             radiance, transmittance);
}

pas_status pas_model_get_sky_radiance_to_point(pas_model* m, int use_luminance, size_t n,
                                               const double* camera, const double* point,
                                               const double* shadow_length, const double* sun_direction,
                                               float* radiance, float* transmittance) {
  return sky_radiance_common(m, use_luminance, true, n, camera, point, shadow_length, sun_direction,
                             radiance, transmittance);
}

pas_status pas_model_get_sun_and_sky_irradiance(pas_model* m, int use_luminance, size_t n,
                                                const double* point, const double* normal,
                                                const double* sun_direction, float* sun_irradiance,
                                                float* sky_irradiance) {
  pas_status st = render_ready(m, use_luminance);
  if (st != PAS_OK) return st;
  if (n == 0) return PAS_OK;
  if (!point || !normal || !sun_direction || !sun_irradiance || !sky_irradiance) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  }
  PAS_CUDA(cudaSetDevice(m->device));
  const void *d_p, *d_n, *d_sun;
  void *d_e0, *d_e1;
  if ((st = stage_in(m, 0, point, n * 3 * sizeof(double), &d_p)) != PAS_OK) return st;
  if ((st = stage_in(m, 1, normal, n * 3 * sizeof(double), &d_n)) != PAS_OK) return st;
  if ((st = stage_in(m, 3, sun_direction, n * 3 * sizeof(double), &d_sun)) != PAS_OK) return st;
  if ((st = stage_out(m, 0, sun_irradiance, n * 3 * sizeof(float), &d_e0)) != PAS_OK) return st;
  if ((st = stage_out(m, 1, sky_irradiance, n * 3 * sizeof(float), &d_e1)) != PAS_OK) return st;
  {
    RenderTimer timer(m);
    cudaError_t e = pas::launch_sun_and_sky_irradiance(
        m->geom, render_tables(m), render_constants(m, use_luminance), n, static_cast<const double*>(d_p),
        static_cast<const double*>(d_n), static_cast<const double*>(d_sun), static_cast<float*>(d_e0),
        static_cast<float*>(d_e1), m->stream);
    timer.stop();
    PAS_CUDA(e);
  }
  if ((st = unstage_out(m, sun_irradiance, d_e0, n * 3 * sizeof(float))) != PAS_OK) return st;
  if ((st = unstage_out(m, sky_irradiance, d_e1, n * 3 * sizeof(float))) != PAS_OK) return st;
  PAS_CUDA(cudaStreamSynchronize(m->stream));
  return PAS_OK;
}

pas_status pas_model_render_context(pas_model* m, int use_luminance, void* out, size_t* bytes) {
  if (bytes == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (out == nullptr) {
    *bytes = sizeof(pas::RenderContext);
    return PAS_OK;
  }
  pas_status st = render_ready(m, use_luminance);
  if (st != PAS_OK) return st;
  if (*bytes < sizeof(pas::RenderContext)) return fail(PAS_ERR_INVALID_ARGUMENT, "context buffer too small");
  const pas::RenderContext ctx{m->geom, render_tables(m), render_constants(m, use_luminance)};
  std::memcpy(out, &ctx, sizeof ctx);
  *bytes = sizeof ctx;
  return PAS_OK;
}

pas_status pas_model_last_render_ms(const pas_model* m, float* ms) {
  if (m == nullptr || ms == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  *ms = m->last_render_ms;
  return PAS_OK;
}

}  // extern "C"

// ---- GLSL source of the rendering shader (atmosphere/model.cc:691-744, 769-772) -------------------
namespace {

std::string glsl_number(double v) {
  // full float precision (the reference prints 6 decimals through std::to_string,
  // model.cc:636-652; the tables here are computed from the exact values, so the renderer gets them too)
  char buf[64];
  std::snprintf(buf, sizeof buf, "%.9g", v);
  std::string t(buf);
  if (t.find_first_of(".eE") == std::string::npos) t += ".0";
  return t;
}

std::string glsl_vec3(const pas_model& m, const std::vector<double>& v, double scale) {
  const double lam[3] = {680.0, 550.0, 440.0};
  std::string t = "vec3(";
  for (int a = 0; a < 3; ++a) {
    t += glsl_number(interpolate(m.wavelengths, v, lam[a]) * scale);
    t += a < 2 ? "," : ")";
  }
  return t;
}

std::string glsl_profile(const PasGeometry& g, int profile) {
  std::string t = "DensityProfile(DensityProfileLayer[2](";
  for (int l = 0; l < 2; ++l) {
    t += "DensityProfileLayer(";
    for (int f = 0; f < 5; ++f) {
      t += glsl_number(g.profiles[profile][l][f]);
      t += f < 4 ? "," : ")";
    }
    t += l == 0 ? "," : "))";
  }
  return t;
}

bool read_text(const std::string& path, std::string* out) {
  std::ifstream in(path, std::ios::binary);
  if (!in) return false;
  std::stringstream ss;
  ss << in.rdbuf();
  *out = ss.str();
  return true;
}

// One GLSL wrapper per public entry point of the rendering API (the functions the demo and the
// integration test call, atmosphere/demo/demo.glsl:304-380): each forwards to the function of the
// same base name in functions.glsl with the ATMOSPHERE constant and the table samplers bound.
struct ApiEntry {
  const char* result;      // return type
  const char* name;        // exported name
  const char* params;      // parameter list
  const char* callee;      // functions.glsl function
  const char* tables;      // sampler arguments
  const char* args;        // forwarded arguments
  const char* out_scale;   // constant applied to the `out` sky irradiance, or ""
  const char* scale;       // constant applied to the result, or ""
  bool radiance_only;      // only when RADIANCE_API_ENABLED
};

std::string api_wrappers() {
  static const char kSkyParams[] =
      "Position camera, Direction view_ray, Length shadow_length, Direction sun_direction, "
      "out DimensionlessSpectrum transmittance";
  static const char kPointParams[] =
      "Position camera, Position point, Length shadow_length, Direction sun_direction, "
      "out DimensionlessSpectrum transmittance";
  static const char kIrrParams[] =
      "Position p, Direction normal, Direction sun_direction, out IrradianceSpectrum sky_irradiance";
  static const char kScatTables[] =
      "transmittance_texture, scattering_texture, single_mie_scattering_texture";
  static const char kIrrTables[] = "transmittance_texture, irradiance_texture";
  static const ApiEntry kApi[] = {
      {"RadianceSpectrum", "GetSkyRadiance", kSkyParams, "GetSkyRadiance", kScatTables,
       "camera, view_ray, shadow_length, sun_direction, transmittance", "", "", true},
      {"RadianceSpectrum", "GetSkyRadianceToPoint", kPointParams, "GetSkyRadianceToPoint", kScatTables,
       "camera, point, shadow_length, sun_direction, transmittance", "", "", true},
      {"IrradianceSpectrum", "GetSunAndSkyIrradiance", kIrrParams, "GetSunAndSkyIrradiance", kIrrTables,
       "p, normal, sun_direction, sky_irradiance", "", "", true},
      {"Luminance3", "GetSkyLuminance", kSkyParams, "GetSkyRadiance", kScatTables,
       "camera, view_ray, shadow_length, sun_direction, transmittance", "",
       "SKY_SPECTRAL_RADIANCE_TO_LUMINANCE", false},
      {"Luminance3", "GetSkyLuminanceToPoint", kPointParams, "GetSkyRadianceToPoint", kScatTables,
       "camera, point, shadow_length, sun_direction, transmittance", "",
       "SKY_SPECTRAL_RADIANCE_TO_LUMINANCE", false},
      {"Illuminance3", "GetSunAndSkyIlluminance", kIrrParams, "GetSunAndSkyIrradiance", kIrrTables,
       "p, normal, sun_direction, sky_irradiance", "SKY_SPECTRAL_RADIANCE_TO_LUMINANCE",
       "SUN_SPECTRAL_RADIANCE_TO_LUMINANCE", false},
  };
  static const char kSolar[] =
      "ATMOSPHERE.solar_irradiance / (PI * ATMOSPHERE.sun_angular_radius * ATMOSPHERE.sun_angular_radius)";
  std::string radiance, luminance;
  radiance += std::string("RadianceSpectrum GetSolarRadiance() {\n  return ") + kSolar + ";\n}\n";
  luminance += std::string("Luminance3 GetSolarLuminance() {\n  return ") + kSolar +
               " * SUN_SPECTRAL_RADIANCE_TO_LUMINANCE;\n}\n";
  for (const ApiEntry& e : kApi) {
    std::string f = std::string(e.result) + " " + e.name + "(" + e.params + ") {\n";
    f += std::string("  ") + e.result + " result = " + e.callee + "(ATMOSPHERE, " + e.tables + ", " +
         e.args + ")" + (e.scale[0] ? std::string(" * ") + e.scale : std::string()) + ";\n";
    if (e.out_scale[0]) f += std::string("  sky_irradiance *= ") + e.out_scale + ";\n";
    f += "  return result;\n}\n";
    (e.radiance_only ? radiance : luminance) += f;
  }
  return "uniform sampler2D transmittance_texture;\nuniform sampler3D scattering_texture;\n"
         "uniform sampler3D single_mie_scattering_texture;\nuniform sampler2D irradiance_texture;\n"
         "#ifdef RADIANCE_API_ENABLED\n" + radiance + "#endif\n" + luminance;
}

}  // namespace

namespace {
pas_status shader_source(const pas_model* m, const char* glsl_directory, std::string* out) {
  std::string definitions, functions;
  const std::string dir(glsl_directory);
  if (!read_text(dir + "/definitions.glsl", &definitions) || !read_text(dir + "/functions.glsl", &functions)) {
    return fail(PAS_ERR_IO, "cannot read definitions.glsl / functions.glsl in " + dir);
  }
  const PasGeometry& g = m->geom;
  std::string src = "#version 330\n#define IN(x) const in x\n#define OUT(x) out x\n"
                    "#define TEMPLATE(x)\n#define TEMPLATE_ARGUMENT(x)\n#define assert(x)\n";
  const std::pair<const char*, int> kSizes[] = {
      {"TRANSMITTANCE_TEXTURE_WIDTH", g.sz.t_w}, {"TRANSMITTANCE_TEXTURE_HEIGHT", g.sz.t_h},
      {"SCATTERING_TEXTURE_R_SIZE", g.sz.r_n}, {"SCATTERING_TEXTURE_MU_SIZE", g.sz.mu_n},
      {"SCATTERING_TEXTURE_MU_S_SIZE", g.sz.mu_s_n}, {"SCATTERING_TEXTURE_NU_SIZE", g.sz.nu_n},
      {"IRRADIANCE_TEXTURE_WIDTH", g.sz.e_w}, {"IRRADIANCE_TEXTURE_HEIGHT", g.sz.e_h}};
  for (const auto& kv : kSizes) {
    src += std::string("const int ") + kv.first + " = " + std::to_string(kv.second) + ";\n";
  }
  if (m->combined) src += "#define COMBINED_SCATTERING_TEXTURES\n";
  src += definitions;
  // field order of AtmosphereParameters (atmosphere/definitions.glsl:213-255)
  src += "const AtmosphereParameters ATMOSPHERE = AtmosphereParameters(\n";
  src += glsl_vec3(*m, m->solar, 1.0) + ",\n" + glsl_number(m->sun_angular_radius) + ",\n" +
         glsl_number(g.bottom) + ",\n" + glsl_number(g.top) + ",\n" + glsl_profile(g, 0) + ",\n" +
         glsl_vec3(*m, m->rayleigh, m->unit) + ",\n" + glsl_profile(g, 1) + ",\n" +
         glsl_vec3(*m, m->mie_sca, m->unit) + ",\n" + glsl_vec3(*m, m->mie_ext, m->unit) + ",\n" +
         glsl_number(m->mie_g) + ",\n" + glsl_profile(g, 2) + ",\n" +
         glsl_vec3(*m, m->absorption, m->unit) + ",\n" + glsl_vec3(*m, m->albedo, 1.0) + ",\n" +
         glsl_number(g.mu_s_min) + ");\n";
  auto vec3_of = [](const double* k) {
    return "vec3(" + glsl_number(k[0]) + "," + glsl_number(k[1]) + "," + glsl_number(k[2]) + ")";
  };
  src += "const vec3 SKY_SPECTRAL_RADIANCE_TO_LUMINANCE = " + vec3_of(m->sky_k) + ";\n";
  src += "const vec3 SUN_SPECTRAL_RADIANCE_TO_LUMINANCE = " + vec3_of(m->sun_k) + ";\n";
  src += functions;
  if (m->num_precomputed_wavelengths <= 3) src += "#define RADIANCE_API_ENABLED\n";
  src += api_wrappers();
  *out = std::move(src);
  return PAS_OK;
}

pas_status copy_out(const std::string& src, char* buffer, size_t* size) {
  const size_t need = src.size() + 1;
  if (buffer == nullptr) {
    *size = need;
    return PAS_OK;
  }
  if (*size < need) return fail(PAS_ERR_INVALID_ARGUMENT, "buffer too small");
  std::memcpy(buffer, src.c_str(), need);
  *size = need;
  return PAS_OK;
}

pas_status write_text(const std::string& path, const std::string& text) {
  std::ofstream out(path);
  out << text;
  out.close();
  return out ? PAS_OK : fail(PAS_ERR_IO, "cannot write " + path);
}
}  // namespace

extern "C" {

pas_status pas_model_shader_source(const pas_model* m, const char* glsl_directory, char* buffer, size_t* size) {
  if (m == nullptr || glsl_directory == nullptr || size == nullptr) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  }
  std::string src;
  pas_status st = shader_source(m, glsl_directory, &src);
  return st != PAS_OK ? st : copy_out(src, buffer, size);
}

pas_status pas_shader_source(const pas_model_params* params, const char* glsl_directory, char* buffer,
                             size_t* size) {
  if (glsl_directory == nullptr || size == nullptr) return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  std::unique_ptr<pas_model> m;
  pas_status st = build_host_model(params, &m);
  if (st != PAS_OK) return st;
  std::string src;
  st = shader_source(m.get(), glsl_directory, &src);
  return st != PAS_OK ? st : copy_out(src, buffer, size);
}

pas_status pas_model_save_webgl(pas_model* m, const char* directory, const char* glsl_directory,
                                const char* vertex_shader_source, const char* fragment_shader_source) {
  if (m == nullptr || directory == nullptr || glsl_directory == nullptr) {
    return fail(PAS_ERR_INVALID_ARGUMENT, "NULL argument");
  }
  pas_status st = pas_model_save_dat(m, directory);
  if (st != PAS_OK) return st;
  std::string src;
  if ((st = shader_source(m, glsl_directory, &src)) != PAS_OK) return st;
  const std::string dir(directory);
  if ((st = write_text(dir + "/atmosphere_shader.txt", src)) != PAS_OK) return st;
  if (vertex_shader_source != nullptr &&
      (st = write_text(dir + "/vertex_shader.txt", vertex_shader_source)) != PAS_OK) return st;
  if (fragment_shader_source != nullptr &&
      (st = write_text(dir + "/fragment_shader.txt", fragment_shader_source)) != PAS_OK) return st;
  return PAS_OK;
}

}  // extern "C"
