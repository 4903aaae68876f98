/* pas_b200 -- C ABI of the B200-native LUT precomputation for
 * ebruneton/precomputed_atmospheric_scattering.
 *
 * This is the drop-in boundary for ONE path of the reference: atmosphere::Model's constructor +
 * Model::Init (atmosphere/model.cc:613-795, 866-975, 1048-1215), which fill the transmittance,
 * scattering (+ optional single Mie) and irradiance tables. Every entry point names the reference
 * interface it replaces. Plain C types only: pointers, sizes, doubles. All functions return a
 * pas_status (0 = PAS_OK); pas_last_error() returns a human-readable message for the calling
 * thread's last failure. Nothing here ever falls back to a CPU implementation: without a CUDA
 * device pas_model_create fails with PAS_ERR_CUDA.
 *
 * Threading: calls on different pas_model handles are independent; one handle must not be used
 * from two threads at once (the reference is single-threaded on its GL context,
 * atmosphere/model.cc:866-975).
 */
#ifndef PAS_B200_H_
#define PAS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAS_B200_ABI_VERSION 1

typedef enum pas_status {
  PAS_OK = 0,
  PAS_ERR_INVALID_ARGUMENT = 1,  /* bad sizes / unsorted wavelengths / top <= bottom / > 2 layers */
  PAS_ERR_CUDA = 2,              /* CUDA runtime failure (message has the CUDA error string) */
  PAS_ERR_NCCL = 3,              /* NCCL failure on the multi-GPU path */
  PAS_ERR_UNSUPPORTED = 4,       /* table sizes outside what the kernels support */
  PAS_ERR_STATE = 5,             /* e.g. reading a table before pas_model_init */
  PAS_ERR_IO = 6                 /* file / shader-source I/O */
} pas_status;

/* atmosphere::DensityProfileLayer (atmosphere/model.h:165-178), SI units. */
typedef struct pas_density_layer {
  double width, exp_term, exp_scale, linear_term, constant_term;
} pas_density_layer;

/* Table sizes. The reference fixes them at compile time (atmosphere/constants.h:47-61); here they
 * are run-time. Any field left 0 takes the reference's value
 * (256x64, R=32, MU=128, MU_S=32, NU=8, 64x16). */
typedef struct pas_lut_sizes {
  int transmittance_width, transmittance_height;
  int scattering_r, scattering_mu, scattering_mu_s, scattering_nu;
  int irradiance_width, irradiance_height;
} pas_lut_sizes;

/* The 19 arguments of atmosphere::Model::Model (atmosphere/model.h:182-281), same meaning, same
 * units (nm, W/m^2/nm, rad, m, m^-1), same rules: spectra are sampled at `wavelengths` and
 * interpolated linearly, clamped at the ends (atmosphere/model.cc:535-552); density profiles have
 * at most 2 layers, missing ones are zero layers below (atmosphere/model.cc:653-666). */
typedef struct pas_model_params {
  size_t num_wavelengths;
  const double* wavelengths;
  const double* solar_irradiance;
  double sun_angular_radius;
  double bottom_radius;
  double top_radius;
  size_t num_rayleigh_layers;
  const pas_density_layer* rayleigh_density;
  const double* rayleigh_scattering;
  size_t num_mie_layers;
  const pas_density_layer* mie_density;
  const double* mie_scattering;
  const double* mie_extinction;
  double mie_phase_function_g;
  size_t num_absorption_layers;
  const pas_density_layer* absorption_density;
  const double* absorption_extinction;
  const double* ground_albedo;
  double max_sun_zenith_angle;
  double length_unit_in_meters;
  unsigned int num_precomputed_wavelengths;
  int combine_scattering_textures;
  int half_precision;
  /* ---- extensions (zero-initialise for reference behaviour) ---- */
  pas_lut_sizes sizes;
  int device;  /* CUDA device ordinal + 1; 0 = the calling thread's current device */
} pas_model_params;

typedef struct pas_model pas_model;

/* The tables a Model owns (atmosphere/model.h:331-334). */
typedef enum pas_texture {
  PAS_TEXTURE_TRANSMITTANCE = 0,    /* RGBA32F, width x height                        */
  PAS_TEXTURE_SCATTERING = 1,       /* RGBA16F/32F, (NU*MU_S) x MU x R                 */
  PAS_TEXTURE_IRRADIANCE = 2,       /* RGBA32F, width x height                        */
  PAS_TEXTURE_SINGLE_MIE = 3        /* RGBA16F/32F; absent with combined textures      */
} pas_texture;

typedef struct pas_texture_info {
  int width, height, depth;   /* depth 1 for 2-D tables */
  int channels;               /* always 4: texels are RGBA interleaved, x fastest, then y, z --
                                 what glGetTexImage(GL_RGBA, ...) returns
                                 (atmosphere/demo/webgl/precompute.cc:63-73) */
  int bytes_per_channel;      /* 4 (float) or 2 (IEEE half) as stored on the device */
  int present;                /* 0 if this model has no such table */
} pas_texture_info;

/* Replaces atmosphere::Model::Model (atmosphere/model.cc:613-795): validates and converts the
 * parameters, allocates the tables in HBM. Does not precompute. */
pas_status pas_model_create(const pas_model_params* params, pas_model** out_model);

/* Replaces atmosphere::Model::~Model (atmosphere/model.cc:801-811). NULL is allowed. */
void pas_model_destroy(pas_model* model);

/* Replaces atmosphere::Model::Init(num_scattering_orders) (atmosphere/model.cc:866-975 and
 * Precompute, :1048-1215): runs every pass on the GPU and returns when the tables are complete in
 * device memory. May be called again (e.g. with another order count). */
pas_status pas_model_init(pas_model* model, unsigned int num_scattering_orders);
/* The same work without blocking the host: pas_model_init_async enqueues the whole precomputation on
 * the model's own streams and returns; pas_model_wait blocks until it is done (pas_model_init is the
 * two calls back to back). Several models (a batch of atmospheres: turbidity / ozone / albedo
 * sweeps) can be in flight at once; their kernels share the GPU. Every other call on a model waits
 * for its Init first. */
pas_status pas_model_init_async(pas_model* model, unsigned int num_scattering_orders);
pas_status pas_model_wait(pas_model* model);
/* Pipelined read-back for callers that upload the tables right after Init (the GL upload of
 * atmosphere/model.cc:747-766, the .dat export of demo/webgl/precompute.cc:85-106): registers host
 * destinations (NULL = not wanted; native texel format and size of pas_model_texture_info, i.e. what
 * pas_model_read_texture(as_float32 = 0) returns; page-locked memory for the copies to overlap).
 * Every later Init copies each table out as soon as its last writer is done -- T after the first
 * pass -- and the last multiple-scattering pass writes the scattering table into a page-locked
 * destination itself, row by row as it finishes them (a pageable destination is copied in bands of
 * layers behind that pass). The buffers are valid when pas_model_init / pas_model_wait returns. Pass
 * four NULLs to unregister. */
pas_status pas_model_set_host_outputs(pas_model* model, void* transmittance, void* scattering,
                                      void* single_mie, void* irradiance);
/* Multi-GPU worlds over peer / symmetric memory: with own_layers_only != 0 a rank copies out only the
 * scattering layers it computed (rank 0 also the transmittance and irradiance tables), pipelined behind
 * its passes like on one GPU, and the last cross-rank barrier of Init comes after those copies. The
 * registered buffers are then meant to be ONE set of host tables shared by the ranks of the box (POSIX
 * shared memory registered with cudaHostRegister by every rank, world.py: shared_host_tables): when
 * Init returns on any rank, every layer of every table is in it. Default 0: every rank copies out
 * complete tables after the final exchange. */
pas_status pas_model_set_host_output_mode(pas_model* model, int own_layers_only);

pas_status pas_model_texture_info(const pas_model* model, pas_texture which,
                                  pas_texture_info* info);

/* Device pointer of a table, for zero-copy consumers (CUDA-GL interop upload in the Model shim's
 * SetProgramUniforms, atmosphere/model.cc:984-1011). Valid until the next init/destroy. */
pas_status pas_model_texture_device_ptr(const pas_model* model, pas_texture which,
                                        const void** device_ptr);

/* Copies a table to host memory as RGBA float32 (as_float32 != 0: what
 * glGetTexImage(GL_RGBA, GL_FLOAT) returns, the layout of the reference's .dat files,
 * atmosphere/demo/webgl/precompute.cc:63-73) or in its stored precision. dst_bytes must match. */
pas_status pas_model_read_texture(pas_model* model, pas_texture which, int as_float32, void* dst,
                                  size_t dst_bytes);

/* Writes transmittance.dat / scattering.dat / irradiance.dat (+ single_mie_scattering.dat) into
 * `directory`, raw little-endian RGBA32F (atmosphere/demo/webgl/precompute.cc:85-106). */
pas_status pas_model_save_dat(pas_model* model, const char* directory);

/* ---- render-time use of the tables --------------------------------------
 * CUDA counterparts of the GLSL rendering API that atmosphere::Model::shader() exports
 * (atmosphere/model.cc:221-281) and of the CPU API atmosphere::reference::Model
 * (atmosphere/reference/model.h:62-77): GetSolarRadiance, GetSkyRadiance, GetSkyRadianceToPoint,
 * GetSunAndSkyIrradiance and their *Luminance / *Illuminance forms. Positions are relative to the
 * planet centre, in the model's length unit; directions are unit vectors; results are RGB at
 * 680/550/440 nm (radiance, W/m^2/sr/nm) or linear-sRGB luminance (cd/m^2, use_luminance != 0).
 * As in the reference (atmosphere/model.h:120-144) the radiance forms exist only for models with
 * num_precomputed_wavelengths <= 3 (PAS_ERR_STATE otherwise). Vector arguments are arrays of n
 * xyz triples of doubles, results arrays of n rgb triples of floats; every pointer may be a host
 * or a device pointer (detected per pointer). pas_model_init must have run. */
pas_status pas_model_get_solar_radiance(const pas_model* model, int use_luminance, double* rgb);
/* functions.glsl:1705-1769. shadow_length (n doubles) and transmittance may be NULL. */
pas_status pas_model_get_sky_radiance(pas_model* model, int use_luminance, size_t n,
                                      const double* camera, const double* view_ray,
                                      const double* shadow_length, const double* sun_direction,
                                      float* radiance, float* transmittance);
/* functions.glsl:1787-1863. */
pas_status pas_model_get_sky_radiance_to_point(pas_model* model, int use_luminance, size_t n,
                                               const double* camera, const double* point,
                                               const double* shadow_length,
                                               const double* sun_direction, float* radiance,
                                               float* transmittance);
/* functions.glsl:1878-1896 (model.cc:272-280 in luminance mode). */
pas_status pas_model_get_sun_and_sky_irradiance(pas_model* model, int use_luminance, size_t n,
                                                const double* point, const double* normal,
                                                const double* sun_direction, float* sun_irradiance,
                                                float* sky_irradiance);

/* For CUDA renderers that call the lookups from their own kernels: fills `out` with the
 * pas::RenderContext of csrc/kernel_render.cuh (geometry, device pointers of the four product tables,
 * the constants at 680/550/440 nm, luminance factors or 1) -- the CUDA counterpart of the reference
 * handing its users GLSL source + sampler uniforms (atmosphere/model.cc:221-281, 984-1011). The context
 * is valid until the next pas_model_init / destroy of the model. Call with out == NULL to get the size.
 * (The reference's integration-test scene, reference/model_test.glsl, is rendered this way by the tests:
 * tests/cuda/scene_kernel.cu. It is test infrastructure, not part of this library.) */
pas_status pas_model_render_context(pas_model* model, int use_luminance, void* out, size_t* bytes);
/* Device time of the last render / lookup kernel in milliseconds (CUDA events on the model's
 * stream, copies excluded). */
pas_status pas_model_last_render_ms(const pas_model* model, float* ms);

/* The GLSL source atmosphere::Model::shader() compiles (atmosphere/model.cc:691-744, 769-772):
 * header with the ATMOSPHERE constant + definitions.glsl + functions.glsl + the API wrappers.
 * definitions.glsl / functions.glsl are read from `glsl_directory` (the reference checkout's
 * atmosphere/ directory); they are inputs, not part of this library. Call with buffer == NULL to
 * get the required size (including the terminating NUL) in *size. */
pas_status pas_model_shader_source(const pas_model* model, const char* glsl_directory,
                                   char* buffer, size_t* size);
/* The same source from the constructor parameters alone -- no model, no device: the shader depends on
 * the parameters only, which is why the reference compiles it in its constructor, before Init
 * (atmosphere/model.cc:769-776). */
pas_status pas_shader_source(const pas_model_params* params, const char* glsl_directory, char* buffer,
                             size_t* size);
/* The hand-off of atmosphere/demo/webgl/precompute.cc:50-61, 81-106 to the WebGL viewer: the .dat
 * files of pas_model_save_dat plus atmosphere_shader.txt (= pas_model_shader_source) and, when the caller
 * passes them, vertex_shader.txt / fragment_shader.txt (the demo's own shaders: they belong to the
 * caller, demo/demo.cc:75-97 and demo/demo.glsl; NULL = not written). */
pas_status pas_model_save_webgl(pas_model* model, const char* directory, const char* glsl_directory,
                                const char* vertex_shader_source, const char* fragment_shader_source);

/* The SKY/SUN_SPECTRAL_RADIANCE_TO_LUMINANCE constants baked into that shader
 * (atmosphere/model.cc:562-595, 668-686). out6 = sky rgb, sun rgb. */
pas_status pas_model_luminance_factors(const pas_model* model, double* out6);

/* Replaces the static atmosphere::Model::ConvertSpectrumToLinearSrgb
 * (atmosphere/model.cc:1020-1040). */
pas_status pas_convert_spectrum_to_linear_srgb(size_t n, const double* wavelengths,
                                               const double* spectrum, double* r, double* g,
                                               double* b);

/* ---- introspection used by the parity tests and the bench (no reference counterpart) -------- */

/* Number of spectral channels the model precomputes and their wavelengths
 * (atmosphere/model.cc:907-924). Pass lambdas == NULL to query the count only. */
pas_status pas_model_channels(const pas_model* model, int* num_channels, double* lambdas);

/* luminance_from_radiance of all channels side by side, row-major [3][num_channels]
 * (atmosphere/model.cc:909, 925-943). */
pas_status pas_model_luminance_matrix(const pas_model* model, float* out);

/* The same two without a model (and without a GPU): the wavelength grid and the luminance matrices
 * Model::Init derives from num_precomputed_wavelengths (atmosphere/model.cc:907-943). lambdas and
 * luminance_from_radiance may be NULL to query *num_channels only. */
pas_status pas_spectral_channels(unsigned int num_precomputed_wavelengths, int* num_channels,
                                 double* lambdas, float* luminance_from_radiance);

/* When enabled, Init keeps a device copy of every intermediate (planar per channel, texel order
 * x fastest): "transmittance", "delta_irradiance_<n>", "delta_rayleigh", "delta_mie",
 * "delta_density_<n>", "delta_multiple_<n>". Costs memory and time; off by default. With captures
 * Init enqueues every pass on one stream in the reference's order; without, it overlaps the
 * irradiance passes with the multiple-scattering passes and reuses the delta_rayleigh buffer for
 * odd orders, so the live buffers below are only meaningful after single passes or captured runs
 * (the products are bit-identical either way). */
pas_status pas_model_set_capture(pas_model* model, int enabled);
/* num_floats: in = capacity of dst, out = floats needed/copied. dst may be NULL to query. */
pas_status pas_model_read_intermediate(pas_model* model, const char* name, float* dst,
                                       size_t* num_floats);

/* Teacher-forced single passes: upload planar per-channel inputs (same names as above, plus
 * "delta_irradiance" / "delta_density" / "delta_multiple" for the live buffers) and run one phase
 * of atmosphere/model.cc:1048-1215. phase: 0 transmittance, 1 direct irradiance, 2 single
 * scattering, 3 scattering density(order), 4 indirect irradiance(order = order of the radiance
 * integrated), 5 multiple scattering, 6 the per-(layer, direction) ground-transmittance tables of
 * the density pass alone, from the transmittance buffer as it stands (phase 0 computes both).
 * Results are read back with pas_model_read_intermediate using the live-buffer names. */
pas_status pas_model_write_intermediate(pas_model* model, const char* name, const float* src,
                                        size_t num_floats);
pas_status pas_model_run_phase(pas_model* model, int phase, int order);

/* Device time of the last pas_model_init, per phase (CUDA events on the model's stream), in ms.
 * names/ms are filled up to *count entries; *count returns the number available. */
pas_status pas_model_last_timings(const pas_model* model, int* count, const char** names,
                                  float* ms);
/* Kernels launched by the last pas_model_init. */
pas_status pas_model_last_launch_count(const pas_model* model, int* launches);

/* Roofline denominators of this path, measured on the device with saturating microbenchmarks (CUDA
 * events): dense FP32 FMA throughput in TFLOP/s and MUFU (SFU) throughput in Gop/s. Used by bench.py;
 * the driver's MEASURED_PEAKS.json only has HBM and tensor-core figures. */
pas_status pas_measure_device_peaks(int device, double* fp32_tflops, double* mufu_gops,
                                    int* sm_count);

/* ---- multi-GPU (one process per GPU; no reference counterpart) ------------------------------ */

#define PAS_NCCL_UNIQUE_ID_BYTES 128
/* Fills a 128-byte NCCL unique id (call on rank 0, broadcast it with any host-side transport). */
pas_status pas_nccl_unique_id(void* id_bytes);
/* Attaches the model to a world of `world_size` processes: scattering layers are split into
 * contiguous r-slabs, one per rank; the scattering-density table is all-gathered over NVLink and
 * the irradiance partial sums all-reduced between orders. Must be called by every rank before
 * pas_model_init. After Init every rank holds the complete final tables. */
pas_status pas_model_attach_world(pas_model* model, int rank, int world_size,
                                  const void* nccl_unique_id_bytes);
/* Communicators are kept per (device, rank, world size) for the life of the process: once one
 * exists, later models attach with nccl_unique_id_bytes == NULL. device = CUDA ordinal. */
int pas_world_is_cached(int device, int rank, int world_size);

/* Peer-memory exchange (preferred on one NVLink / NVSwitch box): instead of NCCL collectives, the
 * scattering-density kernel stores its r-slab straight into the tables of the other ranks, the
 * irradiance partial sums and the final scattering slabs are pushed the same way, and ranks meet at
 * flag barriers in device memory. Protocol, on every rank: pas_model_ipc_export fills
 * PAS_IPC_EXPORT_BYTES (CUDA IPC handles of the model's exchange buffers); the host all-gathers
 * them in rank order (any transport); pas_model_attach_peers maps the others' buffers. Ranks must
 * be separate processes on GPUs with peer access. No NCCL communicator is needed or created. */
#define PAS_IPC_EXPORT_BYTES 464
pas_status pas_model_ipc_export(pas_model* model, int rank, int world_size, void* out, size_t* bytes);
pas_status pas_model_attach_peers(pas_model* model, const void* all_exports, size_t bytes_per_rank);

/* Symmetric-memory worlds (preferred where the host can provide one): every rank owns an arena of
 * `bytes` bytes of device memory that every other rank of the box has mapped -- `arena_bases[r]` is rank
 * r's arena as seen from THIS process (arena_bases[rank] = the local one) -- and, where the NVSwitch
 * supports it, `multicast_base` is ONE address whose stores land in the arenas of all ranks at once
 * (NVLS multicast; NULL = unicast stores to arena_bases). The host side allocates and maps the arenas
 * (torch.distributed._symmetric_memory in world.py: CUDA VMM + fabric handles) ONCE per (rank, world)
 * and hands them to every model it creates afterwards: attaching needs no collective. The arena holds
 * the exchange copies of the model's tables (transmittance, two scattering-density buffers, irradiance
 * partial sums, staging of the final scattering slabs) and the flag words of the barriers; it must be
 * zero when first attached and at least pas_model_exchange_bytes() large. With multicast the density
 * kernel sends each texel of its r-slab once instead of once per peer. Models that share an arena
 * run their Inits one at a time (PAS_ERR_STATE otherwise); their product tables stay their own. */
pas_status pas_model_exchange_bytes(const pas_model* model, int world_size, size_t* bytes);
pas_status pas_model_attach_symmetric(pas_model* model, int rank, int world_size,
                                      void* const* arena_bases, void* multicast_base, size_t bytes);

/* Device memory of destroyed models is kept in a process-wide pool for the next pas_model_create
 * (the demo re-creates its Model on every settings change, atmosphere/demo/demo.cc:446-494);
 * this returns it to the driver. */
void pas_release_cached_memory(void);

const char* pas_last_error(void);
int pas_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif  /* PAS_B200_H_ */
