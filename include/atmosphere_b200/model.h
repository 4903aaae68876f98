// Drop-in C++ counterpart of the reference's atmosphere::Model (atmosphere/model.h:165-337) whose
// Init() runs the B200 CUDA precomputation behind the C ABI of include/pas_b200.h instead of the
// OpenGL fragment shaders of atmosphere/model.cc:1048-1215.
//
// Same namespace, class names, constructor arguments (model.h:182-281), Init(num_scattering_orders)
// (model.h:285), shader() (model.h:287), SetProgramUniforms (model.h:289-294),
// ConvertSpectrumToLinearSrgb (model.h:301-304) and kLambdaR/G/B (model.h:306-308), so that
// atmosphere/demo/demo.cc and the integration test compile against this header unchanged:
//
//   g++ -DPAS_WITH_GL -Iinclude/atmosphere_b200 -Iexternal/glad/include demo.cc ... -lpas_b200
//
// Header only; link with libpas_b200.so. Two build modes:
//   * PAS_WITH_GL defined (needs <glad/glad.h> and a current GL 3.3 context, like the reference):
//     Init() uploads the tables computed on the GPU into GL textures of the reference's formats
//     (model.cc:422-456, 747-766), shader() compiles GetShaderSource(), SetProgramUniforms binds the
//     textures to the uniforms of model.cc:992-1008.
//   * otherwise (headless: servers, tests): GLuint is a plain unsigned int, shader() returns 0 and
//     SetProgramUniforms does nothing; the tables are read with ReadTexture() / SaveDat().
// In both modes failures throw std::runtime_error with the library's message (the reference
// asserts, model.cc:380,400-401,974); there is no CPU fallback.
#ifndef ATMOSPHERE_B200_MODEL_H_
#define ATMOSPHERE_B200_MODEL_H_

#ifdef PAS_WITH_GL
#include <glad/glad.h>
#else
typedef unsigned int GLuint;
#endif

#include <stdexcept>
#include <string>
#include <vector>

#include "../pas_b200.h"

namespace atmosphere {

// atmosphere/model.h:165-178.
class DensityProfileLayer {
 public:
  DensityProfileLayer() : DensityProfileLayer(0.0, 0.0, 0.0, 0.0, 0.0) {}
  DensityProfileLayer(double width, double exp_term, double exp_scale, double linear_term,
                      double constant_term)
      : width(width), exp_term(exp_term), exp_scale(exp_scale), linear_term(linear_term),
        constant_term(constant_term) {}
  double width, exp_term, exp_scale, linear_term, constant_term;
};

class Model {
 public:
  Model(const std::vector<double>& wavelengths, const std::vector<double>& solar_irradiance,
        double sun_angular_radius, double bottom_radius, double top_radius,
        const std::vector<DensityProfileLayer>& rayleigh_density,
        const std::vector<double>& rayleigh_scattering,
        const std::vector<DensityProfileLayer>& mie_density,
        const std::vector<double>& mie_scattering, const std::vector<double>& mie_extinction,
        double mie_phase_function_g, const std::vector<DensityProfileLayer>& absorption_density,
        const std::vector<double>& absorption_extinction, const std::vector<double>& ground_albedo,
        double max_sun_zenith_angle, double length_unit_in_meters,
        unsigned int num_precomputed_wavelengths, bool combine_scattering_textures,
        bool half_precision, const std::string& glsl_directory = "atmosphere")
      : glsl_directory_(glsl_directory) {
    const size_t n = wavelengths.size();
    for (const std::vector<double>* v : {&solar_irradiance, &rayleigh_scattering, &mie_scattering,
                                         &mie_extinction, &absorption_extinction, &ground_albedo}) {
      if (v->size() != n) throw std::invalid_argument("one value per wavelength expected");  // model.cc:539
    }
    auto layers = [](const std::vector<DensityProfileLayer>& in) {
      std::vector<pas_density_layer> out;
      for (const DensityProfileLayer& l : in) {
        out.push_back({l.width, l.exp_term, l.exp_scale, l.linear_term, l.constant_term});
      }
      return out;
    };
    const std::vector<pas_density_layer> ray = layers(rayleigh_density), mie = layers(mie_density),
                                         absorb = layers(absorption_density);
    pas_model_params p{};
    p.num_wavelengths = n;
    p.wavelengths = wavelengths.data();
    p.solar_irradiance = solar_irradiance.data();
    p.sun_angular_radius = sun_angular_radius;
    p.bottom_radius = bottom_radius;
    p.top_radius = top_radius;
    p.num_rayleigh_layers = ray.size();
    p.rayleigh_density = ray.data();
    p.rayleigh_scattering = rayleigh_scattering.data();
    p.num_mie_layers = mie.size();
    p.mie_density = mie.data();
    p.mie_scattering = mie_scattering.data();
    p.mie_extinction = mie_extinction.data();
    p.mie_phase_function_g = mie_phase_function_g;
    p.num_absorption_layers = absorb.size();
    p.absorption_density = absorb.data();
    p.absorption_extinction = absorption_extinction.data();
    p.ground_albedo = ground_albedo.data();
    p.max_sun_zenith_angle = max_sun_zenith_angle;
    p.length_unit_in_meters = length_unit_in_meters;
    p.num_precomputed_wavelengths = num_precomputed_wavelengths;
    p.combine_scattering_textures = combine_scattering_textures ? 1 : 0;
    p.half_precision = half_precision ? 1 : 0;
    Check(pas_model_create(&p, &model_));
#ifdef PAS_WITH_GL
    // like the reference (model.cc:769-776) the shader is compiled here: shader() is valid before Init
    CompileShader();
#endif
  }

  Model(const Model&) = delete;             // owns device memory and GL objects
  Model& operator=(const Model&) = delete;

  ~Model() {
#ifdef PAS_WITH_GL
    if (atmosphere_shader_ != 0) glDeleteShader(atmosphere_shader_);
    glDeleteTextures(4, textures_);
#endif
    pas_model_destroy(model_);
  }

  // atmosphere/model.cc:866-975. Blocks until every table is complete.
  void Init(unsigned int num_scattering_orders = 4) {
    Check(pas_model_init(model_, num_scattering_orders));
#ifdef PAS_WITH_GL
    UploadTextures();
#endif
  }

  GLuint shader() const { return atmosphere_shader_; }

  // atmosphere/model.cc:984-1011: same uniform names, same texture units.
  void SetProgramUniforms(GLuint program, GLuint transmittance_texture_unit,
                          GLuint scattering_texture_unit, GLuint irradiance_texture_unit,
                          GLuint optional_single_mie_scattering_texture_unit = 0) const {
#ifdef PAS_WITH_GL
    Bind(program, "transmittance_texture", GL_TEXTURE_2D, transmittance_texture_unit,
         textures_[PAS_TEXTURE_TRANSMITTANCE]);
    Bind(program, "scattering_texture", GL_TEXTURE_3D, scattering_texture_unit,
         textures_[PAS_TEXTURE_SCATTERING]);
    Bind(program, "irradiance_texture", GL_TEXTURE_2D, irradiance_texture_unit,
         textures_[PAS_TEXTURE_IRRADIANCE]);
    if (textures_[PAS_TEXTURE_SINGLE_MIE] != 0) {
      Bind(program, "single_mie_scattering_texture", GL_TEXTURE_3D,
           optional_single_mie_scattering_texture_unit, textures_[PAS_TEXTURE_SINGLE_MIE]);
    }
#else
    (void)program; (void)transmittance_texture_unit; (void)scattering_texture_unit;
    (void)irradiance_texture_unit; (void)optional_single_mie_scattering_texture_unit;
#endif
  }

  // atmosphere/model.cc:1020-1040.
  static void ConvertSpectrumToLinearSrgb(const std::vector<double>& wavelengths,
                                          const std::vector<double>& spectrum, double* r, double* g,
                                          double* b) {
    if (pas_convert_spectrum_to_linear_srgb(wavelengths.size(), wavelengths.data(), spectrum.data(),
                                            r, g, b) != PAS_OK) {
      throw std::runtime_error(pas_last_error());
    }
  }

  static constexpr double kLambdaR = 680.0;
  static constexpr double kLambdaG = 550.0;
  static constexpr double kLambdaB = 440.0;

  // ---- headless access (no counterpart in the reference, which keeps the tables in GL) ----------
  // The GLSL the reference's shader() compiles (model.cc:691-744, 769-772).
  std::string GetShaderSource() const {
    size_t size = 0;
    Check(pas_model_shader_source(model_, glsl_directory_.c_str(), nullptr, &size));
    std::string text(size, '\0');
    Check(pas_model_shader_source(model_, glsl_directory_.c_str(), &text[0], &size));
    text.resize(size - 1);
    return text;
  }
  pas_texture_info TextureInfo(pas_texture which) const {
    pas_texture_info info;
    Check(pas_model_texture_info(model_, which, &info));
    return info;
  }
  // RGBA float32 texels, x fastest: what glGetTexImage(GL_RGBA, GL_FLOAT) returns for the
  // reference's textures (demo/webgl/precompute.cc:63-73).
  std::vector<float> ReadTexture(pas_texture which) const {
    const pas_texture_info info = TextureInfo(which);
    std::vector<float> texels((size_t)info.width * info.height * info.depth * 4);
    Check(pas_model_read_texture(model_, which, 1, texels.data(), texels.size() * sizeof(float)));
    return texels;
  }
  void SaveDat(const std::string& directory) const { Check(pas_model_save_dat(model_, directory.c_str())); }
  // demo/webgl/precompute.cc:81-106: the .dat files + atmosphere_shader.txt + the caller's own shaders
  void SaveWebGl(const std::string& directory, const std::string& vertex_shader,
                 const std::string& fragment_shader) const {
    Check(pas_model_save_webgl(model_, directory.c_str(), glsl_directory_.c_str(), vertex_shader.c_str(),
                               fragment_shader.c_str()));
  }
  pas_model* handle() const { return model_; }

 private:
  static void Check(pas_status status) {
    if (status != PAS_OK) throw std::runtime_error(pas_last_error());
  }

#ifdef PAS_WITH_GL
  static void Bind(GLuint program, const char* name, GLenum target, GLuint unit, GLuint texture) {
    glActiveTexture(GL_TEXTURE0 + unit);
    glBindTexture(target, texture);
    glUniform1i(glGetUniformLocation(program, name), unit);
  }
  void UploadTextures() {
    for (int which = 0; which < 4; ++which) {
      const pas_texture_info info = TextureInfo(static_cast<pas_texture>(which));
      if (!info.present) continue;
      const bool half = info.bytes_per_channel == 2;
      const bool is3d = info.depth > 1;
      const GLenum target = is3d ? GL_TEXTURE_3D : GL_TEXTURE_2D;
      std::vector<unsigned char> texels((size_t)info.width * info.height * info.depth * 4 *
                                        info.bytes_per_channel);
      Check(pas_model_read_texture(model_, static_cast<pas_texture>(which), 0, texels.data(),
                                   texels.size()));
      if (textures_[which] == 0) glGenTextures(1, &textures_[which]);
      glBindTexture(target, textures_[which]);
      for (GLenum wrap : {GL_TEXTURE_WRAP_S, GL_TEXTURE_WRAP_T, GL_TEXTURE_WRAP_R}) {
        glTexParameteri(target, wrap, GL_CLAMP_TO_EDGE);  // model.cc:425-431, 443-450
      }
      glTexParameteri(target, GL_TEXTURE_MIN_FILTER, GL_LINEAR);
      glTexParameteri(target, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
      glPixelStorei(GL_UNPACK_ALIGNMENT, 1);
      const GLenum internal = half ? GL_RGBA16F : GL_RGBA32F;
      const GLenum type = half ? GL_HALF_FLOAT : GL_FLOAT;
      if (is3d) {
        glTexImage3D(target, 0, internal, info.width, info.height, info.depth, 0, GL_RGBA, type,
                     texels.data());
      } else {
        glTexImage2D(target, 0, internal, info.width, info.height, 0, GL_RGBA, type, texels.data());
      }
    }
  }
  void CompileShader() {
    const std::string source = GetShaderSource();
    const char* text = source.c_str();
    atmosphere_shader_ = glCreateShader(GL_FRAGMENT_SHADER);
    glShaderSource(atmosphere_shader_, 1, &text, nullptr);
    glCompileShader(atmosphere_shader_);
    GLint ok = GL_FALSE;
    glGetShaderiv(atmosphere_shader_, GL_COMPILE_STATUS, &ok);
    if (ok != GL_TRUE) throw std::runtime_error("atmosphere shader does not compile");
  }
#endif

  std::string glsl_directory_;
  pas_model* model_ = nullptr;
  GLuint atmosphere_shader_ = 0;
  GLuint textures_[4] = {0, 0, 0, 0};
};

}  // namespace atmosphere

#endif  // ATMOSPHERE_B200_MODEL_H_
