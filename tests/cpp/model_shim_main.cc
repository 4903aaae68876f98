// Exercises the C++ drop-in atmosphere::Model (include/atmosphere_b200/model.h) the way the
// reference's callers do (atmosphere/demo/demo.cc:278-284, reference/model_test.cc:392-414):
// construct with the 19 arguments, Init(), then use the tables. Headless build (no GL): the tables
// are written as .dat files (demo/webgl/precompute.cc:85-106) into argv[1].
//
//   g++ -std=c++14 -Iinclude/atmosphere_b200 tests/cpp/model_shim_main.cc -L<pkg> -lpas_b200
//
// Input spectra come from a small text file written by the test (argv[2]): one line per
// wavelength "lambda solar rayleigh mie_sca mie_ext absorption albedo".
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "model.h"

int main(int argc, char** argv) {
  if (argc < 4) {
    std::cerr << "usage: model_shim_main <out_dir> <spectra.txt> <num_wavelengths> [orders]\n";
    return 2;
  }
  std::vector<double> wavelengths, solar, rayleigh, mie_sca, mie_ext, absorption, albedo;
  std::ifstream in(argv[2]);
  double v[7];
  while (in >> v[0] >> v[1] >> v[2] >> v[3] >> v[4] >> v[5] >> v[6]) {
    wavelengths.push_back(v[0]); solar.push_back(v[1]); rayleigh.push_back(v[2]);
    mie_sca.push_back(v[3]); mie_ext.push_back(v[4]); absorption.push_back(v[5]); albedo.push_back(v[6]);
  }
  const unsigned n = static_cast<unsigned>(std::atoi(argv[3]));
  const unsigned orders = argc > 4 ? static_cast<unsigned>(std::atoi(argv[4])) : 4;
  using atmosphere::DensityProfileLayer;
  try {
    // Earth, demo parameters (demo.cc:226-250), half precision + combined textures
    atmosphere::Model model(
        wavelengths, solar, 0.00935 / 2.0, 6360000.0, 6420000.0,
        {DensityProfileLayer(0.0, 1.0, -1.0 / 8000.0, 0.0, 0.0)}, rayleigh,
        {DensityProfileLayer(0.0, 1.0, -1.0 / 1200.0, 0.0, 0.0)}, mie_sca, mie_ext, 0.8,
        {DensityProfileLayer(25000.0, 0.0, 0.0, 1.0 / 15000.0, -2.0 / 3.0),
         DensityProfileLayer(0.0, 0.0, 0.0, -1.0 / 15000.0, 8.0 / 3.0)},
        absorption, albedo, 102.0 / 180.0 * 3.1415926, 1000.0, n, true, true);
    model.Init(orders);
    model.SetProgramUniforms(0, 0, 1, 2);  // no-op without GL; must still link
    model.SaveDat(argv[1]);
    const pas_texture_info s = model.TextureInfo(PAS_TEXTURE_SCATTERING);
    const std::vector<float> t = model.ReadTexture(PAS_TEXTURE_TRANSMITTANCE);
    double r, g, b;
    atmosphere::Model::ConvertSpectrumToLinearSrgb(wavelengths, solar, &r, &g, &b);
    std::printf("scattering %dx%dx%d bytes_per_channel %d T[0]=%.9g sun_srgb %.6f %.6f %.6f shader %u\n",
                s.width, s.height, s.depth, s.bytes_per_channel, t[0], r, g, b, model.shader());
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
