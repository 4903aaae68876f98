// Forwarding header: put this directory first on the include path and the reference's callers
// (#include "atmosphere/model.h": demo/demo.cc, reference/model_test.cc, demo/webgl/precompute.cc)
// pick up the B200 drop-in atmosphere::Model instead of atmosphere/model.h. See INTEGRATION.md.
#include "../../include/atmosphere_b200/model.h"
