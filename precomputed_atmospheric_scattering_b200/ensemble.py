"""Batches of atmospheres (BASELINE config 5): a turbidity x ozone x ground-albedo sweep of the demo's
Earth, precomputed with several models in flight at once (pas_model_init_async: every model has its
own CUDA streams, their kernels share the GPU); the tables are then used through the render-time
lookups (Model.GetSkyRadiance ...), by the tests through the reference's test scene (tests/scene_render.py).

The reference has no batch API: its demo re-creates one Model per settings change
(atmosphere/demo/demo.cc:446-494). The sweep below varies what the demo's keys vary -- the Mie scale
height stands for turbidity (demo.cc:230-234), the ozone column (demo.cc:208-222, 272-274) and the
ground albedo (demo.cc:234) -- on a fixed, seeded grid so that tests and benches name the same 64
atmospheres.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .atmospheres import AtmosphereSpec, earth
from .model import Model

SWEEP_SEED = 0
MIE_SCALE_HEIGHTS_M = (800.0, 1200.0, 1800.0, 2700.0)     # turbidity: thinner ... thicker haze layer
OZONE_DOBSON = (200.0, 300.0, 400.0, 500.0)
GROUND_ALBEDO = (0.0, 0.1, 0.3, 0.6)


def sweep(n_turbidity: int = 4, n_ozone: int = 4, n_albedo: int = 4, *, seed: int = SWEEP_SEED,
          jitter: float = 0.05, **earth_kw) -> List[AtmosphereSpec]:
    """``n_turbidity * n_ozone * n_albedo`` Earth atmospheres (64 by default). Every grid value is
    scaled by a factor drawn once from U(1 - jitter, 1 + jitter) with ``numpy.random.default_rng(seed)``
    so that no two atmospheres share a parameter exactly; ``earth_kw`` goes to ``atmospheres.earth``
    (num_precomputed_wavelengths, half_precision, ...)."""
    if not (1 <= n_turbidity <= 4 and 1 <= n_ozone <= 4 and 1 <= n_albedo <= 4):
        raise ValueError("each axis of the sweep has 1..4 values")
    rng = np.random.default_rng(seed)
    factors = rng.uniform(1.0 - jitter, 1.0 + jitter, size=(4, 4, 4, 3))
    specs = []
    for a in range(n_turbidity):
        for b in range(n_ozone):
            for c in range(n_albedo):
                f = factors[a, b, c]
                specs.append(earth(mie_scale_height=MIE_SCALE_HEIGHTS_M[a] * f[0],
                                   ozone_dobson=OZONE_DOBSON[b] * f[1],
                                   ground_albedo=min(1.0, GROUND_ALBEDO[c] * f[2]), **earth_kw))
    return specs


def precompute(specs: Sequence[AtmosphereSpec], num_scattering_orders: int = 4, *,
               device: Optional[int] = None, sizes: Optional[Dict[str, int]] = None,
               max_in_flight: int = 16) -> List[Model]:
    """One initialised ``Model`` per spec. Up to ``max_in_flight`` precomputations are enqueued
    before the oldest is waited for, so the host never blocks between the kernels of one model."""
    models: List[Model] = []
    pending: List[Model] = []
    try:
        for spec in specs:
            m = Model.from_spec(spec, device=device, **({"sizes": sizes} if sizes else {}))
            models.append(m)
            m.InitAsync(num_scattering_orders)
            pending.append(m)
            if len(pending) >= max_in_flight:
                pending.pop(0).Wait()
        for m in pending:
            m.Wait()
    except Exception:
        for m in models:
            m.close()
        raise
    return models
