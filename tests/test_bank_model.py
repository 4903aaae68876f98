"""The CPU model of the shared-memory pipe that chose the staged-row layout of the ray-march kernels
(tools/probe/ms_bank_model.py) stays runnable and keeps telling the same story on a
small sample of rays: the parity split, the lane predicate and the odd slab pitch each remove wavefronts."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_model():
    spec = importlib.util.spec_from_file_location("ms_bank_model", os.path.join(ROOT, "tools", "probe", "ms_bank_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_layouts_of_the_multiple_scattering_gathers_rank_as_measured():
    m = load_model()
    rows = [(k, j) for k in (1, 16, 31) for j in (12, 75, 120)]
    natural, n0 = m.multiple_scattering(rows, False, 32, False)
    split, n1 = m.multiple_scattering(rows, True, 16, False)
    pred, n2 = m.multiple_scattering(rows, True, 16, True)
    odd, n3 = m.multiple_scattering(rows, True, 17, True)
    # ncu measured 1.27 conflict wavefronts per gather load on top of 4 for the natural order
    assert 4.9 < natural < 5.6
    assert n0 > n1 > n2 > n3
    assert n3 < 0.82 * n0


def test_single_scattering_lane_maps_rank_as_measured():
    m = load_model()
    rows = [(k, j) for k in (1, 16, 31) for j in (12, 75, 120)]
    columns, _ = m.single_scattering(rows, False, False)
    nu_lanes, _ = m.single_scattering(rows, True, False)
    split, _ = m.single_scattering(rows, True, True)
    assert columns > nu_lanes > split > 4.0
