"""CPU: the SED-realisation oracle (oracle/sed_oracle.py) against the golden outputs of the reference's own
FLUXES_ARRAYS.fluxes / integrated_luminosities, and the dense PlotSet round trip."""
import numpy as np
import pytest

from conftest import SED_CASES
from helpers import load_sed, sed_component_order, assert_close_with_specials
from agnfitter_b200 import plotset
from oracle import sed_oracle as SO


@pytest.mark.parametrize("case", SED_CASES)
def test_sed_oracle_matches_reference_golden(case):
    ms, ps, g = load_sed(case)
    stride = int(g["stride"])
    order = sed_component_order(ms)
    assert g["ref_nuLnu"].shape[1] == len(order)
    for s in range(len(g["src_z"])):
        sel = g["source_id"] == s
        z, dlum = float(g["src_z"][s]), float(g["src_dlum"][s])
        grid, fl = SO.realization_fluxes(ms, ps, g["pars"][sel], z, dlum)
        assert np.array_equal(grid, g["all_nus_rest"])
        nl = SO.fluxes_to_nuLnu(grid, fl, z, dlum)
        for ci, c in enumerate(order):
            want = g["ref_nuLnu"][sel][:, ci]
            got = nl[c][:, ::stride]
            assert np.array_equal(got == 0.0, want == 0.0), (case, c)
            assert_close_with_specials(got, want, rtol=1e-13, what="%s %s" % (case, c))
        with np.errstate(all="ignore"):
            il = np.log10(SO.integrated_luminosities([str(m) for m in g["intlum_models"]], g["intlum_ranges_hz"],
                                                     grid, nl, ms.radio)).T
        assert_close_with_specials(il, g["ref_int_lums"][sel], rtol=1e-12, what=case + " int_lums")


def test_sed_golden_covers_special_values():
    """The fixtures exercise what the kernel has to reproduce: NaN (negative template samples), exact zeros
    (10**-inf and the torus / radio cut-offs) and the -100 fill outside a template's range."""
    ms, ps, g = load_sed("classic18")
    assert np.isnan(g["ref_nuLnu"]).any() and (g["ref_nuLnu"] == 0.0).any()
    assert np.isnan(g["ref_int_lums"]).any()


def test_plotset_roundtrip_and_dered_rows():
    ms, ps, g = load_sed("xray_gamma49")
    again = plotset.plotset_from_arrays(plotset.plotset_to_arrays(ps))
    for a, b in zip(ps.families, again.families):
        assert a.name == b.name and np.array_equal(a.group, b.group)
        for x, y in zip(a.Fnu, b.Fnu):
            assert np.array_equal(x, y)
    bb = ms.fam("BB")
    d = bb.par_names.index('EBVbbb')
    assert ps.dered_mode == 2
    for t in range(bb.T):
        k = bb.keys_float[:, int(ps.dered_row[t])]
        assert k[d] == 0.0 and np.array_equal(np.delete(k, d), np.delete(bb.keys_float[:, t], d))
    assert np.array_equal(plotset.dered_rows(ms, ps.dered_mode), ps.dered_row)
    ms1, ps1, _ = load_sed("classic18")
    assert ps1.dered_mode == 1 and len(ps1.fam("SB").lognu) == 2      # DH02_CE01: two frequency grids
