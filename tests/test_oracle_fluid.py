"""The oracle's restatement of the equivalent-fluid branch against the golden produced by the unmodified reference."""
import os

import numpy as np

from conftest import GOLDEN_DIR
from oracle import femder_oracle as O


def test_oracle_equivalent_fluid_matches_reference():
    from femder_b200.mesh import Grid
    d = np.load(os.path.join(GOLDEN_DIR, "fluid", "tet4_fluid.npz"))
    g = Grid(d["nos"], d["elem_vol"], d["elem_surf"], d["domain_index_surf"], domain_index_vol=d["domain_index_vol"])
    nf = len(d["freq"])
    mu = {int(k): d["mu"][i] for i, k in enumerate(d["group_ids"])}
    c = {1: np.ones(nf) * float(d["c0"]), 2: np.ones(nf, dtype=complex) * complex(d["cc"])}
    rho = {1: np.ones(nf) * float(d["rho0"]), 2: np.ones(nf, dtype=complex) * complex(d["rhoc"])}
    r = O.compute_fluid(g, d["freq"], mu, c, rho, d["src"], d["src_q"], float32_shape=True)
    rel = np.linalg.norm(r["pN"] - d["pN"], axis=1) / np.linalg.norm(d["pN"], axis=1)
    assert rel.max() < 1e-10                       # same arithmetic as the reference, single-precision shape functions included
    r64 = O.compute_fluid(g, d["freq"], mu, c, rho, d["src"], d["src_q"], float32_shape=False)
    rel64 = np.linalg.norm(r64["pN"] - d["pN"], axis=1) / np.linalg.norm(d["pN"], axis=1)
    assert 1e-9 < rel64.max() < 2e-5               # what that single precision costs the reference
    assert np.abs(O.tri3_unit_matrix() - O.tri3_impedance_unit_matrix()).max() < 1e-16
