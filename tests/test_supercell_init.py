"""CPU test of the host-side supercell initial condition against the compiled reference's own initial state."""
import numpy as np
import _oracle as O
from miniweatherml_b200.supercell import supercell_column


def test_supercell_column_matches_reference(golden):
    for name in ["config1_dycore10.npz", "box3d_vapor_dycore5.npz"]:
        g = golden(name)
        nz = int(g["nz"])
        bg, col = supercell_column(nz, float(g["zlen"]))
        assert np.abs(bg - g["bg"]).max() <= 1e-13 * np.abs(g["bg"]).max()
        s0 = g["s0"]
        for l, nm in [(0, "density_dry"), (1, "uvel"), (2, "vvel"), (3, "wvel"), (5, "water_vapor")]:
            ref = s0[l]
            mine = np.broadcast_to(col[nm][:, None, None], ref.shape)
            den = max(np.abs(ref).max(), 1e-300)
            assert np.abs(mine - ref).max() / den <= 1e-13, nm
        # temp carries the thermal bubble in the golden state: apply the oracle's restatement of it
        t = np.ascontiguousarray(np.broadcast_to(col["temp"][:, None, None], s0[4].shape)).copy()
        nx, ny = int(g["nx"]), int(g["ny"])
        O.perturb_thermal(t, 0, 0, float(g["xlen"]) / nx, float(g["ylen"]) / ny, float(g["zlen"]) / nz,
                          float(g["xlen"]), float(g["ylen"]))
        assert np.abs(t - s0[4]).max() / np.abs(s0[4]).max() <= 1e-13
