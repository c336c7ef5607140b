"""Host pencil-matrix templates vs the reference's per-pencil matrices (subsystems.py:497-602), compared in
the reference's natural (un-permuted) ordering dumped by tests/golden/make_golden.py."""
import numpy as np, pytest
from scipy import sparse
from dedalus_b200 import examples
from dedalus_b200.pencils import PencilSystemBuilder


def _golden_matrix(g, tag, name):
    shape = tuple(g[f"{tag}_{name}_shape"])
    return sparse.coo_matrix((g[f"{tag}_{name}_val"], (g[f"{tag}_{name}_row"], g[f"{tag}_{name}_col"])), shape=shape).tocsr()


@pytest.mark.parametrize("case", [("rb3d_8.npz", 3, 8, 8, 1e6, [(0, 0), (0, 2), (3, 0), (1, 2)]),
                                  ("rb2d_16x16.npz", 2, 16, 16, 2e6, [(0,), (1,), (5,)])])
def test_rb_templates_match_reference(golden, case):
    fname, dim, Nh, Nz, Ra, groups = case
    g = golden(fname)
    pb = examples.rayleigh_benard(dim=dim, Nh=Nh, Nz=Nz, Rayleigh=Ra)
    builder = PencilSystemBuilder(pb['problem'])
    for grp in groups:
        tag = "pen_" + "_".join(str(k) for k in grp)
        cls = builder.find_class(grp)
        vr, vc = g[f"{tag}_valid_rows"], g[f"{tag}_valid_cols"]
        assert np.array_equal(cls.valid_rows, vr), (grp, "valid rows")
        assert np.array_equal(cls.valid_cols, vc), (grp, "valid cols")
        for name in ("M", "L"):
            ref = _golden_matrix(g, tag, name)
            mine = builder.class_matrix(cls, name, grp, restrict=False)
            # zero the invalid rows/cols like the reference's pre_left/pre_right do
            mine = sparse.diags(vr.astype(float)) @ mine @ sparse.diags(vc.astype(float))
            diff = abs(mine - ref)
            scale = abs(ref).max() if ref.nnz else 1.0
            assert diff.max() <= 1e-13 * max(scale, 1.0), (grp, name, diff.max(), scale)
            assert (abs(mine) > 0).sum() == (abs(ref) > 0).sum(), (grp, name, "pattern")
