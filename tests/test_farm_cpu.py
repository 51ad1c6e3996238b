"""File formats of the job farm (cylinder_b200/farm.py) against fixtures written by the reference's own scripts
(parallel_preprocessing.py, parallel_postprocessing.py; recorded by tests/golden/make_golden.py farm)."""
import os
import shutil

import numpy as np
import pandas as pd
import pytest

from cylinder_b200 import farm

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "farm")

# same option syntax as the reference's infile.txt (one `--option value...` per line, free spacing); values are this test's own
INFILE = """--field_type lattice
--method   sequential
--n_steps 120
--measure_every 25
--fieldsteps_per_ampstep  40  
--dims 30 20
--num_field_coeffs 2 3
--wavenumber  1
--radius 1
--amplitude 0.1  
--alpha -1.5
--C 0.5
--u 2
--n 3
--gamma 1
--kappa 0.25
--intrinsic_curvature  0
--temp .5
--temp_final .0000000001
"""


def test_preprocess_writes_the_reference_varfile(tmp_path):
    n = farm.preprocess("golden scan", "alpha", [-2, 1.1, .75], "C", [.3, 1.0, .3], outdir=str(tmp_path))
    assert n == 15
    assert open(tmp_path / "varfile.txt").read() == open(os.path.join(GOLD, "varfile_ref.txt")).read()
    assert open(tmp_path / "notes.txt").read() == open(os.path.join(GOLD, "notes_ref.txt")).read()
    names, lines = farm.read_varfile(str(tmp_path / "varfile.txt"))
    assert names == ("alpha", "C") and lines[0] == ("-2.0", "0.3") and lines[-1] == ("1.0", "0.9") and len(lines) == 15
    assert farm.title_of(names, lines[3]) == "alpha_-1.25_C_0.3"


def test_read_infile_options_and_defaults(tmp_path):
    p = tmp_path / "infile.txt"
    p.write_text(INFILE)
    a = farm.read_infile(str(p))
    assert (a.n_steps, a.field_type, a.method, a.measure_every) == (120, "lattice", "sequential", 25)
    assert (a.alpha, a.C, a.u, a.n, a.kappa, a.gamma, a.temp, a.radius, a.wavenumber) == (-1.5, .5, 2, 3, .25, 1, .5, 1, 1)
    assert a.temp_final == 1e-10 and a.dims == [30, 20] and a.num_field_coeffs == (2, 3) and a.fieldsteps_per_ampstep == 40
    assert a.amplitude == .1
    assert a.temperature_lattice == 1 and a.n_substeps == 600 and a.outdir == "."         # defaults of parallel_single_run.py:62-66
    cols, amp = farm._columns(a, ("wavenumber", "temp"), [("0.5", "0.1"), ("0.75", "0.2")])
    assert cols["wavenumber"].tolist() == [.5, .75] and cols["temperature"].tolist() == [.1, .2] and cols["alpha"].tolist() == [-1.5, -1.5]
    cols, amp = farm._columns(a, ("amplitude", "C"), [("0.5", "2"), ("-0.25", "3")])
    assert amp.tolist() == [.5, -.25] and cols["C"].tolist() == [2, 3]
    with pytest.raises(ValueError):
        farm._columns(a, ("n_steps", "C"), [("1", "2")])


def test_postprocess_equals_the_reference_tables(tmp_path):
    for f in os.listdir(GOLD):
        if f.endswith("_mean.csv"):
            shutil.copy(os.path.join(GOLD, f), tmp_path / f)
    cols = farm.postprocess(str(tmp_path))
    assert cols == ["abs_amplitude", "amplitude_squared", "field_energy"]
    for c in cols:
        ours = pd.read_csv(tmp_path / (c + ".csv"), index_col=0)
        ref = pd.read_csv(os.path.join(GOLD, c + "_ref.csv"), index_col=0)
        # the reference iterates glob() in directory order, so its row / column order is arbitrary: compare as labelled tables
        assert sorted(ours.columns) == sorted(ref.columns) and sorted(ours.index) == sorted(ref.index)
        np.testing.assert_array_equal(ours.loc[ref.index, ref.columns].to_numpy(), ref.to_numpy())
